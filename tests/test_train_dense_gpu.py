"""GPU parity of the dense-prediction-head training path (DensePrediction + SilogLoss + SegLoss):
the loss-gradient kernels against torch.autograd on the same fp32 inputs, then forward values, losses and every
parameter / input gradient of train_dense.DenseHead against torch.autograd over the CPU oracle's `dense_head`.

Tolerances: the loss kernels are fp32 (1e-5; their bf16 outputs to bf16 rounding); the head runs bf16 activations through
7 convolutions per branch, so gradients must agree to a few per cent in relative L2 norm."""
import pytest
import torch
import torch.nn.functional as F

from helpers import oracle, synth_weights

pytestmark = pytest.mark.gpu


def _ops():
    import gwdepth_b200  # noqa: F401
    from gwdepth_b200 import ops
    return ops


def _g(seed):
    return torch.Generator().manual_seed(seed)


def rel_l2(got, ref):
    got, ref = got.double().cpu(), ref.double().cpu()
    return float((got - ref).norm() / ref.norm().clamp_min(1e-30))


@pytest.mark.parametrize("log_only", [False, True])
@pytest.mark.parametrize("h,w,H,W", [(24, 32, 24, 32), (12, 16, 48, 64), (15, 20, 60, 80)])
def test_silog_bwd(h, w, H, W, log_only):
    """gwd_silog_sums + gwd_silog_bwd == autograd of SilogLoss on the nearest-resized ground truth (engine_glassrgbd.py:74-80)"""
    ops = _ops()
    g = _g(h + H)
    B = 3
    pred = torch.rand(B, 1, h, w, generator=g) * 8 + 0.5
    gt = torch.rand(B, 1, H, W, generator=g) * 12          # some pixels outside [0.2, 10): masked out
    pr = pred.clone().requires_grad_(True)
    loss = oracle.depth_losses([pr], gt, weights=(0.25,), log_depth_error=log_only)[0]
    loss.backward()
    sums = ops.silog_sums(pred.cuda(), gt.cuda(), log_only=log_only)
    loss_out = torch.zeros(1, device="cuda")
    d = ops.silog_bwd(pred.cuda(), gt.cuda(), sums, weight=0.25, log_only=log_only, loss_out=loss_out)
    assert abs(float(loss_out) - float(loss.detach())) < 1e-5 * abs(float(loss.detach()))
    assert rel_l2(d.view(B, 1, h, w), pr.grad) < 1e-5
    # through max_depth * sigmoid, as padded bf16 rows
    z = torch.randn(B, 1, h, w, generator=g).requires_grad_(True)
    oracle.depth_losses([10.0 * torch.sigmoid(z)], gt, weights=(1.0,), log_depth_error=log_only)[0].backward()
    p2 = (10.0 * torch.sigmoid(z.detach())).cuda()
    s2 = ops.silog_sums(p2, gt.cuda(), log_only=log_only)
    rows = ops.silog_bwd(p2, gt.cuda(), s2, log_only=log_only, sig_scale=10.0, out_cols=16)
    assert rows.shape == (B * h * w, 16) and float(rows[:, 1:].float().abs().max()) == 0.0
    assert rel_l2(rows[:, 0].float().view(B, 1, h, w), z.grad) < 4e-3


@pytest.mark.parametrize("layout", ["nchw", "nhwc"])
def test_seg_ce(layout):
    """gwd_seg_ce == nn.CrossEntropyLoss (mean over non-ignored pixels) * weight, forward and backward"""
    ops = _ops()
    g = _g(5)
    B, C, H, W = 2, 2, 20, 28
    logits = torch.randn(B, C, H, W, generator=g) * 3
    gt = (torch.rand(B, 1, H, W, generator=g) > 0.4).long()
    gt[0, 0, :2] = -100                                   # ignored pixels
    lr_ = logits.clone().requires_grad_(True)
    loss = F.cross_entropy(lr_, gt.squeeze(1)) * 2.0
    loss.backward()
    lc = logits.cuda() if layout == "nchw" else logits.permute(0, 2, 3, 1).contiguous().cuda().permute(0, 3, 1, 2)
    loss_out = torch.zeros(1, device="cuda")
    sums, d = ops.seg_ce(lc, gt.cuda(), weight=2.0, out_cols=16, loss_out=loss_out)
    assert int(sums[0]) == int((gt >= 0).sum())
    assert abs(float(loss_out) - float(loss.detach())) < 1e-5 * abs(float(loss.detach()))
    assert float(d[:, C:].float().abs().max()) == 0.0
    assert rel_l2(d[:, :C].float().view(B, H, W, C).permute(0, 3, 1, 2), lr_.grad) < 4e-3


def test_act_bwd_from_input_and_scales():
    ops = _ops()
    g = _g(8)
    x = torch.randn(500, 144, generator=g).bfloat16()
    dy = torch.randn(500, 144, generator=g).bfloat16()
    xr = x.float().requires_grad_(True)
    F.gelu(xr).backward(dy.float())
    out = ops.act_bwd(dy.cuda(), x.cuda(), ops.ACT_GELU, from_input=True)
    assert rel_l2(out, xr.grad) < 4e-3
    # y = 10 * sigmoid(z): derivative from the output, dy fp32
    z = torch.randn(700, 1, generator=g)
    y = 10 * torch.sigmoid(z)
    d = torch.randn(700, 1, generator=g)
    out = ops.act_bwd(d.cuda(), y.cuda(), ops.ACT_SIGMOID, out_cols=16, y_mul=0.1, scale=10.0).float().cpu()
    assert rel_l2(out[:, 0], (d * y * (1 - y / 10))[:, 0]) < 4e-3 and float(out[:, 1:].abs().max()) == 0.0
    out = ops.act_bwd(dy.cuda(), None, ops.ACT_NONE, scale=4.0).float().cpu()
    assert torch.equal(out, (dy.float() * 4).bfloat16().float())


def _head_inputs(B, H4, W4, seed):
    g = _g(seed)
    feat = torch.randn(B, 64, H4, W4, generator=g)
    dtok = torch.randn(B, 64, H4, W4, generator=g)
    stok = torch.randn(B, 64, H4, W4, generator=g)
    depth3 = torch.rand(B, 1, H4, W4, generator=g)
    H, W = 4 * H4, 4 * W4
    depth_gt = torch.rand(B, 1, H, W, generator=g) * 10.5 + 0.1
    seg_gt = (torch.rand(B, 1, H, W, generator=g) > 0.5).long()
    return feat, dtok, stok, depth3, depth_gt, seg_gt


def _stage_buffer(feat, dtok, stok, depth3):
    B, _, H4, W4 = feat.shape
    buf = torch.zeros(B, H4, W4, 256)
    buf[..., :64], buf[..., 64:128], buf[..., 128:192] = feat.permute(0, 2, 3, 1), dtok.permute(0, 2, 3, 1), stok.permute(0, 2, 3, 1)
    buf[..., 192] = depth3[:, 0]
    return buf.bfloat16()


def _head_weights(scale=1.0):
    sd = {k: v.clone() for k, v in synth_weights().items() if k.startswith("depth_decoder.")}
    return sd


def test_dense_head_gradients_match_oracle_autograd():
    _ops()
    from gwdepth_b200.train_dense import DenseHead
    B, H4, W4 = 2, 12, 16
    H, W = 4 * H4, 4 * W4
    feat, dtok, stok, depth3, depth_gt, seg_gt = _head_inputs(B, H4, W4, 31)
    buf = _stage_buffer(feat, dtok, stok, depth3)
    sd = _head_weights()
    # oracle (fp32 autograd) on the bf16-rounded inputs
    bufr = buf.float().requires_grad_(True)
    sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    nchw = lambda t: t.permute(0, 3, 1, 2)
    depth_o, seg_o = oracle.dense_head(nchw(bufr[..., :64]), nchw(bufr[..., 192:193]), nchw(bufr[..., 64:128]),
                                       nchw(bufr[..., 128:192]), (H, W), oracle.P(sdr, "depth_decoder."), 10.0)
    l_depth = oracle.depth_losses([depth_o], depth_gt, weights=(1.0,))[0]
    l_seg = oracle.seg_loss(seg_o, seg_gt)
    (l_depth + l_seg).backward()
    # CUDA path
    head = DenseHead({k: v.cuda() for k, v in sd.items()})
    back = head.state_dict()
    for k, v in sd.items():
        assert torch.equal(back[k].cpu(), v), k               # logical <-> physical layout round trip
    depth, seg, losses, d_buf = head.loss_and_grads(buf.cuda(), depth_gt.cuda(), seg_gt.cuda())
    assert rel_l2(depth, depth_o.detach()) < 2e-2 and rel_l2(seg, seg_o.detach()) < 2e-2
    assert abs(float(losses[0]) - float(l_depth)) < 2e-2 * abs(float(l_depth))
    assert abs(float(losses[1]) - float(l_seg)) < 2e-2 * abs(float(l_seg))
    grads = head.grads()
    worst = {}
    for k, v in sdr.items():
        worst[k] = rel_l2(grads[k], v.grad)
    bad = {k: e for k, e in worst.items() if e > 5e-2}
    assert not bad, bad
    ref_dbuf = bufr.grad.view(-1, 256)
    assert rel_l2(d_buf[:, :193], ref_dbuf[:, :193]) < 5e-2
    assert float(d_buf[:, 193:].float().abs().max()) == 0.0       # padding columns carry no gradient


def test_dense_head_training_lowers_the_loss():
    _ops()
    from gwdepth_b200.train_dense import DenseHead
    B, H4, W4 = 2, 12, 16
    feat, dtok, stok, depth3, depth_gt, seg_gt = _head_inputs(B, H4, W4, 32)
    buf = _stage_buffer(feat, dtok, stok, depth3).cuda()
    head = DenseHead({k: v.cuda() for k, v in _head_weights().items()}, lr=1e-3, max_norm=1.0)
    hist = []
    for _ in range(12):
        hist.append(head.train_step(buf, depth_gt.cuda(), seg_gt.cuda()).sum().item())
    assert hist[-1] < 0.9 * hist[0], hist
    # the bf16 mirror follows the fp32 master
    assert torch.equal(head.Wb, head.P.to(torch.bfloat16))


# ------------------------------------------------------------------------------------------ PyramidLayer (A19)
@pytest.mark.parametrize("H,W,h,w,C", [(24, 32, 3, 4, 64), (30, 40, 15, 20, 160), (17, 23, 1, 1, 32), (16, 16, 8, 8, 64), (21, 37, 5, 9, 16), (120, 160, 7, 10, 160), (60, 80, 3, 5, 160), (60, 80, 7, 10, 168)])
def test_bilinear_up_bwd(H, W, h, w, C):
    """gwd_bilinear_up_bwd == autograd of F.interpolate(bilinear, align_corners=True), reading a channel slice"""
    ops = _ops()
    g = _g(H + w)
    B = 2
    wide = torch.randn(B, H, W, C + 16, generator=g).bfloat16()
    x = torch.zeros(B, C, h, w, requires_grad=True)
    F.interpolate(x, size=(H, W), mode="bilinear", align_corners=True).backward(wide[..., 8:8 + C].float().permute(0, 3, 1, 2))
    dx = ops.bilinear_up_bwd(wide.cuda()[..., 8:8 + C], h, w)
    assert rel_l2(dx, x.grad.permute(0, 2, 3, 1)) < 4e-3


@pytest.mark.parametrize("H,W,k", [(24, 32, 2), (30, 40, 4), (120, 160, 16), (17, 23, 8)])
def test_avgpool_bwd(H, W, k):
    ops = _ops()
    g = _g(H + k)
    B, C = 2, 32
    d = torch.randn(B, H // k, W // k, C, generator=g).bfloat16()
    add = torch.randn(B, H, W, C + 8, generator=g).bfloat16()
    x = torch.zeros(B, C, H, W, requires_grad=True)
    F.avg_pool2d(x, k, k).backward(d.float().permute(0, 3, 1, 2))
    out = ops.avgpool_bwd(d.cuda(), k, H, W, add=add.cuda()[..., :C])
    assert rel_l2(out, x.grad.permute(0, 2, 3, 1) + add[..., :C].float()) < 4e-3


def test_layernorm_bwd_with_channel_padding():
    """LayerNorm over n = 30 logical channels of a 32-wide buffer (K = 30 mixture components), GELU behind it"""
    ops = _ops()
    g = _g(77)
    rows, n, C = 999, 30, 32
    z = torch.zeros(rows, C)
    z[:, :n] = torch.randn(rows, n, generator=g) * 2 + 0.3
    z = z.bfloat16()
    dy = torch.randn(rows, C, generator=g).bfloat16()
    gamma, beta = torch.rand(n, generator=g) + 0.5, torch.randn(n, generator=g) * 0.3
    zr, gr, br = z[:, :n].float().requires_grad_(True), gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    F.gelu(F.layer_norm(zr, (n,), gr, br, 1e-5)).backward(dy[:, :n].float())
    gp, bp = torch.zeros(C), torch.zeros(C)
    gp[:n], bp[:n] = gamma, beta
    dg, db = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    dz = ops.layernorm_bwd(dy.cuda(), z.cuda(), gp.cuda(), dg, db, beta=bp.cuda(), post_act=ops.ACT_GELU, n=n)
    assert rel_l2(dz[:, :n], zr.grad) < 6e-3 and float(dz[:, n:].float().abs().max()) == 0.0
    assert rel_l2(dg[:n], gr.grad) < 2e-3 and rel_l2(db[:n], br.grad) < 2e-3
    assert float(dg[n:].abs().max()) == 0.0


@pytest.mark.parametrize("stage,K,B,H,W", [(1, 30, 2, 24, 32), (2, 80, 1, 18, 24)])
def test_pyramid_gradients_match_oracle_autograd(stage, K, B, H, W):
    """train_pyramid.Pyramid (forward + backward) against torch.autograd over the oracle's `pyramid` on the same bf16
    input and cotangent: logits, d(input) and the gradient of every parameter (K = 30: channel counts padded to 16;
    K = 80: LayerNorm over 320 channels as a separate pass, 800 -> 320 convolution)"""
    _ops()
    from gwdepth_b200.train_pyramid import Pyramid
    prefix = "dense_encoder.point_based_pred%d.pyramid." % stage
    sd = {k: v.clone() for k, v in synth_weights().items() if k.startswith(prefix)}
    g = _g(K)
    Kp = (K + 15) // 16 * 16
    rg = torch.zeros(B, H, W, Kp)
    rg[..., :K] = torch.randn(B, H, W, K, generator=g)
    rg = rg.bfloat16()
    dl = torch.zeros(B, H, W, Kp)
    dl[..., :K] = torch.randn(B, H, W, K, generator=g)
    dl = dl.bfloat16()
    xr = rg[..., :K].float().permute(0, 3, 1, 2).requires_grad_(True)
    sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    ref = oracle.pyramid(xr, oracle.P(sdr, prefix))
    ref.backward(dl[..., :K].float().permute(0, 3, 1, 2))
    py = Pyramid({k: v.cuda() for k, v in sd.items()}, prefix, K)
    back = py.state_dict()
    for k, v in sd.items():
        if ".layer4." not in k:
            assert torch.equal(back[k].cpu(), v), k
    logits = py.forward(rg.cuda())
    assert rel_l2(logits[..., :K], ref.detach().permute(0, 2, 3, 1)) < 3e-2
    assert float(logits[..., K:].float().abs().max()) == 0.0 if Kp > K else True
    d_rg = py.backward(dl.cuda())
    assert rel_l2(d_rg[..., :K], xr.grad.permute(0, 2, 3, 1)) < 6e-2
    grads = py.grads()
    bad = {}
    for k, v in sdr.items():
        if ".layer4." in k:
            assert v.grad is None            # constructed by the reference, never run
            continue
        e = rel_l2(grads[k], v.grad)
        if e > 6e-2:
            bad[k] = e
    assert not bad, bad


# ------------------------------------------------------------------------------------------ PointBasedPred (A18)
@pytest.mark.parametrize("K,B,HW", [(30, 2, 24 * 32), (80, 3, 999), (20, 1, 70)])
def test_anchor_mix_bwd(K, B, HW):
    """gwd_anchor_mix_bwd == autograd of sum_k softmax_k(logits) * anchor (points_sample.py:277-279)"""
    ops = _ops()
    g = _g(K)
    Kp = (K + 15) // 16 * 16
    logits = torch.zeros(B * HW, Kp)
    logits[:, :K] = torch.randn(B * HW, K, generator=g) * 2
    logits = logits.bfloat16()
    anchor = torch.rand(B, K, generator=g)
    dpred = torch.randn(B, HW, generator=g)
    lr_ = logits[:, :K].float().view(B, HW, K).requires_grad_(True)
    ar = anchor.clone().requires_grad_(True)
    pred = (torch.softmax(lr_, dim=-1) * ar[:, None, :]).sum(-1)
    pred.backward(dpred)
    fwd = ops.anchor_mix(logits.cuda(), anchor.cuda(), B, HW, K)
    assert rel_l2(fwd, pred.detach()) < 1e-5
    dl, da = ops.anchor_mix_bwd(logits.cuda(), anchor.cuda(), dpred.cuda(), B, HW, K)
    assert rel_l2(dl[:, :K].view(B, HW, K), lr_.grad) < 4e-3
    assert Kp == K or float(dl[:, K:].float().abs().max()) == 0.0
    assert rel_l2(da, ar.grad) < 1e-4


def _sample_coords(B, K, g):
    coords = torch.rand(B, K, 2, generator=g) * 2 - 1
    coords[:, 0] = torch.tensor([-0.9999, 0.9999])        # footprints that hang over the border (zero padding)
    coords[:, 1] = torch.tensor([1.0, -1.0])
    coords[:, 2] = coords[:, 3]                            # two points on the same pixel
    return coords


@pytest.mark.parametrize("B,H,W,C,K", [(2, 24, 32, 64, 30), (1, 15, 20, 128, 80), (3, 9, 7, 16, 5)])
def test_sample_bilinear_and_scalar_bwd(B, H, W, C, K):
    """gwd_sample_bilinear_bwd / gwd_sample_scalar_bwd == autograd of F.grid_sample(bilinear, align_corners=False, zeros)"""
    ops = _ops()
    g = _g(H * W + K)
    coords = _sample_coords(B, K, g)
    d = torch.randn(B, K, C, generator=g)
    x = torch.zeros(B, C, H, W, requires_grad=True)
    F.grid_sample(x, coords.view(B, K, 1, 2), align_corners=False).backward(d.permute(0, 2, 1)[..., None])
    out = torch.full((B * H * W, C + 32), 7.0, dtype=torch.bfloat16, device="cuda")
    ops.sample_bilinear_bwd(d.cuda(), coords.cuda(), out, 16, H, W)
    assert rel_l2(out[:, 16:16 + C].view(B, H, W, C), x.grad.permute(0, 2, 3, 1)) < 4e-3
    assert float((out[:, :16].float() - 7).abs().max()) == 0.0 and float((out[:, 16 + C:].float() - 7).abs().max()) == 0.0
    ds = torch.randn(B, K, generator=g)
    add = torch.randn(B, H, W, generator=g)
    xs = torch.zeros(B, 1, H, W, requires_grad=True)
    F.grid_sample(xs, coords.view(B, K, 1, 2), align_corners=False).backward(ds.view(B, 1, K, 1))
    got = ops.sample_scalar_bwd(ds.cuda(), coords.cuda(), H, W, add=add.cuda())
    assert rel_l2(got, xs.grad[:, 0] + add) < 1e-5
    assert rel_l2(ops.sample_scalar_bwd(ds.cuda(), coords.cuda(), H, W), xs.grad[:, 0]) < 1e-5


@pytest.mark.parametrize("stage,dim,K,B,H,W", [(1, 128, 30, 2, 24, 32), (2, 64, 80, 1, 18, 24)])
def test_point_pred_gradients_match_oracle_autograd(stage, dim, K, B, H, W):
    """train_points.PointPred (forward + backward) against torch.autograd over the oracle's `point_based_pred` on the same
    bf16 stage buffer, previous depth, sample points and cotangent: the depth map, d(stage buffer), d(previous depth) and the
    gradient of every parameter (pre_proj, refer_proj, the whole pyramid)"""
    _ops()
    from gwdepth_b200.engine import sine_table
    from gwdepth_b200.train_points import PointPred
    prefix = "dense_encoder.point_based_pred%d." % stage
    sd = {k: v.clone() for k, v in synth_weights().items() if k.startswith(prefix)}
    g = _g(100 + K)
    td, width = 64, dim + 3 * 64
    buf = torch.zeros(B * H * W, width)
    buf[:, :dim + 2 * td] = torch.randn(B * H * W, dim + 2 * td, generator=g)
    buf = buf.bfloat16()
    h, w = H // 2, W // 2
    pre_depth = torch.rand(B, h, w, generator=g) * 0.9 + 0.05
    coords = _sample_coords(B, K, g)
    d_pred = torch.randn(B, H, W, generator=g)
    d_pre_own = torch.randn(B, h, w, generator=g) * 0.1
    # oracle under autograd
    bufr = buf.float().view(B, H * W, width).requires_grad_(True)
    prer = pre_depth.clone().requires_grad_(True)
    sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    pos = oracle.sine_position(torch.zeros(B, H, W, dtype=torch.bool), dim // 2, False)
    ref = oracle.point_based_pred(bufr[..., :dim], bufr[..., dim:dim + td], prer[:, None], coords.view(B, K, 1, 2), H, W, pos,
                                  oracle.P(sdr, prefix), dim)
    ref.backward(d_pred[:, None])
    # CUDA path
    pp = PointPred({k: v.cuda() for k, v in sd.items()}, prefix, dim, td, K, in_width=width)
    back = pp.state_dict()
    for k, v in sd.items():
        if ".layer4." not in k:
            assert torch.equal(back[k].cpu(), v), k
    pred = pp.forward(buf.cuda(), pre_depth.cuda(), coords.cuda(), sine_table(H, W, dim // 2, False, "cuda"), B, H, W)
    assert rel_l2(pred, ref.detach()[:, 0]) < 2e-2
    d_buf, d_pre = pp.backward(d_pred.cuda(), d_pre_depth=d_pre_own.cuda())
    ref_dbuf = bufr.grad.view(-1, width)
    assert rel_l2(d_buf[:, :dim + td], ref_dbuf[:, :dim + td]) < 6e-2
    assert float(d_buf[:, dim + td:].float().abs().max()) == 0.0       # seg token / padding columns: no gradient
    assert rel_l2(d_pre, prer.grad + d_pre_own) < 3e-2
    grads = pp.grads()
    bad = {}
    for k, v in sdr.items():
        if ".layer4." in k:
            continue
        e = rel_l2(grads[k], v.grad)
        if e > 6e-2:
            bad[k] = e
    assert not bad, bad


def test_point_pred_training_lowers_a_depth_loss():
    """a few steps of forward -> silog gradient -> backward -> shared-norm clip + AdamW on both flat buffers"""
    ops = _ops()
    from gwdepth_b200.engine import sine_table
    from gwdepth_b200.train_points import PointPred
    prefix, dim, K, B, H, W, td = "dense_encoder.point_based_pred1.", 128, 30, 2, 24, 32, 64
    sd = {k: v.cuda() for k, v in synth_weights().items() if k.startswith(prefix)}
    g = _g(9)
    buf = torch.randn(B * H * W, dim + td, generator=g).bfloat16().cuda()
    pre_depth = (torch.rand(B, H // 2, W // 2, generator=g) * 0.9 + 0.05).cuda()
    coords = (torch.rand(B, K, 2, generator=g) * 2 - 1).cuda()
    gt = (torch.rand(B, 1, H, W, generator=g) * 8 + 0.5).cuda()
    pos = sine_table(H, W, dim // 2, False, "cuda")
    pp = PointPred(sd, prefix, dim, td, K, lr=1e-3, max_norm=1.0)
    hist = []
    loss = torch.zeros(1, device="cuda")
    for _ in range(10):
        pred = pp.forward(buf, pre_depth, coords, pos, B, H, W) * 10.0                 # metres (max_depth 10)
        sums = ops.silog_sums(pred.view(B, 1, H, W), gt)
        d = ops.silog_bwd(pred.view(B, 1, H, W), gt, sums, weight=1.0, loss_out=loss)
        pp.backward(d.view(B, H, W) * 10.0)
        pp.step()
        hist.append(float(loss))
    assert hist[-1] < 0.9 * hist[0], hist
    assert torch.equal(pp.Wb, pp.P.to(torch.bfloat16)) and torch.equal(pp.pyramid.Wb, pp.pyramid.P.to(torch.bfloat16))


# ------------------------------------------------------------------------------------------ dense tail (A18-A22 chained)
def _tail_case(B, H4, W4, seed):
    g = _g(seed)
    C, td, K = 64, 64, 80
    buf = torch.zeros(B, H4, W4, 256)
    buf[..., :C + 2 * td] = torch.randn(B, H4, W4, C + 2 * td, generator=g)
    depth2 = torch.rand(B, H4 // 2, W4 // 2, generator=g) * 0.9 + 0.05
    coords = _sample_coords(B, K, g)
    H, W = 4 * H4, 4 * W4
    depth_gt = torch.rand(B, 1, H, W, generator=g) * 10.5 + 0.1
    seg_gt = (torch.rand(B, 1, H, W, generator=g) > 0.5).long()
    return buf.bfloat16(), depth2, coords, depth_gt, seg_gt


def _tail_weights():
    return {k: v.clone() for k, v in synth_weights().items()
            if k.startswith("depth_decoder.") or k.startswith("dense_encoder.point_based_pred2.")}


def test_dense_tail_gradients_match_oracle_autograd():
    """train_tail.DenseTail: point_based_pred2 -> dense head -> silog(depth_pred3) * 0.25 + silog(depth) + 2 * seg CE, against
    torch.autograd over the oracle's functions chained the same way: the three losses, d(stage buffer), d(depth_pred2) and
    every parameter gradient"""
    _ops()
    from gwdepth_b200.engine import sine_table
    from gwdepth_b200.train_tail import DenseTail
    B, H4, W4 = 2, 16, 20
    H, W, C, td, K = 4 * H4, 4 * W4, 64, 64, 80
    buf, depth2, coords, depth_gt, seg_gt = _tail_case(B, H4, W4, 41)
    sd = _tail_weights()
    bufr = buf.float().requires_grad_(True)
    d2r = depth2.clone().requires_grad_(True)
    sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    tok = bufr.view(B, H4 * W4, 256)
    pos = oracle.sine_position(torch.zeros(B, H4, W4, dtype=torch.bool), C // 2, False)
    depth3_o = oracle.point_based_pred(tok[..., :C], tok[..., C:C + td], d2r[:, None], coords.view(B, K, 1, 2), H4, W4, pos,
                                       oracle.P(sdr, "dense_encoder.point_based_pred2."), C)
    nchw = lambda t: t.permute(0, 3, 1, 2)
    depth_o, seg_o = oracle.dense_head(nchw(bufr[..., :C]), depth3_o, nchw(bufr[..., C:C + td]), nchw(bufr[..., C + td:C + 2 * td]),
                                       (H, W), oracle.P(sdr, "depth_decoder."), 10.0)
    l3, l4 = oracle.depth_losses([depth3_o, depth_o], depth_gt, weights=(0.25, 1.0))
    ls = oracle.seg_loss(seg_o, seg_gt)
    (l3 + l4 + ls).backward()
    tail = DenseTail({k: v.cuda() for k, v in sd.items()})
    depth3, depth, seg, losses, d_buf, d_depth2 = tail.loss_and_grads(buf.cuda(), depth2.cuda(), coords.cuda(),
                                                                      sine_table(H4, W4, C // 2, False, "cuda"),
                                                                      depth_gt.cuda(), seg_gt.cuda())
    assert rel_l2(depth3, depth3_o.detach()[:, 0]) < 2e-2 and rel_l2(depth, depth_o.detach()) < 2e-2
    for got, want in zip(losses.tolist(), (l3, l4, ls)):
        assert abs(got - float(want)) < 2e-2 * abs(float(want)), (losses.tolist(), float(l3), float(l4), float(ls))
    ref_dbuf = bufr.grad.view(-1, 256)
    assert rel_l2(d_buf[:, :C + 2 * td], ref_dbuf[:, :C + 2 * td]) < 6e-2
    assert float(d_buf[:, C + 2 * td:].float().abs().max()) == 0.0
    assert rel_l2(d_depth2, d2r.grad) < 6e-2
    grads = tail.grads()
    bad = {}
    for k, v in sdr.items():
        if ".layer4." in k:
            continue
        e = rel_l2(grads[k], v.grad)
        if e > 7e-2:
            bad[k] = e
    assert not bad, bad


def test_dense_tail_training_keeps_unmapped_columns_zero_and_lowers_the_loss():
    """weights that read the shared stage buffer through a column map must not grow entries for the other consumers'
    channels (their gradient is masked), and a dozen steps lower the summed loss"""
    _ops()
    from gwdepth_b200.engine import sine_table
    from gwdepth_b200.train_tail import DenseTail
    B, H4, W4 = 2, 16, 20
    buf, depth2, coords, depth_gt, seg_gt = _tail_case(B, H4, W4, 42)
    tail = DenseTail({k: v.cuda() for k, v in _tail_weights().items()}, lr=1e-3, max_norm=1.0)
    args = (buf.cuda(), depth2.cuda(), coords.cuda(), sine_table(H4, W4, 32, False, "cuda"), depth_gt.cuda(), seg_gt.cuda())
    hist = [tail.train_step(*args).sum().item() for _ in range(12)]
    assert hist[-1] < 0.95 * hist[0], hist
    for m in tail.modules():
        for short in m.index:
            phys = m.view(m.P, short)
            assert torch.equal(m._to_physical(short, m._to_logical(short, phys)), phys), short   # nothing outside the logical entries
        assert torch.equal(m.Wb, m.P.to(torch.bfloat16))

