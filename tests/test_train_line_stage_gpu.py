"""GPU parity of the training path of the 1/32 LINE-WINDOW STAGE (train_line_stage.LineStage):
the backward kernels of the "glass-structure context" (reference re-query, three diffusion rounds, reference scores,
reference-token scatter) each against torch.autograd on the same operands, then the module's forward values and every input
/ parameter gradient against torch.autograd over the CPU oracle's `swin_stage` with reference points
(WindowAttention.forward, src/models/multiscale_transformerr.py:267-332; block :646-755).

Tolerances: fp32 kernels 1e-4 relative L2 (the diffusion convolution runs as 3xTF32), bf16-output kernels 6e-3, the module a
few per cent (bf16 activations through four blocks of ~35 kernels each)."""
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


@pytest.mark.parametrize("B,T,R", [(2, 441, 40), (1, 98, 60), (3, 130, 6)])
def test_ref_requery_and_scores_bwd(B, T, R):
    """gwd_ref_requery_bwd / gwd_ref_scores_bwd == autograd of q_new = scale softmax_R(a) ref_v and a = scale q ref_k^T"""
    ops = _ops()
    heads, hd, scale = 16, 32, 32 ** -0.5
    D = heads * hd
    g = _g(B * T + R)
    a = torch.randn(B, heads, T, R, generator=g) * 2
    ref = torch.randn(B * R, 2 * D, generator=g)                        # ref_k | ref_v
    dqn = torch.randn(B * T, 3 * D, generator=g).bfloat16()               # only the first D columns are d q_new
    q = torch.randn(B * T, D, generator=g).bfloat16()
    # ---- requery
    ar, vr = a.clone().requires_grad_(True), ref[:, D:].clone().requires_grad_(True)
    v4 = vr.view(B, R, heads, hd).permute(0, 2, 1, 3)
    qn = (torch.softmax(ar, -1) @ v4) * scale                             # [B, heads, T, hd]
    qn.backward(dqn[:, :D].float().view(B, T, heads, hd).permute(0, 2, 1, 3))
    refc, dqc = ref.cuda(), dqn.cuda()
    d_kv = torch.zeros(B * R, 2 * D, device="cuda")
    d_a = ops.ref_requery_bwd(a.cuda(), refc[:, D:], 2 * D, dqc, 3 * D, d_kv[:, D:], 2 * D, B, T, heads, hd, R, scale)
    assert rel_l2(d_a, ar.grad) < 1e-4
    assert rel_l2(d_kv[:, D:], vr.grad) < 1e-4 and float(d_kv[:, :D].abs().max()) == 0.0
    # the forward kernel agrees with the same formula
    out = torch.zeros(B * T, 3 * D, dtype=torch.bfloat16, device="cuda")
    ops.ref_requery(a.cuda(), refc[:, D:], 2 * D, out, 3 * D, B, 1, T, heads, hd, R, scale)
    assert rel_l2(out[:, :D], qn.detach().permute(0, 2, 1, 3).reshape(B * T, D)) < 6e-3
    # ---- scores
    d_a0 = torch.randn(B, heads, T, R, generator=g)
    qr, kr = q.float().clone().requires_grad_(True), ref[:, :D].clone().requires_grad_(True)
    s = (qr.view(B, T, heads, hd).permute(0, 2, 1, 3) * scale) @ kr.view(B, R, heads, hd).permute(0, 2, 3, 1)
    s.backward(d_a0)
    refk = ref[:, :D].contiguous().cuda()
    ops.ref_scores_bwd(d_a0.cuda(), refk, D, q.cuda(), D, dqc, 3 * D, d_kv, 2 * D, B, T, heads, hd, R, scale)
    assert rel_l2(dqc[:, :D], qr.grad) < 6e-3
    assert rel_l2(d_kv[:, :D], kr.grad) < 1e-4 and rel_l2(d_kv[:, D:], vr.grad) < 1e-4
    a0 = torch.empty(B, heads, T, R, device="cuda")
    ops.ref_scores(q.cuda(), D, refk, D, a0, B, 1, T, heads, hd, R, scale=scale)
    assert rel_l2(a0, s.detach()) < 1e-5


@pytest.mark.parametrize("B,P,R", [(2, 441, 40), (1, 60, 60), (1, 15, 6), (1, 23, 16)])
def test_ref_diffuse_round_bwd(B, P, R):
    """gwd_ref_diffuse_dev + gwd_ref_diffuse_bwd == autograd of a + gelu(layer_norm_plane(conv3x3(a))) with the filter in the
    flat-buffer layout [tap = kx*3+ky][oc][ic]"""
    ops = _ops()
    heads = 16
    g = _g(B * P + R)
    a = torch.randn(B, heads, P, R, generator=g)
    w = torch.randn(heads, heads, 3, 3, generator=g) * 0.1
    b = torch.randn(heads, generator=g) * 0.1
    gout = torch.randn(B, heads, P, R, generator=g)
    ar, wr, br = a.clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    out = ar + F.gelu(F.layer_norm(F.conv2d(ar, wr, br, padding=1), [P, R]))
    out.backward(gout)
    w_phys = w.permute(3, 2, 0, 1).reshape(9, heads, heads).contiguous().cuda()         # [kx][ky][oc][ic]
    fwd, bwd = ops.diffuse_filter_pack(w_phys, b.cuda())
    a_out, raw, stats = ops.ref_diffuse_dev(a.cuda(), fwd, B, heads, P, R)
    assert rel_l2(a_out, out.detach()) < 2e-3      # tanh-form GELU on the forward (gwd_common.cuh), exact derivative on the backward
    dw, db = torch.zeros(9, heads, heads, device="cuda"), torch.zeros(heads, device="cuda")
    d_in = ops.ref_diffuse_bwd(gout.cuda(), raw, stats, a.cuda(), bwd, dw, db, B, heads, P, R)
    assert rel_l2(d_in, ar.grad) < 2e-3
    assert rel_l2(dw.view(3, 3, heads, heads).permute(2, 3, 1, 0), wr.grad) < 2e-3
    # the plane normalisation removes any constant shift, so the bias gradient is identically zero (round-off on both sides)
    assert float(db.abs().max()) < 1e-3 * float(dw.abs().max()) and float(br.grad.abs().max()) < 1e-3 * float(wr.grad.abs().max())


def test_ref_affine_and_scatter():
    """gwd_ref_affine(_bwd) == autograd of mu + exp(logsigma) * ref; gwd_line_ref_scatter == adjoint of gwd_line_ref_gather's
    feature sample (index_add over the window rows the points hit, shifted and un-shifted)"""
    ops = _ops()
    g = _g(5)
    rows, D = 80, 512
    ref = torch.randn(rows, 2 * D, generator=g)
    mu, ls = torch.randn(D, generator=g), torch.randn(D, generator=g) * 0.3
    d_kv = torch.randn(rows, 2 * D, generator=g)
    rr, mr, lr = ref.clone().requires_grad_(True), mu.clone().requires_grad_(True), ls.clone().requires_grad_(True)
    k = mr + lr.exp() * rr[:, :D]
    (k * d_kv[:, :D]).sum().add((rr[:, D:] * d_kv[:, D:]).sum()).backward()
    assert rel_l2(ops.ref_affine(ref.cuda(), mu.cuda(), ls.cuda(), D), k.detach()) < 1e-6
    dmu, dls = torch.zeros(D, device="cuda"), torch.zeros(D, device="cuda")
    d_ref = ops.ref_affine_bwd(d_kv.cuda(), ref.cuda(), ls.cuda(), dmu, dls, D)
    assert rel_l2(d_ref, rr.grad) < 6e-3 and rel_l2(dmu, mr.grad) < 1e-5 and rel_l2(dls, lr.grad) < 1e-5
    # scatter: gather a one-hot-free random window map, compare <gather(win), d> with <win, scatter(d)>
    B, H, W, ws, R, C = 2, 15, 20, 7, 40, 64
    Hp, Wp = 21, 21
    coords = (torch.rand(B, R, 2, generator=g) * 2.2 - 1.1).cuda()       # some points fall outside the map
    pos = torch.zeros(H * W, C, device="cuda")
    for shift in (0, 3):
        win = torch.randn(B * Hp * Wp, C, generator=g).bfloat16().cuda()
        d = torch.randn(B * R, C, generator=g).bfloat16().cuda()
        gath = ops.line_ref_gather(win, pos, coords, R, B, H, W, ws, shift, C).view(B * R, C)
        d_win = torch.zeros_like(win)
        ops.line_ref_scatter(d, coords, R, d_win, B, H, W, ws, shift, C)
        lhs, rhs = float((gath.float() * d.float()).sum()), float((win.float() * d_win.float()).sum())
        assert abs(lhs - rhs) < 2e-2 * max(abs(lhs), 1.0), (shift, lhs, rhs)


@pytest.mark.parametrize("B,H,W,center", [(2, 15, 20, False), (1, 7, 10, True), (1, 15, 20, False)])
def test_line_stage_gradients_match_oracle_autograd(B, H, W, center):
    """train_line_stage.LineStage (dense_input_proj + four line-window Swin blocks, forward + backward) against torch.autograd
    over the oracle: x32, depth0, the gradient of C5 and of every parameter (depth_pred32 gets none)"""
    _ops()
    from gwdepth_b200.engine import DEFAULT_CFG
    from gwdepth_b200.train_line_stage import LineStage
    cfg = dict(DEFAULT_CFG, with_dense_center=center)
    pre = ("dense_encoder.dense_transformer.", "dense_input_proj.", "dense_encoder.depth_pred32.")
    sd = {k: v.clone() for k, v in synth_weights().items() if k.startswith(pre)}
    g = _g(B * H + W)
    D, L = 512, H * W
    c5 = (torch.randn(B, H, W, 2048, generator=g).abs() * 0.5).bfloat16()           # post-ReLU backbone map
    R = 60 if center else 40
    ref_xy = torch.rand(B, R, 2, generator=g) * 2 - 1
    gx = torch.randn(B * L, D, generator=g).bfloat16()
    c5r = c5.float().requires_grad_(True)
    sdr = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in sd.items()}
    p = oracle.P(sdr)
    dense_in = F.conv2d(c5r.permute(0, 3, 1, 2), p["dense_input_proj.weight"], p["dense_input_proj.bias"])
    pos32 = oracle.sine_position(torch.zeros(B, H, W, dtype=torch.bool), D // 2, False)
    xo, _, _ = oracle.swin_stage(dense_in.flatten(2).permute(0, 2, 1), H, W, p.sub("dense_encoder.dense_transformer"), 4, 16, 7,
                                 ref=ref_xy.view(B, R // (3 if center else 2), -1, 2), ref_pos=pos32)
    d0 = oracle.depth_head(xo, p.sub("dense_encoder"), "depth_pred32").detach().view(B, H, W)
    (xo * gx.float().view(B, L, D)).sum().backward()
    st = LineStage({k: v.cuda() for k, v in sd.items()}, cfg)
    for k, v in st.state_dict().items():
        assert torch.equal(v.cpu(), sd[k]), k
    x32, depth0 = st.forward(c5.cuda(), ref_xy.cuda().contiguous())
    assert rel_l2(x32.view(B * L, D), xo.detach().view(B * L, D)) < 2e-2
    assert rel_l2(depth0, d0) < 5e-2        # a sigmoid near 0.05 of a 512 -> 64 -> 1 head on bf16 tokens: feeds the sampling only
    d_c5 = st.backward(gx.cuda())
    assert rel_l2(d_c5, c5r.grad.view(B * L, -1)) < 5e-2
    grads = st.grads()
    bad = {}
    for k, v in sdr.items():
        if not v.is_floating_point():
            continue
        if v.grad is None:
            assert k.startswith("dense_encoder.depth_pred32."), k
            continue
        if k.endswith("ref_attn_diffusion.bias"):      # analytically zero (the plane normalisation removes constant shifts)
            assert float(grads[k].abs().max()) < 1e-3 * float(grads[k[:-4] + "weight"].abs().max()), k
            continue
        e = rel_l2(grads[k], v.grad)
        if e > (6e-2 if H * W >= 300 else 0.1):        # 70 tokens average less bf16 noise than 300
            bad[k] = round(e, 3)
    assert not bad, bad
