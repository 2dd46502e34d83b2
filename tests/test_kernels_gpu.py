"""GPU parity of every non-GEMM C-ABI kernel against the CPU oracle / plain torch fp32 on the same seeded inputs.
Index and selection kernels must be bit-exact; floating-point kernels carry their tolerance next to the assert."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from helpers import golden, oracle

pytestmark = pytest.mark.gpu


def _ops():
    import gwdepth_b200  # noqa: F401
    from gwdepth_b200 import ops
    return ops


def _g(seed):
    return torch.Generator().manual_seed(seed)


def _bf(t):
    return t.bfloat16()


def close(got, ref, tol, what=""):
    got, ref = got.float().cpu(), ref.float().cpu()
    err = (got - ref).abs().max().item()
    scale = max(1.0, ref.abs().max().item())
    assert err <= tol * scale, "%s: max err %.4g > %.4g" % (what, err, tol * scale)


# ------------------------------------------------------------------------------------------ attention
@pytest.mark.parametrize("B,Lq,Lk,heads,hd,pad", [(2, 300, 300, 8, 32, False), (3, 100, 300, 8, 32, True), (2, 100, 100, 8, 32, False),
                                                 (1, 37, 53, 4, 16, True),
                                                 # long key axes (960x1280: L = 1 200): key tiles + online soft-max on tcgen05
                                                 (2, 1200, 1200, 8, 32, False), (2, 100, 1200, 8, 32, True), (3, 700, 700, 8, 32, True),
                                                 (1, 513, 481, 8, 32, False)])
def test_attention_detr(B, Lq, Lk, heads, hd, pad):
    ops = _ops()
    g = _g(Lq + Lk + hd)
    E = heads * hd
    q, k, v = (_bf(torch.randn(B, L, E, generator=g)) for L in (Lq, Lk, Lk))
    kpm = None
    if pad:
        kpm = torch.zeros(B, Lk, dtype=torch.bool)
        kpm[:, Lk - 7:] = True
        kpm[0, 3] = True
    qh = q.float().view(B, Lq, heads, hd).permute(0, 2, 1, 3)
    kh = k.float().view(B, Lk, heads, hd).permute(0, 2, 1, 3)
    vh = v.float().view(B, Lk, heads, hd).permute(0, 2, 1, 3)
    s = qh @ kh.transpose(-1, -2)
    if pad:
        s = s.masked_fill(kpm[:, None, None, :], float("-inf"))
    ref = (torch.softmax(s, -1) @ vh).permute(0, 2, 1, 3).reshape(B, Lq, E)
    o = torch.empty(B * Lq, E, dtype=torch.bfloat16, device="cuda")
    ops.attention(q.cuda().view(-1, E), k.cuda().view(-1, E), v.cuda().view(-1, E), o, items=B, heads=heads, Lq=Lq, Lk=Lk,
                  hd=hd, q_strides=(Lq * E, E), k_strides=(Lk * E, E), v_strides=(Lk * E, E), o_strides=(Lq * E, E),
                  key_padding=kpm.to(torch.uint8).cuda() if pad else None)
    close(o.view(B, Lq, E), ref, 1e-2, "attention")     # bf16 output rounding


@pytest.mark.parametrize("hd,heads,shifted", [(32, 16, True), (16, 16, False), (8, 16, True), (4, 16, True)])
def test_attention_window(hd, heads, shifted):
    ops = _ops()
    g = _g(hd)
    N, nW, B = 49, 6, 2
    C = heads * hd
    qkv = _bf(torch.randn(B * nW * N, 3 * C, generator=g))
    bias = torch.randn(heads, N, N, generator=g)
    mask = torch.where(torch.rand(nW, N, N, generator=g) > 0.7, torch.tensor(-100.0), torch.tensor(0.0)) if shifted else None
    x = qkv.float().view(B * nW, N, 3, heads, hd).permute(2, 0, 3, 1, 4)
    ref = oracle.softmax_attention(x[0], x[1], x[2], bias, mask, nW).transpose(1, 2).reshape(B * nW * N, C)
    d = qkv.cuda()
    o = torch.empty(B * nW * N, C, dtype=torch.bfloat16, device="cuda")
    ops.attention(d, d[:, C:], d[:, 2 * C:], o, items=B * nW, heads=heads, Lq=N, Lk=N, hd=hd, q_strides=(N * 3 * C, 3 * C),
                  k_strides=(N * 3 * C, 3 * C), v_strides=(N * 3 * C, 3 * C), o_strides=(N * C, C), bias=bias.cuda(),
                  mask=mask.cuda() if shifted else None)
    close(o, ref, 1e-2, "window attention")


def test_token_attention():
    """class-token channel attention vs the oracle's formulation (multiscale_transformerr.py:561-578)"""
    ops = _ops()
    g = _g(11)
    items, N, heads, td_all, C = 5, 49, 16, 64, 128
    tC = C + 2 * td_all
    dq, sq = _bf(torch.randn(items * N, td_all, generator=g)), _bf(torch.randn(items * N, td_all, generator=g))
    kv = _bf(torch.randn(items * N, 2 * tC, generator=g))
    scale = (C // heads) ** -0.5

    def ref_one(tq):
        q = tq.float().view(items, N, heads, td_all // heads).permute(0, 2, 1, 3) * scale
        tk = kv.float()[:, :tC].reshape(items, N, heads, tC // heads).permute(0, 2, 1, 3)
        tv = kv.float()[:, tC:].reshape(items, N, heads, tC // heads).permute(0, 2, 1, 3)
        a = torch.softmax(q.transpose(-2, -1) @ tk, dim=-1)
        return (a @ tv.transpose(-2, -1)).reshape(items, -1, N).permute(0, 2, 1).reshape(items * N, td_all)

    dout = torch.empty(items * N, td_all, dtype=torch.bfloat16, device="cuda")
    sout = torch.empty_like(dout)
    kvd = kv.cuda()
    ops.token_attention(dq.cuda(), sq.cuda(), kvd, kvd[:, tC:], dout, sout, items=items, N=N, heads=heads, td=td_all // heads,
                        tc=tC // heads, q_rs=td_all, k_rs=2 * tC, v_rs=2 * tC, o_rs=td_all, scale=scale)
    close(dout, ref_one(dq), 1e-2, "depth token")
    close(sout, ref_one(sq), 1e-2, "seg token")


def test_line_requery_chain():
    """ref_scores -> 3 x ref_diffuse -> ref_requery vs the oracle's WindowAttention re-query (mst.py:295-310)"""
    ops = _ops()
    g = _g(5)
    B, nW, N, heads, hd, R = 2, 9, 49, 16, 32, 40
    C = heads * hd
    P = nW * N
    q = _bf(torch.randn(B * P, C, generator=g) * 0.3)
    refkv = torch.randn(B * R, 2 * C, generator=g) * 0.5
    w = torch.randn(heads, heads, 3, 3, generator=g) * 0.1
    b = torch.randn(heads, generator=g) * 0.1
    scale = hd ** -0.5
    qh = q.float().view(B, nW, N, heads, hd).permute(0, 3, 1, 2, 4).reshape(B, heads, P, hd)
    rk = refkv[:, :C].reshape(B, R, heads, hd).permute(0, 2, 1, 3)
    rv = refkv[:, C:].reshape(B, R, heads, hd).permute(0, 2, 1, 3)
    a = qh @ rk.transpose(-1, -2)
    for _ in range(3):
        a = a + F.gelu(F.layer_norm(F.conv2d(a, w, b, padding=1), [P, R]))
    ref = ((torch.softmax(a, -1) @ rv) * scale).permute(0, 2, 1, 3).reshape(B * P, C)

    a0 = torch.empty(B, heads, P, R, device="cuda")
    a1, raw = torch.empty_like(a0), torch.empty_like(a0)
    st = torch.empty(B * heads * 2, dtype=torch.float64, device="cuda")
    rd = refkv.cuda()
    ops.ref_scores(q.cuda(), C, rd, 2 * C, a0, B, nW, N, heads, hd, R)
    ops.ref_diffuse(a0, a1, w.contiguous(), b.contiguous(), raw, st, B, heads, P, R)
    ops.ref_diffuse(a1, a0, w.contiguous(), b.contiguous(), raw, st, B, heads, P, R)
    ops.ref_diffuse(a0, a1, w.contiguous(), b.contiguous(), raw, st, B, heads, P, R)
    close(a1, a, 2e-4, "diffused reference scores")      # fp32 path
    out = torch.empty(B * P, C, dtype=torch.bfloat16, device="cuda")
    ops.ref_requery(a1, rd[:, C:], 2 * C, out, C, B, nW, N, heads, hd, R, scale)
    close(out, ref, 1e-2, "re-queried q")


# ------------------------------------------------------------------------------------------ bandwidth kernels
@pytest.mark.parametrize("rows,C,n", [(1000, 256, 256), (777, 64, 64), (513, 320, 320), (300, 128, 120), (64, 512, 512), (99, 80, 80), (1001, 192, 190), (35, 448, 448), (130, 384, 384)])
def test_layernorm_and_add(rows, C, n):
    ops = _ops()
    g = _g(rows + C)
    x, r = _bf(torch.randn(rows, C, generator=g)), _bf(torch.randn(rows, C, generator=g))
    gam, bet = 1 + 0.1 * torch.randn(n, generator=g), 0.1 * torch.randn(n, generator=g)
    ref = F.gelu(F.layer_norm((x.float() + r.float())[:, :n], (n,), gam, bet, 1e-5))
    out = ops.layernorm(x.cuda(), ops.pad_vec(gam.cuda(), C), ops.pad_vec(bet.cuda(), C), res=r.cuda(), act=ops.ACT_GELU, n=n)
    close(out[:, :n], ref, 1e-2, "layernorm")
    assert (out[:, n:] == 0).all()
    period = 50
    add = _bf(torch.randn(period, C, generator=g))
    out2 = ops.add_rows(x.cuda(), add.cuda(), period)
    idx = torch.arange(rows) % period
    close(out2, x.float() + add.float()[idx], 1e-2, "add_rows")


@pytest.mark.parametrize("H,W,C,shift", [(15, 20, 512, 0), (15, 20, 512, 3), (30, 40, 256, 3), (17, 23, 64, 3), (14, 21, 128, 0)])
def test_window_gather_merge(H, W, C, shift):
    ops = _ops()
    g = _g(H * W + C + shift)
    B, ws = 2, 7
    x = _bf(torch.randn(B, H * W, C, generator=g))
    gam, bet = 1 + 0.1 * torch.randn(C, generator=g), 0.1 * torch.randn(C, generator=g)
    ln = F.layer_norm(x.float(), (C,), gam, bet, 1e-5)
    ref_w = oracle.to_windows(oracle.pad_and_shift(ln, H, W, ws, shift), ws).reshape(-1, C)
    got_w = ops.window_gather(x.cuda().view(B, H, W, C), B, H, W, ws, shift, gam.cuda(), bet.cuda())
    close(got_w, ref_w, 1e-2, "window gather")
    Hp, Wp = math.ceil(H / ws) * ws, math.ceil(W / ws) * ws
    win = _bf(torch.randn(B * Hp * Wp, C, generator=g))
    merged = x.float() + oracle.unshift_and_crop(oracle.from_windows(win.float().view(-1, ws * ws, C), ws, B, Hp, Wp), H, W, shift).reshape(B, H * W, C)
    out, out_ln = ops.window_merge(win.cuda(), x.cuda().view(B * H * W, C), B, H, W, ws, shift, gam.cuda(), bet.cuda(), want_ln=True)
    close(out, merged.view(-1, C), 1e-2, "window merge")
    close(out_ln, F.layer_norm(merged, (C,), gam, bet, 1e-5).view(-1, C), 2e-2, "window merge LN")


def test_resampling():
    ops = _ops()
    g = _g(3)
    B, h, w, C = 2, 15, 20, 64
    x = _bf(torch.randn(B, h, w, C, generator=g))
    nchw = x.float().permute(0, 3, 1, 2)
    add = _bf(torch.randn(B, 30, 40, C, generator=g))
    up = ops.upsample_nearest(x.cuda(), 30, 40, add=add.cuda())
    close(up, F.interpolate(nchw, size=(30, 40), mode="nearest").permute(0, 2, 3, 1) + add.float(), 1e-2, "nearest")
    up2 = ops.upsample_nearest(x.cuda(), 37, 41)    # non-integer ratio: legacy floor(dst*in/out) rule
    close(up2, F.interpolate(nchw, size=(37, 41), mode="nearest").permute(0, 2, 3, 1), 1e-2, "nearest ragged")
    big = _bf(torch.randn(B, 33, 47, C, generator=g))
    for k in (16, 8, 4, 2):
        close(ops.avgpool(big.cuda(), k), F.avg_pool2d(big.float().permute(0, 3, 1, 2), k, k).permute(0, 2, 3, 1), 1e-2, "avgpool")
    out = torch.zeros(B, 33, 47, 3 * C, dtype=torch.bfloat16, device="cuda")
    ops.bilinear_up_into(x.cuda(), out, C, 33, 47)
    ref = F.interpolate(nchw, size=(33, 47), mode="bilinear", align_corners=True).permute(0, 2, 3, 1)
    close(out[..., C:2 * C], ref, 1e-2, "bilinear")
    assert (out[..., :C] == 0).all() and (out[..., 2 * C:] == 0).all()


@pytest.mark.parametrize("H,W,C", [(33, 47, 64), (120, 160, 160), (60, 80, 168), (16, 16, 8)])
def test_pyramid_pool_and_upsample(H, W, C):
    """the one-pass pooling pyramid (k = 16, 8, 4, 2) == the per-k kernel (k = 2 bit for bit, the tree-summed levels to the bf16
    rounding of an fp32 re-association) == F.avg_pool2d; the four-branch bilinear up-sampling == four gwd_bilinear_up launches bit
    for bit.  Extents include blocks cut by the border and channel counts that are not a multiple of 32"""
    ops = _ops()
    g = _g(H + C)
    B = 2
    big = _bf(torch.randn(B, H, W, C + 16, generator=g)).cuda()     # pooled channels are a slice of a wider buffer
    pyr = ops.avgpool_pyramid(big, C=C)
    for k, got in zip((16, 8, 4, 2), pyr):
        ref = F.avg_pool2d(big[..., :C].float().permute(0, 3, 1, 2), k, k).permute(0, 2, 3, 1)
        assert tuple(got.shape) == tuple(ref.shape)
        close(got, ref, 4e-3, "pyramid pool %d" % k)
        one = ops.avgpool(big, k, C=C)
        if k == 2:
            assert torch.equal(got, one)
        else:
            close(got, one, 4e-3, "pyramid pool %d vs single" % k)
    out = torch.zeros(B, H, W, 5 * C + 8, dtype=torch.bfloat16, device="cuda")
    ref = torch.zeros_like(out)
    ops.bilinear_up4_into(pyr, out, C, H, W)
    for j, p_ in enumerate(pyr, start=1):
        ops.bilinear_up_into(p_, ref, j * C, H, W)
    assert torch.equal(out, ref)


def test_point_sampling_and_mixture():
    ops = _ops()
    g = _g(4)
    B, H, W, C, K = 2, 12, 17, 48, 30
    feat = _bf(torch.randn(B, H, W, 2 * C, generator=g))
    table = torch.randn(H * W, C, generator=g)
    coords = torch.rand(B, K, 1, 2, generator=g) * 2.2 - 1.1          # some points fall outside
    f = feat.float()[..., C:].permute(0, 3, 1, 2)
    t = table.view(1, H, W, C).permute(0, 3, 1, 2).expand(B, -1, -1, -1)
    ref = (F.grid_sample(f, coords, align_corners=False) + F.grid_sample(t, coords, align_corners=False)).flatten(2).permute(0, 2, 1)
    got = ops.sample_bilinear(feat.cuda(), C, table.cuda(), B, H, W, C, coords.cuda().contiguous(), K)
    close(got, ref, 1e-5, "bilinear point sample")
    # per-image tables (padded batches): image b samples table[b]
    tables = torch.randn(B, H * W, C, generator=g)
    tb = tables.view(B, H, W, C).permute(0, 3, 1, 2)
    ref_b = (F.grid_sample(f, coords, align_corners=False) + F.grid_sample(tb, coords, align_corners=False)).flatten(2).permute(0, 2, 1)
    got_b = ops.sample_bilinear(feat.cuda(), C, tables.cuda(), B, H, W, C, coords.cuda().contiguous(), K)
    close(got_b, ref_b, 1e-5, "bilinear point sample, per-image table")
    depth = torch.rand(B, 9, 11, generator=g)
    ref_a = F.grid_sample(depth[:, None], coords, align_corners=False).flatten(1)
    close(ops.sample_scalar(depth.cuda(), coords.cuda().contiguous(), K), ref_a, 1e-6, "anchor depth")
    logits = _bf(torch.randn(B, H * W, 32, generator=g))
    ref_m = (torch.softmax(logits.float()[..., :K], -1) * ref_a[:, None, :]).sum(-1)
    close(ops.anchor_mix(logits.cuda(), ref_a.cuda().contiguous(), B, H * W, K), ref_m, 1e-5, "anchor mixture")


@pytest.mark.parametrize("shift", [0, 3])
def test_line_reference_gather(shift):
    """nearest sampling of the shifted, padded, normalised map + shifted position table (mst.py:676-701)"""
    ops = _ops()
    g = _g(6 + shift)
    B, H, W, C, ws, R = 2, 15, 20, 64, 7, 40
    x = _bf(torch.randn(B, H * W, C, generator=g))
    pos = torch.randn(H * W, C, generator=g)
    ref_pts = torch.rand(B, R // 2, 2, 2, generator=g) * 2 - 1
    xs = oracle.pad_and_shift(x.float(), H, W, ws, shift)
    Hp, Wp = xs.shape[1:3]
    pos_nchw = pos.view(1, H, W, C).permute(0, 3, 1, 2).expand(B, -1, -1, -1)
    if shift:
        rr = torch.zeros_like(ref_pts)
        rr[..., 0] = ref_pts[..., 0] - (shift / (Wp - 1)) * 2
        rr[..., 1] = ref_pts[..., 1] - (shift / (Hp - 1)) * 2
        rr = torch.where(rr < -1, -1 - (1 + rr), rr)
        rpos = torch.roll(pos_nchw, shifts=(-shift, -shift), dims=(2, 3))
    else:
        rr, rpos = ref_pts, pos_nchw
    ref = (oracle.nearest_points(xs.permute(0, 3, 1, 2), rr) + oracle.nearest_points(rpos, rr)).reshape(B, C, -1).permute(0, 2, 1)
    xw = ops.window_gather(x.cuda().view(B, H, W, C), B, H, W, ws, shift)
    got = ops.line_ref_gather(xw, pos.cuda(), ref_pts.reshape(B, R, 2).cuda().contiguous(), R, B, H, W, ws, shift, C)
    close(got, ref, 1e-2, "line reference tokens")


# ------------------------------------------------------------------------------------------ selection / reduction kernels
@pytest.mark.parametrize("h,w,K,seed", [(15, 20, 30, 0), (30, 40, 80, 1), (7, 10, 30, 2), (14, 20, 30, 3)])
def test_certain_sample_bit_exact(h, w, K, seed):
    ops = _ops()
    g = _g(100 + seed)
    B = 3
    small = torch.rand(B, 1, h, w, generator=g)
    large = torch.rand(B, 1, 2 * h, 2 * w, generator=g)
    if seed == 2:
        large = large * 0.08            # almost everything in the first bin: exercises the repeat / complement rules
    if seed == 3:
        large[1] = 2.0                  # no pixel in any bin for image 1: global top-K fallback
    interval = [0.1, 0.3, 0.5, 0.7, 0.9]
    ref_xy, ref_idx = oracle.certain_sample(small, large, K, interval, 1e-4)
    xy, idx = ops.certain_sample(small[:, 0].cuda().contiguous(), large[:, 0].cuda().contiguous(), K, [1e-4] + interval + [1.0])
    assert torch.equal(idx.cpu().long(), ref_idx), "sample indices differ"
    assert torch.equal(xy.cpu(), ref_xy), "sample coordinates differ"


def test_match_cost_and_assignment():
    ops = _ops()
    import gwdepth_b200  # noqa: F401
    from gwdepth_b200 import model as M
    g = _g(9)
    B, Q = 4, 100
    logits = torch.randn(B, Q, 2, generator=g) * 2
    lines = torch.rand(B, Q, 6, generator=g)
    targets = [{"lines": torch.rand(12 + 5 * b, 6, generator=g), "labels": torch.zeros(12 + 5 * b, dtype=torch.int64)} for b in range(B)]
    ref = oracle.matcher_cost(logits, lines, [t["lines"] for t in targets], 1.0, 5.0)
    matcher = M.HungarianMatcher_Line(cost_class=1.0, cost_line=5.0)
    out = {"pred_logits": logits.cuda(), "pred_lines": lines.cuda()}
    tg = [{k: v.cuda() for k, v in t.items()} for t in targets]
    got = matcher.cost_matrices(out, tg)
    for a, b_ in zip(got, ref):
        assert (a - b_).abs().max().item() <= 2e-6          # fp32, same operation order up to exp rounding
    mine = matcher(out, tg)
    theirs = oracle.hungarian(ref)
    for (i, j), (ri, rj), c in zip(mine, theirs, ref):
        assert torch.equal(i, ri) and torch.equal(j, rj), "assignment differs"        # random uniform lines: no exact L1 ties


def test_depth_metrics_and_silog():
    ops = _ops()
    gld = golden("depth_metrics.npz")
    pred, gt = torch.from_numpy(gld["pred"]), torch.from_numpy(gld["gt"])
    got = ops.depth_metrics(pred.cuda(), gt.cuda()).cpu().numpy()
    assert np.allclose(got, gld["metrics"], rtol=2e-5, atol=1e-7), (got, gld["metrics"])   # reference sums in fp32 (numpy)
    g = _g(12)
    p = torch.rand(2, 1, 30, 40, generator=g) * 0.9 + 0.05
    d = torch.rand(2, 1, 120, 160, generator=g) * 11.0
    s = ops.silog_sums(p.cuda(), d.cuda(), 0.2, 10.0, False).cpu()
    loss = math.sqrt(s[2] / s[0] - 0.85 * (s[1] / s[0]) ** 2) * 10.0
    ref = float(oracle.depth_losses([p], d, weights=(1.0,))[0])
    assert abs(loss - ref) <= 1e-4 * abs(ref), (loss, ref)


# ------------------------------------------------------------------------------------------ ResNet stem
@pytest.mark.parametrize("B,H,W", [(2, 64, 96), (1, 480, 640), (2, 50, 70), (1, 33, 47)])
def test_stem_conv_pool(B, H, W):
    """7x7/2 conv + folded BN shift + ReLU + 3x3/2 max-pool in one launch vs torch (bf16-rounded operands, fp32 math);
    sizes include tiles cut by the right / bottom border and odd extents"""
    import torch.nn.functional as F
    ops = _ops()
    g = _g(H + W)
    x = torch.randn(B, 3, H, W, generator=g)
    w = torch.randn(64, 3, 7, 7, generator=g) * 0.1
    shift = torch.randn(64, generator=g) * 0.2
    wp, bp = ops.pack_stem(w.cuda(), shift.cuda())
    got = ops.stem_conv_pool(x.cuda().contiguous(), wp, bp)
    conv = F.conv2d(_bf(x).float(), _bf(w).float(), shift, stride=2, padding=3)
    ref = F.max_pool2d(_bf(F.relu(conv)).float(), 3, 2, 1).permute(0, 2, 3, 1)
    assert tuple(got.shape) == tuple(ref.shape)
    close(got, ref, 1e-2, "stem")


def test_seg_confusion_bit_exact():
    """argmax + confusion matrix of the segmentation evaluation == torch.argmax + np.bincount (src/util/metrics.py:43-78),
    on NCHW logits and on the channels-last view the forward returns, with ignored pixels and exact ties"""
    ops = _ops()
    g = _g(31)
    B, C, H, W = 3, 2, 37, 53
    logits = torch.randn(B, C, H, W, generator=g)
    logits[:, 1, ::5] = logits[:, 0, ::5]                      # exact ties -> class 0, as torch.argmax
    gt = (torch.rand(B, 1, H, W, generator=g) > 0.4).long()
    gt[:, :, ::7, ::3] = 255                                   # ignored pixels
    pred = logits.argmax(1).reshape(-1).numpy()
    gtn = gt.reshape(-1).numpy()
    keep = gtn != 255
    ref = np.bincount(gtn[keep] * C + pred[keep], minlength=C * C).reshape(C, C)
    got = ops.seg_confusion(logits.cuda(), gt.cuda())
    assert np.array_equal(got.cpu().numpy(), ref)
    nhwc = logits.permute(0, 2, 3, 1).contiguous().cuda().permute(0, 3, 1, 2)      # what GlassRGBD.forward returns
    acc = ops.seg_confusion(nhwc, gt.cuda(), confusion=got.clone())
    assert np.array_equal(acc.cpu().numpy(), 2 * ref)
    iou, pix, macc, miou = ops.seg_scores(got)
    cm = ref.astype(np.float64)
    tp, pos, res = np.diag(cm), cm.sum(1), cm.sum(0)
    assert np.allclose(iou.cpu().numpy(), tp / np.maximum(1.0, pos + res - tp) * 100)
    assert abs(float(pix) - tp.sum() / pos.sum() * 100) < 1e-9 and abs(float(miou) - (tp / np.maximum(1.0, pos + res - tp)).mean() * 100) < 1e-9


def test_dense_evaluator_batches_equal_the_batch_1_loop():
    """evaluation.DenseEvaluator over batches of 3 and 2 == the reference's batch-1 bookkeeping (oracle.depth_metrics per
    image, numpy confusion matrix over all images)"""
    import gwdepth_b200  # noqa: F401
    from gwdepth_b200 import evaluation
    g = _g(41)
    N, H, W = 5, 48, 64
    pred = torch.rand(N, 1, H, W, generator=g) * 11 - 0.2          # some values outside [1e-3, 10]
    pred[0, 0, 0, 0], pred[1, 0, 1, 1] = float("nan"), float("inf")
    gt = torch.rand(N, 1, H, W, generator=g) * 10.5
    seg = torch.randn(N, 2, H, W, generator=g)
    seg_gt = (torch.rand(N, 1, H, W, generator=g) > 0.5).long()
    ev = evaluation.DenseEvaluator()
    for sl in (slice(0, 3), slice(3, 5)):
        ev.update({"pred_depth": [None, pred[sl].cuda()], "pred_seg": seg[sl].cuda()}, gt[sl].cuda(), seg_gt[sl].cuda())
    got = ev.summary()
    ref = torch.stack([torch.as_tensor(oracle.depth_metrics(pred[i, 0], gt[i, 0])) for i in range(N)]).double().mean(0)
    for k, name in enumerate(evaluation.DEPTH_METRICS):
        assert abs(got[name] - float(ref[k])) <= 1e-5 * max(1.0, abs(float(ref[k]))), name
    p, t = seg.argmax(1).reshape(-1).numpy(), seg_gt.reshape(-1).numpy()
    cm = np.bincount(t * 2 + p, minlength=4).reshape(2, 2).astype(np.float64)
    tp, pos, res = np.diag(cm), cm.sum(1), cm.sum(0)
    assert abs(got["Mean IU"] - (tp / np.maximum(1.0, pos + res - tp)).mean() * 100) < 1e-9
    assert abs(got["Glass"] - tp[1] / max(1.0, pos[1] + res[1] - tp[1]) * 100) < 1e-9 and got["images"] == N


def test_images_to_batch_bit_exact():
    """uint8 HWC images -> normalised, padded NCHW batch + mask == torchvision's to_tensor / normalize + the reference's
    nested_tensor_from_tensor_list, bit for bit (uniform batch and a ragged list)"""
    import gwdepth_b200  # noqa: F401
    from gwdepth_b200 import model as M
    ops = _ops()
    g = _g(51)
    mean, std = torch.tensor(ops.IMAGE_MEAN), torch.tensor(ops.IMAGE_STD)

    def reference(u8):      # torchvision.transforms.functional.to_tensor + normalize
        t = u8.permute(2, 0, 1).contiguous().to(torch.float32).div(255)
        return t.sub(mean[:, None, None]).div(std[:, None, None])
    batch = torch.randint(0, 256, (3, 40, 56, 3), generator=g, dtype=torch.uint8)
    out, mask, padded = ops.images_to_batch(batch.cuda())
    assert not padded and not bool(mask.any())
    assert torch.equal(out.cpu(), torch.stack([reference(b) for b in batch]))
    ragged = [torch.randint(0, 256, (h, w, 3), generator=g, dtype=torch.uint8) for h, w in ((40, 56), (33, 56), (40, 21))]
    out, mask, padded = ops.images_to_batch([t.cuda() for t in ragged])
    nt = M.nested_tensor_from_tensor_list([reference(t) for t in ragged])
    assert padded and torch.equal(out.cpu(), nt.tensors) and torch.equal(mask.cpu(), nt.mask)


@pytest.mark.parametrize("B,Q,num_ref,npts", [(3, 100, 20, 2), (2, 100, 20, 3), (1, 37, 37, 2)])
def test_select_lines_matches_topk_gather(B, Q, num_ref, npts):
    """gwd_select_lines == torch.topk over the raw line logit + gather + * 2 - 1 (multiscale_transformerr.py:1165-1179),
    index-identical (descending order), ties broken by the lower index"""
    import gwdepth_b200  # noqa: F401
    from gwdepth_b200 import ops
    g = torch.Generator().manual_seed(Q + npts)
    logits = torch.randn(B, Q, 2, generator=g)
    logits[0, 5, 0] = logits[0, 9, 0] = logits[0].max() + 1.0           # an exact tie at the top
    lines = torch.rand(B, Q, 6, generator=g)
    ref_xy, ids = ops.select_lines(logits.cuda(), lines.cuda(), num_ref, npts)
    want = torch.sort(logits[:, :, 0], dim=-1, descending=True, stable=True).indices[:, :num_ref]
    assert torch.equal(ids.cpu(), want)
    pts = torch.gather(lines, 1, want[:, :, None].expand(-1, -1, 6)).reshape(B, num_ref, 3, 2)[:, :, :npts] * 2 - 1.0
    assert torch.equal(ref_xy.cpu(), pts.reshape(B, -1, 2))
