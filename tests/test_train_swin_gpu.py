"""GPU parity of the class-window Swin stage training path (train_swin.ClassStage) and of
the stage entries between the Swin stages (train_entry.StageEntry, A12 glue): the two attention-backward kernels against
torch.autograd on the same bf16 operands, then forward values and every input / parameter gradient of the modules against
torch.autograd over the CPU oracle's `swin_stage` / stage-entry expressions.

Tolerances: single kernels to bf16 output rounding (6e-3 relative L2; fp32 outputs 1e-4); modules a few per cent (bf16
activations through two blocks of ~25 kernels each)."""
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


# ------------------------------------------------------------------------------------------ class-window Swin stage (A13/A14/A16)
@pytest.mark.parametrize("heads,hd,N,items,nWm", [(16, 4, 49, 12, 4), (16, 8, 49, 6, 0), (8, 32, 49, 5, 0), (4, 16, 64, 3, 3)])
def test_window_attention_bwd(heads, hd, N, items, nWm):
    """gwd_window_attention_bwd == autograd of softmax(scale q k^T + bias + mask) v on fused q | k | v rows"""
    ops = _ops()
    g = _g(heads * hd + items)
    C = heads * hd
    scale = hd ** -0.5
    qkv = torch.randn(items * N, 3 * C, generator=g).bfloat16()
    d_o = torch.randn(items * N, C, generator=g).bfloat16()
    bias = torch.randn(heads, N, N, generator=g)
    mask = None
    if nWm:
        mask = torch.where(torch.rand(nWm, N, N, generator=g) < 0.3, torch.full((nWm, N, N), -100.0), torch.zeros(nWm, N, N))
        mask[:, torch.arange(N), torch.arange(N)] = 0.0
    qr = qkv.float().view(items, N, 3, heads, hd).permute(2, 0, 3, 1, 4).clone().requires_grad_(True)   # [3, items, heads, N, hd]
    br = bias.clone().requires_grad_(True)
    a = (qr[0] * scale) @ qr[1].transpose(-2, -1) + br[None]
    if mask is not None:
        a = a + mask.repeat(items // nWm, 1, 1)[:, None]
    out = torch.softmax(a, dim=-1) @ qr[2]
    out.backward(d_o.float().view(items, N, heads, hd).permute(0, 2, 1, 3))
    dbias = torch.full((heads, N, N), 0.5, device="cuda")
    dqkv = ops.window_attention_bwd(qkv.cuda(), d_o.cuda(), items=items, heads=heads, N=N, hd=hd, scale=scale, bias=bias.cuda(),
                                    mask=mask.cuda() if mask is not None else None, dbias=dbias)
    ref = qr.grad.permute(1, 3, 0, 2, 4).reshape(items * N, 3 * C)
    assert rel_l2(dqkv, ref) < 6e-3
    assert rel_l2(dbias - 0.5, br.grad) < 1e-4


@pytest.mark.parametrize("heads,td,tc,N,items", [(16, 4, 12, 49, 7), (16, 4, 24, 49, 3), (8, 8, 32, 49, 2)])
def test_token_attention_bwd(heads, td, tc, N, items):
    """gwd_token_attention_bwd == autograd of the class-token channel attention (multiscale_transformerr.py:561-578)"""
    ops = _ops()
    g = _g(heads + tc)
    TD, TC = heads * td, heads * tc
    scale = 0.37
    rows = items * N
    dq, sq = torch.randn(rows, TD, generator=g).bfloat16(), torch.randn(rows, TD, generator=g).bfloat16()
    gkv = torch.randn(rows, 2 * TC, generator=g).bfloat16()
    d_dout, d_sout = torch.randn(rows, TD, generator=g).bfloat16(), torch.randn(rows, TD, generator=g).bfloat16()
    leaf = lambda t: t.float().clone().requires_grad_(True)
    dqr, sqr, gr = leaf(dq), leaf(sq), leaf(gkv)
    tk = gr[:, :TC].reshape(items, N, heads, tc).permute(0, 2, 1, 3)
    tv = gr[:, TC:].reshape(items, N, heads, tc).permute(0, 2, 1, 3)

    def chan(tok):
        tq = tok.reshape(items, N, heads, td).permute(0, 2, 1, 3) * scale
        a = torch.softmax(tq.transpose(-2, -1) @ tk, dim=-1)
        return (a @ tv.transpose(-2, -1)).reshape(items, -1, N).permute(0, 2, 1).reshape(rows, TD)
    (chan(dqr) * d_dout.float()).sum().add((chan(sqr) * d_sout.float()).sum()).backward()
    # forward kernel agrees with the same formula
    dout, sout = torch.empty(rows, TD, dtype=torch.bfloat16, device="cuda"), torch.empty(rows, TD, dtype=torch.bfloat16, device="cuda")
    gc = gkv.cuda()
    ops.token_attention(dq.cuda(), sq.cuda(), gc, gc[:, TC:], dout, sout, items=items, N=N, heads=heads, td=td, tc=tc, q_rs=TD,
                        k_rs=2 * TC, v_rs=2 * TC, o_rs=TD, scale=scale)
    assert rel_l2(dout, chan(dqr).detach()) < 1e-2
    g_dq, g_sq, g_gkv = ops.token_attention_bwd(dq.cuda(), sq.cuda(), gc, d_dout.cuda(), d_sout.cuda(), items=items, N=N, heads=heads,
                                                td=td, tc=tc, scale=scale)
    assert rel_l2(g_dq, dqr.grad) < 6e-3 and rel_l2(g_sq, sqr.grad) < 6e-3
    assert rel_l2(g_gkv, gr.grad) < 6e-3


@pytest.mark.parametrize("stage,C,depth,B,H,W,tok_scale", [(3, 64, 1, 2, 10, 12, 1.0), (2, 128, 2, 1, 14, 9, 1.0), (1, 256, 2, 1, 7, 8, 1.0),
                                                          (3, 64, 1, 1, 32, 40, 1.0), (3, 64, 1, 1, 32, 40, 0.01), (3, 64, 1, 1, 32, 40, 100.0)])
def test_class_stage_gradients_match_oracle_autograd(stage, C, depth, B, H, W, tok_scale):
    """train_swin.ClassStage (forward + backward) against torch.autograd over the oracle's `swin_stage` with class tokens:
    the three output streams, the gradients of the three input streams and of every live parameter (window padding, the
    shifted block with its mask, relative-position bias tables, the shared proj_dth)"""
    _ops()
    from gwdepth_b200.train_swin import ClassStage
    prefix = "dense_encoder.class_transformer%d." % stage
    sd = {k: v.clone() for k, v in synth_weights().items() if k.startswith(prefix)}
    g = _g(stage * 11)
    td, L = 64, H * W
    x, d, s = (torch.randn(B * L, n, generator=g).bfloat16() for n in (C, td, td))
    gx, gd, gs = (torch.randn(B * L, n, generator=g).bfloat16() for n in (C, td, td))
    gd, gs = (gd.float() * tok_scale).bfloat16(), (gs.float() * tok_scale).bfloat16()      # cross paths must be right at any ratio
    xr, dr, sr = (t.float().view(B, L, -1).requires_grad_(True) for t in (x, d, s))
    sdr = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in sd.items()}
    xo, do, so = oracle.swin_stage(xr, H, W, oracle.P(sdr, prefix), depth, 16, 7, dtok=dr, stok=sr)
    ((xo * gx.float().view(B, L, -1)).sum() + (do * gd.float().view(B, L, -1)).sum() + (so * gs.float().view(B, L, -1)).sum()).backward()
    st = ClassStage({k: v.cuda() for k, v in sd.items()}, prefix, C, depth)
    back = st.state_dict()
    for k, v in back.items():
        assert torch.equal(v.cpu(), sd[k]), k
    x1, d1, s1 = st.forward(x.cuda(), d.cuda(), s.cuda(), B, H, W)
    assert rel_l2(x1, xo.detach().view(B * L, -1)) < 2e-2 and rel_l2(d1, do.detach().view(B * L, -1)) < 2e-2
    assert rel_l2(s1, so.detach().view(B * L, -1)) < 2e-2
    g_x, g_d, g_s = st.backward(gx.cuda(), gd.cuda(), gs.cuda())
    assert rel_l2(g_x, xr.grad.view(B * L, -1)) < 5e-2 and rel_l2(g_d, dr.grad.view(B * L, -1)) < 5e-2
    assert rel_l2(g_s, sr.grad.view(B * L, -1)) < 5e-2
    grads = st.grads()
    bad = {}
    for k, v in sdr.items():
        if not v.is_floating_point():
            continue
        if v.grad is None:
            assert k not in grads, k                   # dead parameters (proj_seg, diff_*, border_*) are not stored
            continue
        e = rel_l2(grads[k], v.grad)
        if e > 6e-2:
            bad[k] = round(e, 3)
    assert not bad, bad


# ------------------------------------------------------------------------------------------ stage entries (A12 glue)
@pytest.mark.parametrize("si,Cprev,C,Cb", [(3, 128, 64, 256), (2, 256, 128, 512)])
def test_stage_entry_gradients_match_oracle_autograd(si, Cprev, C, Cb):
    """train_entry.StageEntry against torch.autograd over the oracle's stage-entry expressions (dense_encoder lines for the
    1/8 and 1/4 stages): outputs, gradients of the previous features / tokens / backbone map and of every parameter"""
    _ops()
    from gwdepth_b200.train_entry import StageEntry
    B, h, w, td = 2, 6, 8, 64
    H, W = 2 * h, 2 * w
    sc = {2: "8", 3: "4"}[si]
    g = _g(si)
    prev_x = torch.randn(B, h, w, Cprev, generator=g).bfloat16()
    prev_d, prev_s = torch.randn(B * h * w, td, generator=g).bfloat16(), torch.randn(B * h * w, td, generator=g).bfloat16()
    feat = torch.randn(B, H, W, Cb, generator=g).bfloat16()
    gx, gd, gs = (torch.randn(B * H * W, n, generator=g).bfloat16() for n in (C, td, td))
    sd = {k: v.clone() for k, v in synth_weights().items() if k.startswith("dense_encoder.") and "transformer" not in k and "point_based" not in k}
    sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    p = oracle.P(sdr, "dense_encoder.")
    pxr, pdr, psr, fr = (t.float().clone().requires_grad_(True) for t in (prev_x, prev_d, prev_s, feat))
    up = F.interpolate(pxr.permute(0, 3, 1, 2), size=(H, W), mode="nearest")
    xo = oracle.linear(up.flatten(2).permute(0, 2, 1), p, "proj_class%d" % si) + \
        oracle.conv_a(fr.permute(0, 3, 1, 2), p, "proj_backbn%d" % si).flatten(2).permute(0, 2, 1)
    do = oracle.mlp_norm(oracle.up_tokens(pdr.view(B, h * w, td), h, w, (H, W)).flatten(2).permute(0, 2, 1), p, "old_depth_token_proj" + sc)
    so = oracle.mlp_norm(oracle.up_tokens(psr.view(B, h * w, td), h, w, (H, W)).flatten(2).permute(0, 2, 1), p, "old_seg_token_proj" + sc)
    ((xo.reshape(-1, C) * gx.float()).sum() + (do.reshape(-1, td) * gd.float()).sum() + (so.reshape(-1, td) * gs.float()).sum()).backward()
    ent = StageEntry({k: v.cuda() for k, v in sd.items()}, si)
    for k, v in ent.state_dict().items():
        assert torch.equal(v.cpu(), sd[k]), k
    x, d, s = ent.forward(prev_x.cuda(), prev_d.cuda(), prev_s.cuda(), feat.cuda())
    assert rel_l2(x, xo.detach().reshape(-1, C)) < 1e-2 and rel_l2(d, do.detach().reshape(-1, td)) < 1e-2
    assert rel_l2(s, so.detach().reshape(-1, td)) < 1e-2
    d_px, d_pd, d_ps, d_feat = ent.backward(gx.cuda(), gd.cuda(), gs.cuda(), need_dfeat=True)
    assert rel_l2(d_px, pxr.grad) < 2e-2 and rel_l2(d_pd, pdr.grad) < 3e-2 and rel_l2(d_ps, psr.grad) < 3e-2
    assert rel_l2(d_feat[..., :Cb], fr.grad) < 2e-2
    grads = ent.grads()
    bad = {k: rel_l2(grads[k], sdr[k].grad) for k in grads}
    bad = {k: e for k, e in bad.items() if e > 3e-2}
    assert not bad, bad


# ------------------------------------------------------------------------------------------ the dense branch behind the 1/32 stage
def _oracle_dense_branch(sdr, x32, feats, coords1, coords2, depth_gt, seg_gt, size):
    """the body of oracle.dense_encoder from the 1/16 stage on (sample points pinned) + dense head + the five losses"""
    cfg = oracle.DEFAULT_CFG
    p = oracle.P(sdr, "dense_encoder.")
    heads, ws, D = cfg["dense_trans_heads"], cfg["window"], cfg["dense_trans_dim"]
    B = x32.shape[0]
    tokens = lambda m: m.flatten(2).permute(0, 2, 1)
    depths, prev, dtok, stok = [], x32, None, None
    for si, (feat, pts) in enumerate(zip(feats, (None, coords1, coords2)), start=1):
        Hs, Ws = feat.shape[-2:]
        hp, wp = prev.shape[-2:]
        up = F.interpolate(prev, size=(Hs, Ws), mode="nearest")
        x = oracle.linear(tokens(up), p, "proj_class%d" % si) + tokens(oracle.conv_a(feat, p, "proj_backbn%d" % si))
        if si == 1:
            dtok, stok = p["depth_token"].expand(B, Hs * Ws, -1), p["seg_token"].expand(B, Hs * Ws, -1)
        else:
            sc = {2: "8", 3: "4"}[si]
            dtok = oracle.mlp_norm(tokens(oracle.up_tokens(dtok, hp, wp, (Hs, Ws))), p, "old_depth_token_proj" + sc)
            stok = oracle.mlp_norm(tokens(oracle.up_tokens(stok, hp, wp, (Hs, Ws))), p, "old_seg_token_proj" + sc)
        x, dtok, stok = oracle.swin_stage(x, Hs, Ws, p.sub("class_transformer%d" % si), cfg["class_trans_layers"][si - 1], heads, ws,
                                          dtok=dtok, stok=stok)
        if si == 1:
            depth = oracle.depth_head(torch.cat([x, dtok], dim=-1), p, "depth_pred16").permute(0, 2, 1).reshape(B, -1, Hs, Ws)
        else:
            pos = oracle.sine_position(torch.zeros(B, Hs, Ws, dtype=torch.bool), D // (8 if si == 2 else 16), False)
            depth = oracle.point_based_pred(x, dtok, depths[-1], pts.view(B, -1, 1, 2), Hs, Ws, pos, p.sub("point_based_pred%d" % (si - 1)),
                                            D // (4 if si == 2 else 8))
        depths.append(depth)
        prev = x.permute(0, 2, 1).reshape(B, -1, Hs, Ws)
    img = lambda t: t.permute(0, 2, 1).reshape(B, -1, *feats[2].shape[-2:])
    depth, seg = oracle.dense_head(prev, depths[-1], img(dtok), img(stok), size, oracle.P(sdr, "depth_decoder."), 10.0)
    losses = oracle.depth_losses(depths + [depth], depth_gt) + [oracle.seg_loss(seg, seg_gt)]
    return depths + [depth], seg, losses


def _branch_case():
    B, h5, w5 = 1, 4, 5
    g = _g(77)
    x32 = torch.randn(B, h5, w5, 512, generator=g).bfloat16()
    feats = [torch.randn(B, (2 ** k) * h5, (2 ** k) * w5, c, generator=g).bfloat16() for k, c in ((1, 1024), (2, 512), (3, 256))]
    coords1, coords2 = torch.rand(B, 30, 2, generator=g) * 2 - 1, torch.rand(B, 80, 2, generator=g) * 2 - 1
    depth_gt = torch.rand(B, 1, 32 * h5, 32 * w5, generator=g) * 10.5 + 0.1
    seg_gt = (torch.rand(B, 1, 32 * h5, 32 * w5, generator=g) > 0.5).long()
    return x32, feats, coords1, coords2, depth_gt, seg_gt


def _oracle_branch_grads(sd, emulate_bf16):
    """losses, depth maps and all gradients of the oracle chain; emulate_bf16: weights and the outputs of every linear / conv /
    LayerNorm / activation rounded to bf16 (straight-through), i.e. the oracle run at the CUDA path's storage precision"""
    x32, feats, coords1, coords2, depth_gt, seg_gt = _branch_case()
    H, W = depth_gt.shape[-2:]
    rnd = lambda t: t + (t.bfloat16().float() - t).detach()  # noqa: E731
    sdr = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in sd.items()}
    use = {k: (rnd(v) if (emulate_bf16 and v.is_floating_point() and v.dim() > 1) else v) for k, v in sdr.items()}
    nchw = lambda t: t.float().permute(0, 3, 1, 2).clone().requires_grad_(True)  # noqa: E731
    x32r, featr = nchw(x32), [nchw(f) for f in feats]
    names = ["linear", "conv2d", "layer_norm", "gelu", "elu"]
    orig = {n: getattr(F, n) for n in names}
    try:
        if emulate_bf16:
            for n in names:
                setattr(F, n, (lambda f: (lambda *a, **k: rnd(f(*a, **k))))(orig[n]))
        depths, seg, losses = _oracle_dense_branch(use, x32r, featr, coords1, coords2, depth_gt, seg_gt, (H, W))
        sum(losses).backward()
    finally:
        for n in names:
            setattr(F, n, orig[n])
    grads = {k: v.grad for k, v in sdr.items() if v.is_floating_point() and v.grad is not None}
    grads.update({"d_x32": x32r.grad.permute(0, 2, 3, 1), "d_c4": featr[0].grad.permute(0, 2, 3, 1), "d_c3": featr[1].grad.permute(0, 2, 3, 1)})
    return [d.detach() for d in depths], [float(l.detach()) for l in losses], grads


def _module_of(k):
    return ".".join(k.split(".")[:4] if k.startswith("dense_encoder.class_transformer") else k.split(".")[:2])


def test_dense_branch_gradients_match_oracle_autograd():
    """train_branch.DenseBranch (three class-window stages + entries + coarse depth head + both point predictions + dense head +
    five losses) against torch.autograd over the oracle's functions chained as `dense_encoder` chains them.  Forward values and
    losses must agree closely.  The gradients of this network are sensitive to bf16 storage (peaked channel soft-maxes in the
    token attention): the ORACLE ITSELF, run with weights and layer outputs rounded to bf16, moves its Swin-stage gradients by
    20-50 % -- so the bar for the CUDA path is the distance of that emulated run (per module: <= 1.5 x emulated + 3 %),
    while modules whose gradients are insensitive (heads, point predictions, pyramids) must match to a few per cent."""
    _ops()
    from gwdepth_b200.train_branch import DenseBranch
    live = ("dense_encoder.class_transformer", "dense_encoder.point_based_pred", "dense_encoder.proj_", "dense_encoder.old_",
            "dense_encoder.depth_pred16", "dense_encoder.depth_token", "dense_encoder.seg_token", "depth_decoder.")
    sd = {k: v.clone() for k, v in synth_weights().items() if k.startswith(live)}
    depths_o, losses_o, g_ref = _oracle_branch_grads(sd, False)
    _, _, g_emu = _oracle_branch_grads(sd, True)
    x32, feats, coords1, coords2, depth_gt, seg_gt = _branch_case()
    B, h5, w5 = x32.shape[:3]
    br = DenseBranch({k: v.cuda() for k, v in sd.items()})
    outs, losses, d_x32, d_c4, d_c3 = br.loss_and_grads(x32.cuda(), torch.zeros(B, h5, w5, device="cuda"), [f.cuda() for f in feats],
                                                        depth_gt.cuda(), seg_gt.cuda(),
                                                        pinned={"sample1": coords1.cuda(), "sample2": coords2.cuda()})
    for i, (d_, o) in enumerate(zip(outs["pred_depth"], depths_o)):
        assert rel_l2(d_.reshape(-1), o.reshape(-1)) < 2e-2, i
    for a, b in zip(losses.tolist(), losses_o):
        assert abs(a - b) < 5e-3 * abs(b), (losses.tolist(), losses_o)
    got = dict(br.grads(), d_x32=d_x32, d_c4=d_c4[..., :1024], d_c3=d_c3[..., :512])
    missing = [k for k in g_ref if k not in got]
    assert not missing, missing
    e_cuda, e_emu = {}, {}
    for k, ref in g_ref.items():
        m = _module_of(k)
        e_cuda[m] = max(e_cuda.get(m, 0.0), rel_l2(got[k], ref))
        e_emu[m] = max(e_emu.get(m, 0.0), rel_l2(g_emu[k], ref))
    print("dense branch gradient distance to the fp32 oracle, per module (CUDA / bf16-emulated oracle):",
          {m: (round(e_cuda[m], 3), round(e_emu[m], 3)) for m in e_cuda})
    bad = {m: (e_cuda[m], e_emu[m]) for m in e_cuda if e_cuda[m] > 1.5 * e_emu[m] + 3e-2}
    assert not bad, bad
    for m in ("dense_encoder.depth_pred16", "dense_encoder.point_based_pred1", "dense_encoder.point_based_pred2"):
        assert e_cuda[m] < 7e-2, (m, e_cuda[m])
    assert all(e < 8e-2 for m, e in e_cuda.items() if m.startswith("depth_decoder.")), e_cuda


def test_dense_branch_forward_equals_the_inference_engine_on_real_activations():
    """train_branch.DenseBranch fed with the inference engine's own x32 / depth_pred32 / backbone maps of a synthetic image batch
    (sample points pinned to the engine's) reproduces the engine's four depth maps and seg logits: the training forward (un-folded
    weights, soft-max scale as a kernel argument, separate token streams) and the parity-tested inference forward are the same
    function on real data; then one optimizer step lowers the summed dense loss on that batch."""
    _ops()
    from helpers import synth
    from gwdepth_b200 import model as M
    from gwdepth_b200.train_branch import DenseBranch
    net, _, _ = M.build_model(M.default_args(device="cuda"))
    sd = synth_weights()
    net.load_state_dict(sd)
    net.cuda().eval()
    B, H, W = 2, 224, 320
    images, _, depth_gt, seg_gt = synth.synth_batch(B, H, W, seed=1)
    eng = net.plan()
    trace = {}
    with torch.no_grad():
        out = eng.forward(images.cuda().float().contiguous(), trace=trace)
        feats = eng.backbone(images.cuda().float().contiguous())
    h5, w5 = feats[3].shape[1:3]
    br = DenseBranch({k: v.cuda() for k, v in sd.items()}, lr=1e-4, max_norm=0.1)
    args = (trace["x32"].view(B, h5, w5, -1).contiguous(), trace["depth0"].contiguous(), [feats[2], feats[1], feats[0]],
            depth_gt.cuda().float(), seg_gt.cuda().long().view(B, 1, H, W))
    pinned = {"sample1": trace["sample1"].cuda(), "sample2": trace["sample2"].cuda()}
    outs, losses, d_x32, d_c4, d_c3 = br.loss_and_grads(*args, pinned=pinned)
    for i, (a, b) in enumerate(zip(outs["pred_depth"], out["pred_depth"])):
        assert rel_l2(a.reshape(-1), b.reshape(-1)) < 2e-2, i
    assert rel_l2(outs["pred_seg"], out["pred_seg"]) < 3e-2
    assert all(torch.isfinite(t.float()).all() for t in (d_x32, d_c4, d_c3, losses))
    first = float(losses.sum())
    for _ in range(6):
        last = float(br.train_step(*args, pinned=pinned).sum())
    assert last < first, (first, last)
