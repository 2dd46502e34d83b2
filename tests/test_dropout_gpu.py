"""GPU: train-mode dropout of the DETR layers (src/args.py:51 `--dropout 0.1`; src/models/transformer.py:149-162,212-233;
src/models/multi_head_attention.py:368).  Masks come from another generator than torch's, so parity is checked through the
MASK ITSELF: the element-wise kernel's mask is read off its output, the attention kernels' mask is recovered with one-hot value
matrices; forward and backward are then compared with torch.autograd using exactly that mask."""
import pytest
import torch

from helpers import synth, synth_weights

pytestmark = pytest.mark.gpu


def _ops():
    import gwdepth_b200  # noqa: F401
    from gwdepth_b200 import ops
    return ops


def rel_l2(got, ref):
    got, ref = got.double().cpu(), ref.double().cpu()
    return float((got - ref).norm() / ref.norm().clamp_min(1e-30))


def test_elementwise_dropout_is_its_own_backward():
    ops = _ops()
    g = torch.Generator().manual_seed(0)
    p = 0.1
    x = (torch.randn(4096, 256, generator=g).abs() + 0.5).bfloat16().cuda()
    res = torch.randn(4096, 256, generator=g).bfloat16().cuda()
    seed = torch.tensor([12345], dtype=torch.int32, device="cuda")
    y = ops.dropout(x, seed, 7, p)
    keep = y != 0
    frac = float(keep.float().mean())
    assert abs(frac - (1 - p)) < 3e-3, frac                                   # 1 M samples: sigma = 3e-4
    assert rel_l2(y[keep], x[keep].float() / (1 - p)) < 4e-3                   # kept elements carry 1 / (1 - p)
    assert torch.equal(ops.dropout(x, seed, 7, p), y)                         # deterministic in (seed, site)
    assert not torch.equal(ops.dropout(x, seed, 8, p) != 0, keep)             # another site, another mask
    seed2 = seed + 1
    assert not torch.equal(ops.dropout(x, seed2, 7, p) != 0, keep)            # another step, another mask
    z = ops.dropout(x, seed, 7, p, res=res)
    assert rel_l2(z, res.float() + y.float()) < 4e-3
    dz = torch.randn(4096, 256, generator=g).bfloat16().cuda()
    dx = ops.dropout(dz, seed, 7, p)                                          # the backward: the same mask on the gradient
    assert torch.equal(dx != 0, keep & (dz != 0))
    # in place
    h = x.clone()
    ops.dropout(h, seed, 7, p, out=h)
    assert torch.equal(h, y)
    # rows and columns are decorrelated (no stripe patterns): per-row and per-column keep rates
    assert float(keep.float().mean(0).std()) < 0.01 and float(keep.float().mean(1).std()) < 0.03


@pytest.mark.parametrize("B,Lq,Lk", [(2, 100, 96), (1, 300, 320), (2, 128, 64)])
def test_attention_dropout_forward_and_backward(B, Lq, Lk):
    """the tcgen05 attention kernels (single pass and key-tiled) with dropout on the probabilities, and gwd_attention_bwd
    regenerating the same mask"""
    ops = _ops()
    heads, hd, p = 8, 32, 0.1
    E = heads * hd
    g = torch.Generator().manual_seed(Lq + Lk)
    q, k, v = (torch.randn(B * L, E, generator=g).bfloat16().cuda() for L in (Lq, Lk, Lk))
    q = (q.float() * hd ** -0.5).bfloat16()
    seed = torch.tensor([99], dtype=torch.int32, device="cuda")
    site = 31

    def run(vv, drop=True):
        o = torch.empty(B * Lq, E, dtype=torch.bfloat16, device="cuda")
        ops.attention(q, k, vv, o, items=B, heads=heads, Lq=Lq, Lk=Lk, hd=hd, q_strides=(Lq * E, E), k_strides=(Lk * E, E),
                      v_strides=(Lk * E, E), o_strides=(Lq * E, E), scale=1.0, dropout=(seed, site, p) if drop else None)
        return o
    # recover the dropped probabilities with one-hot values: V_j[key, d] = 1 iff key == 32 j + d  ->  O_j = P_drop[:, 32 j : 32 j + 32]
    pd = torch.zeros(B, heads, Lq, Lk)
    for j in range((Lk + 31) // 32):
        vj = torch.zeros(B, Lk, heads, hd)
        n = min(32, Lk - 32 * j)
        vj[:, 32 * j + torch.arange(n), :, torch.arange(n)] = 1.0
        oj = run(vj.reshape(B * Lk, E).bfloat16().cuda()).float().cpu().view(B, Lq, heads, hd).permute(0, 2, 1, 3)
        pd[..., 32 * j: 32 * j + n] = oj[..., :n]
    h4 = lambda t, L: t.float().cpu().view(B, L, heads, hd).permute(0, 2, 1, 3)  # noqa: E731
    qr, kr, vr = (h4(t, L).clone().requires_grad_(True) for t, L in ((q, Lq), (k, Lk), (v, Lk)))
    P = torch.softmax(qr @ kr.transpose(-1, -2), -1)
    mask = (pd > 0.5 * P.detach() / (1 - p) - 1e-6) & (P.detach() > 1e-4) | ((pd > 0) & (P.detach() <= 1e-4))
    sure = P.detach() > 1e-3                                   # probabilities large enough to read the mask off bf16 outputs
    frac = float(mask[sure].float().mean())
    assert abs(frac - (1 - p)) < 0.02, frac
    assert rel_l2(pd[sure & mask], (P.detach() / (1 - p))[sure & mask]) < 1e-2
    mfull = torch.where(sure, mask, pd > 0).float()            # tiny probabilities: whatever the kernel kept (bf16 flushes some)
    out_ref = ((P * mfull / (1 - p)) @ vr).permute(0, 2, 1, 3).reshape(B * Lq, E)
    o = run(v)
    assert rel_l2(o, out_ref.detach()) < 1e-2
    d_o = torch.randn(B * Lq, E, generator=g).bfloat16()
    out_ref.backward(d_o.float())
    dq, dk, dv = (torch.empty(B * L, E, dtype=torch.bfloat16, device="cuda") for L in (Lq, Lk, Lk))
    ops.attention_bwd(q, k, v, d_o.cuda(), dq, dk, dv, items=B, heads=heads, Lq=Lq, Lk=Lk, hd=hd, q_strides=(Lq * E, E),
                      k_strides=(Lk * E, E), v_strides=(Lk * E, E), do_strides=(Lq * E, E), dq_strides=(Lq * E, E),
                      dk_strides=(Lk * E, E), dv_strides=(Lk * E, E), scale=1.0, o=o, o_strides=(Lq * E, E), dropout=(seed, site, p))
    back = lambda gr, L: gr.permute(0, 2, 1, 3).reshape(B * L, E)  # noqa: E731
    assert rel_l2(dv, back(vr.grad, Lk)) < 1.5e-2
    assert rel_l2(dq, back(qr.grad, Lq)) < 2e-2 and rel_l2(dk, back(kr.grad, Lk)) < 2e-2
    # and without dropout the kernels are what they were
    assert rel_l2(run(v, drop=False), ((P.detach()) @ h4(v, Lk)).permute(0, 2, 1, 3).reshape(B * Lq, E)) < 1e-2


def test_training_step_with_reference_default_dropout():
    """the reference's default training configuration (--dropout 0.1) through the fused step: finite losses, new masks every
    step (also under CUDA-graph replay), every trained tensor receives a gradient, and p = 0 is bit-for-bit the no-dropout path"""
    import gwdepth_b200  # noqa: F401
    from gwdepth_b200 import model as M
    B, H, W = 2, 128, 160
    images, targets, depth_gt, seg_gt = synth.synth_batch(B, H, W, seed=0)
    tg = [{k: v.cuda() for k, v in t.items()} for t in targets]
    args = M.default_args(device="cuda")                       # dropout = 0.1 as src/args.py:51
    assert args.dropout == 0.1
    net, crit, _ = M.build_model(args)
    net.load_state_dict(synth_weights())
    net.cuda().train()
    tr = net.trainer()
    assert tr.line.p_drop == 0.1 and tr.use_cuda_graph
    x, dg, sg = images.cuda(), depth_gt.cuda(), seg_gt.cuda()
    totals, logits = [], []
    for _ in range(3):
        total, losses = tr.train_step(x, tg, dg, sg, crit[0].cuda())
        totals.append(float(total))
        logits.append(tr.last["logits"].clone())
        assert all(torch.isfinite(v).all() for v in losses.values())
    grads = tr.grads()
    assert len(grads) == 684 and all(torch.isfinite(g).all() for g in grads.values())
    assert not torch.equal(logits[0], logits[1]) and not torch.equal(logits[1], logits[2])
    # a second engine with lr = 0: two steps on the same batch with the same weights differ only through the masks
    net2, crit2, _ = M.build_model(args)
    net2.load_state_dict(synth_weights())
    net2.cuda().train()
    tr2 = net2.trainer(lr=0.0, lr_backbone=0.0, weight_decay=0.0)
    a, _ = tr2.train_step(x, tg, dg, sg, crit2[0].cuda())
    la = tr2.last["logits"].clone()
    b, _ = tr2.train_step(x, tg, dg, sg, crit2[0].cuda())
    assert not torch.equal(la, tr2.last["logits"]), "the replayed graph re-used the dropout masks of the previous step"
    assert abs(float(a) - float(b)) < 0.2 * abs(float(a))
    # eval mode ignores dropout: same forward as a model built with --dropout 0.0
    net.sync_from_trainer()
    net.eval()
    net0, _, _ = M.build_model(M.default_args(device="cuda", dropout=0.0))
    net0.load_state_dict(net.state_dict())
    net0.cuda().eval()
    with torch.no_grad():
        assert torch.equal(net(x)["pred_depth"][3], net0(x)["pred_depth"][3])
