"""GPU parity of the WHOLE-MODEL training step (train_model.Trainer behind GlassRGBD, SURVEY 8b / VERDICT r1 items 1-3):
  * gradients of all 684 trained tensors (738 trainable minus the 54 the reference never reaches) against
    torch.autograd over the fp32 CPU oracle on the same batch, selections and matchings pinned to the oracle's;
  * the drop-in surface: the reference's own `train_one_epoch` (src/engine_glassrgbd.py:22-171, imported unmodified from
    baseline/_ref when it is staged; a line-by-line mirror of its loop otherwise) drives `model(samples)` ->
    criteria -> `losses.backward()` -> clip_grad_norm_ -> torch AdamW for two steps;
  * the fused step (Trainer.train_step) equals the drop-in step's losses and gradients.

Gradient tolerances: this synthetic network is sensitive to bf16 storage itself -- the ORACLE's own gradients move by 5-50 % per
module when its weights, stored activations and activation gradients are rounded to bf16, see DESIGN.md section 4 -- so every
module is held to 1.5 x that distance + 3 % (the modules next to the losses, where no amplification happens, to 8 % absolute)."""
import collections
import math
import os
import sys

import pytest
import torch
import torch.nn.functional as F

from helpers import oracle, synth, synth_weights

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _mods():
    import gwdepth_b200  # noqa: F401
    from gwdepth_b200 import model as M
    from gwdepth_b200.train_model import Trainer
    return M, Trainer


def rel_l2(got, ref):
    got, ref = got.double().cpu(), ref.double().cpu()
    return float((got - ref).norm() / ref.norm().clamp_min(1e-30))


def _group(k):
    return ".".join(k.split(".")[:3]) if k.startswith(("backbone", "transformer")) else ".".join(k.split(".")[:2])


def _trainable(sd):
    """the reference's trainable set: everything but the stem, layer1 and the FrozenBatchNorm buffers (backbone.py:62-64)"""
    frozen = ("backbone.0.body.conv1", "backbone.0.body.layer1")
    return {k for k, v in sd.items() if v.is_floating_point() and "running" not in k and ".bn" not in k and "downsample.1" not in k
            and not k.startswith(frozen)}


class _Bf16Store(torch.autograd.Function):
    """what storing a tensor in bf16 does to a training step: the value is rounded on the way forward, its gradient on the way back"""

    @staticmethod
    def forward(ctx, x):
        return x.bfloat16().float()

    @staticmethod
    def backward(ctx, g):
        return g.bfloat16().float()


def _oracle_grads(sd, images, targets, depth_gt, seg_gt, wd, pin=None, emulate=False, mask=None):
    """total loss of the engine's loop and its gradients by torch.autograd over the oracle; emulate=True stores the weights and
    every activation (and activation gradient) in bf16"""
    # what the CUDA path stores in bf16: the outputs of every Linear / convolution / LayerNorm / activation, attention
    # probabilities and attention outputs (torch.softmax / bmm / @ in the oracle), and the weights themselves
    patches = [(F, n) for n in ("linear", "conv2d", "layer_norm", "gelu", "relu", "elu")] + [(torch, "softmax"), (torch, "bmm"),
                                                                                           (torch.Tensor, "__matmul__")]
    orig = [(o, n, getattr(o, n)) for o, n in patches]
    st = _Bf16Store.apply
    train = _trainable(sd)
    leaves = {k: (v.clone().requires_grad_(True) if k in train else v) for k, v in sd.items()}
    use = {k: (st(v) if (emulate and k in train and v.dim() > 1) else v) for k, v in leaves.items()}
    try:
        if emulate:
            for o, n, f in orig:
                setattr(o, n, (lambda f: (lambda *a, **k: st(f(*a, **k))))(f))
        trace = {}
        out = oracle.forward(use, images, mask=mask, pinned=pin, trace=trace, grad=True)
        tl = [t["lines"] for t in targets]
        set_l, idx = oracle.set_criterion(out, tl)
        dl = oracle.depth_losses(out["pred_depth"], depth_gt)
        total = sum(v * wd[k] for k, v in set_l.items()) + sum(dl) + oracle.seg_loss(out["pred_seg"], seg_gt)
        total.backward()
    finally:
        for o, n, f in orig:
            setattr(o, n, f)
    return float(total), {k: v.grad for k, v in leaves.items() if k in train}, trace, idx


def test_whole_model_gradients_match_oracle_autograd():
    M, Trainer = _mods()
    B, H, W = 2, 128, 160
    images, targets, depth_gt, seg_gt = synth.synth_batch(B, H, W, seed=0)
    sd = synth_weights()
    _, crit, _ = M.build_model(M.default_args(device="cuda", dropout=0.0))
    wd = crit[0].weight_dict
    total, ref, trace, idx = _oracle_grads(sd, images, targets, depth_gt, seg_gt, wd)
    pin_o = {"line_ids": trace["line_ids"], "sample1": (trace["sample1"], trace["sample1_idx"]), "sample2": (trace["sample2"], trace["sample2_idx"])}
    _, emu, _, _ = _oracle_grads(sd, images, targets, depth_gt, seg_gt, wd, pin=pin_o, emulate=True)
    assert sum(g is not None for g in ref.values()) == 684 and len(ref) == 738       # 54 never-used tensors

    tr = Trainer(sd)
    pinned = {"line_ids": trace["line_ids"].cuda(), "sample1": trace["sample1"].cuda(), "sample2": trace["sample2"].cuda()}
    tg = [{k: v.cuda() for k, v in t.items()} for t in targets]
    logits, lines, outs = tr.forward(images.cuda(), pinned)
    g = tr.dense.loss_grads(outs, depth_gt.cuda(), seg_gt.cuda())
    tr.backward_dense(*g)
    stacked = idx[1:] + idx[:1]                    # the oracle lists the final stage first, the stacked layout has it last
    _, dlogits, dlines = crit[0].cuda().forward_backward_stacked(logits, lines, tg, pinned_pairs=stacked)
    tr.backward_line(dlogits, dlines)
    got_total = float(crit[0].last_total + tr.dense.losses().sum())
    assert abs(got_total - total) < 5e-3 * abs(total), (got_total, total)
    grads = tr.grads()
    live = {k for k, v in ref.items() if v is not None}
    assert set(grads) == live, (sorted(live - set(grads))[:5], sorted(set(grads) - live)[:5])
    agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0])
    for k in live:
        a = agg[_group(k)]
        a[0] += float((grads[k].double().cpu() - ref[k].double()).pow(2).sum())
        a[1] += float((emu[k].double() - ref[k].double()).pow(2).sum())
        a[2] += float(ref[k].double().pow(2).sum())
        assert grads[k].shape == ref[k].shape and torch.isfinite(grads[k]).all(), k
    bad = {}
    near_loss = ("depth_decoder.", "dense_encoder.point_based_pred", "dense_encoder.depth_pred16", "class_embed.")
    for grp, (e, y, n) in agg.items():
        e, y = math.sqrt(e / n), math.sqrt(y / n)
        bar = 0.08 if grp.startswith(near_loss) else 1.5 * y + 0.03
        if e > bar:
            bad[grp] = (round(e, 3), round(y, 3))
    assert not bad, bad


def test_ragged_batch_gradients_match_oracle_autograd():
    """training on a PADDED batch (two images of different sizes, zero padding + mask as nested_tensor_from_tensor_list builds them;
    per-image position codes at every level, key-padding masks in the encoder self-attention and the decoder cross-attention,
    forward and backward): total loss and the gradients of all 684 trained tensors against torch.autograd over the oracle with
    the same mask, same yard-stick as the equal-size test"""
    M, Trainer = _mods()
    B, H, W = 2, 128, 160
    images, targets, depth_gt, seg_gt = synth.synth_batch(B, H, W, seed=5)
    mask = torch.zeros(B, H, W, dtype=torch.bool)
    mask[1, 96:, :] = True
    mask[1, :, 128:] = True
    images = images.masked_fill(mask[:, None], 0.0)
    depth_gt = depth_gt.masked_fill(mask[:, None], 0.0)
    seg_gt = seg_gt.masked_fill(mask[:, None], 0)
    sd = synth_weights()
    _, crit, _ = M.build_model(M.default_args(device="cuda", dropout=0.0))
    wd = crit[0].weight_dict
    total, ref, trace, idx = _oracle_grads(sd, images, targets, depth_gt, seg_gt, wd, mask=mask)
    pin_o = {"line_ids": trace["line_ids"], "sample1": (trace["sample1"], trace["sample1_idx"]), "sample2": (trace["sample2"], trace["sample2_idx"])}
    _, emu, _, _ = _oracle_grads(sd, images, targets, depth_gt, seg_gt, wd, pin=pin_o, emulate=True, mask=mask)
    tr = Trainer(sd)
    pinned = {"line_ids": trace["line_ids"].cuda(), "sample1": trace["sample1"].cuda(), "sample2": trace["sample2"].cuda()}
    tg = [{k: v.cuda() for k, v in t.items()} for t in targets]
    logits, lines, outs = tr.forward(images.cuda(), pinned, mask=mask.cuda())
    g = tr.dense.loss_grads(outs, depth_gt.cuda(), seg_gt.cuda())
    tr.backward_dense(*g)
    _, dlogits, dlines = crit[0].cuda().forward_backward_stacked(logits, lines, tg, pinned_pairs=idx[1:] + idx[:1])
    tr.backward_line(dlogits, dlines)
    got_total = float(crit[0].last_total + tr.dense.losses().sum())
    assert abs(got_total - total) < 5e-3 * abs(total), (got_total, total)
    grads = tr.grads()
    live = {k for k, v in ref.items() if v is not None}
    assert set(grads) == live
    agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0])
    for k in live:
        a = agg[_group(k)]
        a[0] += float((grads[k].double().cpu() - ref[k].double()).pow(2).sum())
        a[1] += float((emu[k].double() - ref[k].double()).pow(2).sum())
        a[2] += float(ref[k].double().pow(2).sum())
        assert torch.isfinite(grads[k]).all(), k
    bad = {}
    near_loss = ("depth_decoder.", "dense_encoder.point_based_pred", "dense_encoder.depth_pred16", "class_embed.")
    for grp, (e, y, n) in agg.items():
        e, y = math.sqrt(e / n), math.sqrt(y / n)
        bar = max(0.08 if grp.startswith(near_loss) else 0.0, 1.5 * y + 0.03)      # (here the oracle's own bf16 distance of point_based_pred2 is 10 %)
        if e > bar:
            bad[grp] = (round(e, 3), round(y, 3))
    assert not bad, bad
    # the un-masked run on the same pixels differs: the mask really entered (position codes + attention)
    logits2, _, _ = tr.forward(images.cuda(), pinned)
    assert float((logits2 - logits).abs().max()) > 1e-3


def test_drop_in_forward_backward_on_a_ragged_batch():
    """model([img_a, img_b]) under train(): the NestedTensor built with size_divisibility=32 goes through the autograd edge with its
    mask; the criterion's backward fills .grad of the 684 live parameters; without the rounding the call explains itself"""
    M, _ = _mods()
    model, crit, _ = M.build_model(M.default_args(device="cuda", dropout=0.0))
    model.load_state_dict(synth_weights())
    model.cuda().train()
    a = synth.synth_batch(1, 128, 160, seed=1)[0][0]
    b = synth.synth_batch(1, 96, 120, seed=2)[0][0]
    nt = M.nested_tensor_from_tensor_list([a.cuda(), b.cuda()], size_divisibility=32)
    assert tuple(nt.tensors.shape) == (2, 3, 128, 160) and nt.padded and bool(nt.mask[1, 96:].all()) and not bool(nt.mask[0].any())
    out = model(nt)
    loss = out["pred_logits"].float().square().mean() + out["pred_depth"][-1].mean() + out["pred_seg"].float().mean()
    loss.backward()
    got = [n for n, p_ in model.named_parameters() if p_.grad is not None]
    assert len(got) == 684 and all(torch.isfinite(p_.grad).all() for p_ in model.parameters() if p_.grad is not None)
    c = synth.synth_batch(1, 100, 120, seed=2)[0][0]
    with pytest.raises(NotImplementedError):
        model(M.nested_tensor_from_tensor_list([a.cuda()[:, :100, :150], c.cuda()]))      # 100 x 150: not a multiple of 32


def test_training_step_beyond_512_tokens():
    """a training step on 544 x 1024 images (17 x 32 = 544 tokens at 1/32: the long-sequence attention backward, the key-tiled forward
    kernel): total loss equals the oracle's, every trained tensor gets a finite, non-zero gradient"""
    M, Trainer = _mods()
    B, H, W = 1, 544, 1024
    images, targets, depth_gt, seg_gt = synth.synth_batch(B, H, W, seed=3)
    sd = synth_weights()
    _, crit, _ = M.build_model(M.default_args(device="cuda", dropout=0.0))
    wd = crit[0].weight_dict
    trace = {}
    with torch.no_grad():
        out = oracle.forward(sd, images, trace=trace)
        set_l, idx = oracle.set_criterion(out, [t["lines"] for t in targets])
        total = float(sum(v * wd[k] for k, v in set_l.items()) + sum(oracle.depth_losses(out["pred_depth"], depth_gt)) +
                      oracle.seg_loss(out["pred_seg"], seg_gt))
    tr = Trainer(sd)
    pinned = {"line_ids": trace["line_ids"].cuda(), "sample1": trace["sample1"].cuda(), "sample2": trace["sample2"].cuda()}
    tg = [{k: v.cuda() for k, v in t.items()} for t in targets]
    logits, lines, outs = tr.forward(images.cuda(), pinned)
    g = tr.dense.loss_grads(outs, depth_gt.cuda(), seg_gt.cuda())
    tr.backward_dense(*g)
    _, dlogits, dlines = crit[0].cuda().forward_backward_stacked(logits, lines, tg, pinned_pairs=idx[1:] + idx[:1])
    tr.backward_line(dlogits, dlines)
    got_total = float(crit[0].last_total + tr.dense.losses().sum())
    assert abs(got_total - total) < 1e-2 * abs(total), (got_total, total)
    grads = tr.grads()
    assert len(grads) == 684
    for k, v in grads.items():
        assert torch.isfinite(v).all(), k
    enc = [v for k, v in grads.items() if k.startswith("transformer.encoder.layers.0.self_attn")]
    assert enc and all(float(v.abs().max()) > 0 for v in enc)


def _loader(B, H, W, steps, NestedTensor):
    for s in range(steps):
        images, targets, depth_gt, seg_gt = synth.synth_batch(B, H, W, seed=s)
        mask = torch.zeros(B, H, W, dtype=torch.bool)
        yield (NestedTensor(images, mask), NestedTensor(depth_gt, mask), NestedTensor(seg_gt, mask), targets, ["synthetic_%d" % s] * B)


def _mirror_train_loop(model, criterions, loader, optimizer, device, max_norm, args):
    """src/engine_glassrgbd.py:22-171 without the logging (used only when the reference tree is not staged under baseline/_ref)"""
    model.train()
    criterion, criterion_depth, criterion_seg, _ = criterions
    log = []
    for samples, depth_gt, seg_gt, targets, _ in loader:
        samples, depth_gt, seg_gt = samples.to(device), depth_gt.to(device), seg_gt.to(device)
        targets = [{k: v.to(device) for k, v in t.items()} for t in targets]
        outputs = model(samples, reflc_mat=None, img_name="x")
        loss_dict = criterion(outputs, targets, depth_gt=depth_gt.tensors)
        mask = (depth_gt.tensors >= 0.2) & (depth_gt.tensors < 10.0)
        loss_depth = 0.0
        for i, pd in enumerate(outputs["pred_depth"]):
            size = pd.shape[-2:]
            d_gt = F.interpolate(depth_gt.tensors, size=size, mode="nearest")
            m_rs = F.interpolate(mask.to(torch.uint8), size=size, mode="nearest")
            loss_depth = loss_depth + criterion_depth(pd, d_gt, m_rs.to(torch.bool)) * args.depth_loss_weights[i]
        loss_seg = criterion_seg(outputs["pred_seg"], seg_gt.tensors.squeeze(1)) * args.seg_loss_weight
        losses = sum(loss_dict[k] * criterion.weight_dict[k] for k in loss_dict if k in criterion.weight_dict) + loss_depth + loss_seg
        assert math.isfinite(float(losses))
        optimizer.zero_grad()
        losses.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm)
        optimizer.step()
        log.append(float(losses))
    return {"loss": sum(log) / len(log)}


def test_drop_in_training_loop_and_fused_step():
    M, Trainer = _mods()
    B, H, W = 2, 128, 160
    sd = synth_weights()
    args = M.default_args(device="cuda", dropout=0.0, lr=1e-4, weight_decay=1e-4, clip_max_norm=0.1, input_log_freq=2.0,
                          with_plane_norm_loss=False)
    net, criterions, _ = M.build_model(args)
    net.load_state_dict(sd)
    net.cuda()
    # the reference's optimizer (src/main_glassrgbd.py:59-66)
    groups = [{"params": [p for n, p in net.named_parameters() if "backbone" not in n and p.requires_grad]},
              {"params": [p for n, p in net.named_parameters() if "backbone" in n and p.requires_grad], "lr": args.lr_backbone}]
    assert sum(len(g["params"]) for g in groups) == 738
    opt = torch.optim.AdamW(groups, lr=args.lr, weight_decay=args.weight_decay)
    before = {n: p.detach().clone() for n, p in net.named_parameters()}
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_shims
    if ref_shims.reference_available():        # the reference's own loop, unmodified
        ref_shims.install()
        from engine_glassrgbd import train_one_epoch
        from util.misc import NestedTensor
        stats = train_one_epoch(net, criterions, None, list(_loader(B, H, W, 2, NestedTensor)), opt, torch.device("cuda"), 0,
                                args.clip_max_norm, args)
    else:
        stats = _mirror_train_loop(net, criterions, _loader(B, H, W, 2, M.NestedTensor), opt, torch.device("cuda"), args.clip_max_norm, args)
    assert math.isfinite(stats["loss"])
    with_grad = {n for n, p in net.named_parameters() if p.grad is not None}
    assert len(with_grad) == 684, len(with_grad)                 # the 54 never-used tensors keep grad None, as in the reference
    # (the bias of the diffusion convolution has an analytically zero gradient -- the plane normalisation removes constant
    # shifts -- so whether round-off moves it is an accident of the summation order, here as in the reference)
    still = [n for n, p in net.named_parameters() if n in with_grad and torch.equal(p.detach(), before[n])
             and not n.endswith("ref_attn_diffusion.bias")]
    assert not still, still
    assert all(torch.equal(p.detach(), before[n]) for n, p in net.named_parameters() if n not in with_grad)

    # fused step == drop-in step on the same batch (same forward / backward kernels; torch criteria vs the fused loss kernels)
    net2, criterions2, _ = M.build_model(args)
    net2.load_state_dict(sd)
    net2.cuda().train()
    images, targets, depth_gt, seg_gt = synth.synth_batch(B, H, W, seed=0)
    tg = [{k: v.cuda() for k, v in t.items()} for t in targets]
    log = _mirror_train_loop(net2, criterions2, [(M.NestedTensor(images, torch.zeros(B, H, W, dtype=torch.bool)),
                                                  M.NestedTensor(depth_gt, None), M.NestedTensor(seg_gt, None), targets, ["x"])],
                             torch.optim.SGD([p for p in net2.parameters() if p.requires_grad], lr=0.0), torch.device("cuda"), 1e9, args)
    drop_grads = {n: p.grad.clone() for n, p in net2.named_parameters() if p.grad is not None}
    tr = Trainer(sd)
    total, losses = tr.train_step(images.cuda(), tg, depth_gt.cuda(), seg_gt.cuda(), criterions2[0].cuda())
    assert abs(float(total) - log["loss"]) < 2e-3 * abs(log["loss"]), (float(total), log["loss"])
    assert set(losses) >= {"loss_ce", "loss_line", "loss_ce_4", "loss_depth", "loss_seg"}
    # gradients were taken before the optimizer step of train_step modified the parameters: compare the kept flat buffers
    fused = tr.grads()
    errs = {n: rel_l2(fused[n], g) for n, g in drop_grads.items() if not n.endswith("ref_attn_diffusion.bias")}
    worst = sorted(errs.items(), key=lambda kv: -kv[1])[:5]
    assert worst[0][1] < 5e-2, worst
    # the CUDA-graph replay of the step (the default) equals the kernel-by-kernel step
    tr_e = Trainer(sd)
    tr_e.use_cuda_graph = False
    total_e, _ = tr_e.train_step(images.cuda(), tg, depth_gt.cuda(), seg_gt.cuda(), criterions2[0].cuda())
    assert tr.use_cuda_graph and abs(float(total) - float(total_e)) < 1e-4 * abs(float(total_e)), (float(total), float(total_e))
    eager = tr_e.grads()
    errs = {n: rel_l2(fused[n], eager[n]) for n in fused if not n.endswith("ref_attn_diffusion.bias")}
    worst = sorted(errs.items(), key=lambda kv: -kv[1])[:5]
    assert worst[0][1] < 1e-2, worst
    # the same for a RAGGED batch (mask as a static graph input), and a second shape gets its own captured step
    mask = torch.zeros(B, H, W, dtype=torch.bool)
    mask[-1, H - 32:, :] = True
    mask[-1, :, W - 32:] = True
    im_r = images.masked_fill(mask[:, None], 0.0).cuda()
    tr_g, tr_k = Trainer(sd), Trainer(sd)
    tr_k.use_cuda_graph = False
    tot_g, _ = tr_g.train_step(im_r, tg, depth_gt.cuda(), seg_gt.cuda(), criterions2[0].cuda(), mask=mask.cuda())
    tot_k, _ = tr_k.train_step(im_r, tg, depth_gt.cuda(), seg_gt.cuda(), criterions2[0].cuda(), mask=mask.cuda())
    assert abs(float(tot_g) - float(tot_k)) < 1e-4 * abs(float(tot_k)), (float(tot_g), float(tot_k))
    gg, gk = tr_g.grads(), tr_k.grads()
    worst = sorted(((rel_l2(gg[n], gk[n]), n) for n in gg if not n.endswith("ref_attn_diffusion.bias")), reverse=True)[:3]
    assert worst[0][0] < 1e-2, worst
    assert abs(float(tot_g) - float(total)) > 1e-6 * abs(float(total))            # the mask really entered
    tr_g.train_step(images.cuda(), tg, depth_gt.cuda(), seg_gt.cuda(), criterions2[0].cuda())      # un-masked: another graph key
    assert len(tr_g._graphs) == 2
    # parameters moved, and the module can be re-synchronised for evaluation / checkpoints
    net2.sync_from_trainer()
    sd_after = tr.state_dict()
    assert any(not torch.equal(sd_after[k].cpu(), sd[k]) for k in sd_after)


def test_reference_evaluate_loop_on_the_drop_in_and_fused_evaluator(tmp_path):
    """the reference's own `evaluate` (src/engine_glassrgbd.py:174-342, imported unmodified from baseline/_ref: batch 1, maps to
    the host, numpy compute_depth_errors, compute_mean_ioU, eval_results.txt) runs on the drop-in model, and the fused device-side
    bookkeeping (`evaluation.DenseEvaluator`: gwd_depth_metrics + gwd_seg_confusion) reports the same numbers"""
    M, _ = _mods()
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_shims
    if not ref_shims.reference_available():
        pytest.skip("reference tree not staged under baseline/_ref (oracle/stage_ref.sh)")
    ref_shims.install()
    from engine_glassrgbd import evaluate
    from util.misc import NestedTensor
    from gwdepth_b200 import evaluation
    n_img, H, W = 3, 224, 320
    args = M.default_args(device="cuda", dropout=0.0, coco_path=None, append_word=None, resume="", dataset="val")
    net, criterions, post = M.build_model(args)
    net.load_state_dict(synth_weights())
    net.cuda()
    batches = []
    for i in range(n_img):
        images, targets, depth_gt, seg_gt = synth.synth_batch(1, H, W, seed=40 + i)
        targets[0]["image_id"] = torch.tensor([i])
        mask = torch.zeros(1, H, W, dtype=torch.bool)
        batches.append((NestedTensor(images, mask), NestedTensor(depth_gt, mask), NestedTensor(seg_gt, mask), targets, ["img_%d" % i]))

    class Loader(list):
        dataset = type("D", (), {"id_to_img": {i: "img_%d" % i for i in range(n_img)}})()
    stats = evaluate(net, criterions, post, Loader(batches), None, torch.device("cuda"), str(tmp_path), args, save_dir=str(tmp_path))
    assert os.path.exists(os.path.join(str(tmp_path), "eval_results.txt"))
    ev = evaluation.DenseEvaluator(min_depth_eval=args.min_depth_eval, max_depth_eval=args.max_depth_eval)
    net.eval()
    with torch.no_grad():
        for samples, depth_gt, seg_gt, _, _ in batches:
            ev.update(net(samples.tensors.cuda()), depth_gt.tensors.cuda(), seg_gt.tensors.cuda())
    mine = ev.summary()
    for k in ("silog", "abs_rel", "log10", "rms", "sq_rel", "log_rms", "d1", "d2", "d3"):
        assert abs(float(stats[k]) - mine[k]) <= 2e-5 * max(1.0, abs(mine[k])), (k, float(stats[k]), mine[k])
    assert abs(float(stats["Mean IU"]) - mine["Mean IU"]) < 1e-3 and abs(float(stats["Pixel accuracy"]) - mine["Pixel accuracy"]) < 1e-3
    assert math.isfinite(float(stats["loss"])) and mine["images"] == n_img
