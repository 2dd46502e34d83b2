"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/* by running the UNMODIFIED reference.

Run in the build container (where /root/reference is mounted):   python oracle/make_golden.py
The GPU box has no reference tree, so what it produces is committed:
  tests/golden/state_dict_spec.json   the reference model's state_dict keys / shapes / dtypes, parameter
                                      names and which of them are trainable (boundary contract)
  tests/golden/fwd_224x320_b2.npz     full outputs of reference model(images) + criterion/loss values,
                                      synthetic weights seed 0, synthetic batch seed 0 (oracle/synth.py)
  tests/golden/fwd_480x640_b1.npz     BASELINE config 1 (1x3x480x640): small outputs in full, dense maps
                                      sub-sampled every 4th pixel
  tests/golden/depth_metrics.npz      reference util/metrics.compute_depth_errors on synthetic maps
Inputs are not stored: they are regenerated from the seeds by oracle/synth.py.
"""
import json
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shims  # noqa: E402
import synth  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")


def reference_model(seed=0):
    torch.manual_seed(0)
    model, criterions, post, args = ref_shims.build_reference()
    model.eval()
    sd0 = model.state_dict()
    spec = [(k, list(v.shape), str(v.dtype).replace("torch.", "")) for k, v in sd0.items()]
    model.load_state_dict(synth.synth_state_dict(spec, seed=seed), strict=False)
    return model, criterions, args, spec


def dump_spec(model, spec):
    trainable = sorted(n for n, p in model.named_parameters() if p.requires_grad)
    params = sorted(n for n, _ in model.named_parameters())
    with open(os.path.join(GOLDEN, "state_dict_spec.json"), "w") as f:
        json.dump({"keys": spec, "params": params, "trainable": trainable}, f)


def reference_losses(criterions, args, out, targets, depth_gt, seg_gt):
    """the loss terms the reference's training step forms (engine_glassrgbd.py:62-90), computed with the
    reference's own criterion modules"""
    set_crit, depth_crit, seg_crit, _ = criterions
    vals = {k: float(v) for k, v in set_crit(out, targets, depth_gt=depth_gt).items()}
    valid = (depth_gt >= 0.2) & (depth_gt < 10.0)
    for i, pd in enumerate(out["pred_depth"]):
        size = pd.shape[-2:]
        g = F.interpolate(depth_gt, size=size, mode="nearest")
        m = F.interpolate(valid.to(torch.uint8), size=size, mode="nearest").to(torch.bool)
        vals["loss_depth_%d" % i] = float(depth_crit(pd, g, m) * args.depth_loss_weights[i])
    vals["loss_seg"] = float(seg_crit(out["pred_seg"], seg_gt.squeeze(1)) * args.seg_loss_weight)
    return vals


def run_case(model, criterions, args, B, H, W, stride, name):
    images, targets, depth_gt, seg_gt = synth.synth_batch(B, H, W, seed=0)
    with torch.no_grad():
        out = model(images)
        losses = reference_losses(criterions, args, out, targets, depth_gt, seg_gt)
        match = criterions[0].matcher({k: v for k, v in out.items() if k != "aux_outputs"}, targets)
    arrs = {
        "pred_logits": out["pred_logits"].numpy(), "pred_lines": out["pred_lines"].numpy(),
        "aux_logits": np.stack([a["pred_logits"].numpy() for a in out["aux_outputs"]]),
        "aux_lines": np.stack([a["pred_lines"].numpy() for a in out["aux_outputs"]]),
        "pred_seg": out["pred_seg"][..., ::stride, ::stride].numpy(),
        "dense_stride": np.int64(stride),
        "loss_names": np.array(sorted(losses)), "loss_values": np.array([losses[k] for k in sorted(losses)], dtype=np.float64),
    }
    for i, d in enumerate(out["pred_depth"]):
        arrs["pred_depth_%d" % i] = (d[..., ::stride, ::stride] if i == 3 else d).numpy()
    # L1 matching costs are piecewise linear, so optimal assignments can be EXACTLY tied; record which images
    # have an assignment that survives 1e-4 cost noise (only those are compared index-by-index)
    from scipy.optimize import linear_sum_assignment
    prob = out["pred_logits"].softmax(-1)
    robust = []
    for b, (i, j) in enumerate(match):
        arrs["match_pred_%d" % b] = i.numpy()
        arrs["match_tgt_%d" % b] = j.numpy()
        C = (args.set_cost_line * torch.cdist(out["pred_lines"][b], targets[b]["lines"], p=1)
             - args.set_cost_class * prob[b][:, :1]).numpy()
        arrs["match_cost_%d" % b] = np.float64(C[i.numpy(), j.numpy()].sum())
        ok = True
        for trial in range(16):
            rng = np.random.default_rng(trial)
            i2, j2 = linear_sum_assignment(C + 1e-4 * rng.standard_normal(C.shape))
            ok = ok and np.array_equal(i2, i.numpy()) and np.array_equal(j2, j.numpy())
        robust.append(ok)
    arrs["match_robust"] = np.array(robust)
    np.savez_compressed(os.path.join(GOLDEN, name), **arrs)
    print(name, {k: float(v) for k, v in list(losses.items())[:3]}, "depth range",
          float(out["pred_depth"][3].min()), float(out["pred_depth"][3].max()))


def depth_metric_case():
    from util.metrics import compute_depth_errors  # reference module
    g = torch.Generator().manual_seed(77)
    gt = (torch.rand(3, 60, 80, generator=g) * 11.0).numpy().astype(np.float32)          # some pixels invalid (>10)
    gt[:, :5, :7] = 0.0
    pred = (gt * (1 + 0.2 * torch.randn(3, 60, 80, generator=g).numpy()) + 0.05).astype(np.float32)
    pred[0, 10, 10] = np.nan
    pred[1, 11, 11] = np.inf
    pred[2, 12, 12] = -3.0
    rows = []
    for b in range(3):
        p = pred[b].copy()
        p[p < 1e-3] = 1e-3
        p[p > 10.0] = 10.0
        p[np.isinf(p)] = 10.0
        p[np.isnan(p)] = 1e-3
        valid = np.logical_and(gt[b] > 1e-3, gt[b] < 10.0)
        rows.append(compute_depth_errors(gt[b][valid], p[valid]))
    np.savez_compressed(os.path.join(GOLDEN, "depth_metrics.npz"), pred=pred, gt=gt,
                        metrics=np.array(rows, dtype=np.float64))


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    model, criterions, args, spec = reference_model()
    dump_spec(model, spec)
    run_case(model, criterions, args, 2, 224, 320, 1, "fwd_224x320_b2.npz")
    run_case(model, criterions, args, 1, 480, 640, 4, "fwd_480x640_b1.npz")
    depth_metric_case()


if __name__ == "__main__":
    main()
