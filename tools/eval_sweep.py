"""BASELINE config 5: evaluation sweep, batch 1..64 at 480x640 and 960x1280 -- forward (one CUDA-graph replay per batch) + the
whole evaluation bookkeeping of `evaluate` (src/engine_glassrgbd.py:232-264,309-322) on the device: `evaluation.DenseEvaluator`
= gwd_depth_metrics (the per-image compute_depth_errors of src/util/metrics.py:197-218) + gwd_seg_confusion.
Per size, a parity check first: image 0 of the batch against the fp32 CPU oracle (selections pinned), and the per-image metrics
of the batch against the batch-1 loop the reference runs.
usage: python tools/eval_sweep.py [max_batch_480] [max_batch_960] > profiles/rN_eval_sweep_config5.jsonl"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
from helpers import oracle, synth, synth_weights  # noqa: E402
import gwdepth_b200  # noqa: F401,E402
from gwdepth_b200 import evaluation, model as M  # noqa: E402

net, _, _ = M.build_model(M.default_args(device="cuda"))
net.load_state_dict(synth_weights())
net.cuda().eval()
limits = {(480, 640): int(sys.argv[1]) if len(sys.argv) > 1 else 64, (960, 1280): int(sys.argv[2]) if len(sys.argv) > 2 else 64}
for (H, W), bmax in limits.items():
    # ---- parity at this size
    images, _, depth_gt, seg_gt = synth.synth_batch(2, H, W, seed=5)
    trace = {}
    ref = oracle.forward(synth_weights(), images[:1], trace=trace)
    pin = {"line_ids": trace["line_ids"].cuda(), "sample1": trace["sample1"].cuda(), "sample2": trace["sample2"].cuda()}
    with torch.no_grad():
        out = net(images[:1].cuda(), _pinned=pin)
        d, r = out["pred_depth"][3].float().cpu(), ref["pred_depth"][3]
        err = float((d - r).abs().mean() / r.abs().mean())
        both = net(images.cuda())
        ev2 = evaluation.DenseEvaluator()
        per2 = ev2.update(both, depth_gt.cuda(), seg_gt.cuda()).cpu()
        solo = torch.cat([evaluation.DenseEvaluator().update({"pred_depth": [both["pred_depth"][3][b:b + 1]], "pred_seg": both["pred_seg"][b:b + 1]},
                                                             depth_gt[b:b + 1].cuda(), seg_gt[b:b + 1].cuda()).cpu() for b in range(2)])
    oracle_metrics = torch.tensor(oracle.depth_metrics(both["pred_depth"][3][0], depth_gt[0]))
    print(json.dumps({"size": [H, W], "parity": {"depth_mean_rel_vs_fp32_oracle": round(err, 5),
                                                  "batch_metrics_equal_batch1_loop_rtol_1e-9": bool(torch.allclose(per2, solo, rtol=1e-9, atol=0)),
                                                  "metrics_vs_numpy_oracle_max_rel": float(((per2[0] - oracle_metrics).abs() / oracle_metrics.abs().clamp_min(1e-9)).max())}}),
          flush=True)
    assert err < 2e-2 and torch.allclose(per2, solo, rtol=1e-9, atol=0)      # (fp64 atomics: the summation order may differ)
    for B in (1, 2, 4, 8, 16, 32, 64):
        if B > bmax:
            continue
        try:
            images, _, depth_gt, seg_gt = synth.synth_batch(B, H, W, seed=5)
            x, gt, sg = images.cuda(), depth_gt.cuda(), seg_gt.cuda()
            ev = evaluation.DenseEvaluator()
            with torch.no_grad():
                for _ in range(2):
                    ev.update(net(x), gt, sg)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                iters = 5
                e0.record()
                for _ in range(iters):
                    ev.update(net(x), gt, sg)
                e1.record()
                torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / iters
            s = ev.summary()
            print(json.dumps({"size": [H, W], "batch": B, "ms_per_step": round(ms, 3), "images_per_s": round(B / ms * 1000, 1),
                              "abs_rel": round(s["abs_rel"], 5), "Mean IU": round(s["Mean IU"], 3), "images": s["images"],
                              "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 1)}), flush=True)
        except torch.OutOfMemoryError as e:
            print(json.dumps({"size": [H, W], "batch": B, "error": "out of memory"}), flush=True)
        x = gt = sg = ev = None
        net._plan._graphs.clear() if net._plan is not None else None
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats()
