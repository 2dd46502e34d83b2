"""BASELINE config 5: evaluation sweep, batch 1..64 at 480x640 and 960x1280 -- forward + the fused depth-metric
reduction (gwd_depth_metrics, the per-image compute_depth_errors of src/util/metrics.py:197-218).
usage: python tools/eval_sweep.py [max_batch_480] [max_batch_960]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
from helpers import synth, synth_weights  # noqa: E402
import gwdepth_b200  # noqa: F401,E402
from gwdepth_b200 import model as M, ops  # noqa: E402

net, _, _ = M.build_model(M.default_args(device="cuda"))
net.load_state_dict(synth_weights())
net.cuda().eval()
limits = {(480, 640): int(sys.argv[1]) if len(sys.argv) > 1 else 64, (960, 1280): int(sys.argv[2]) if len(sys.argv) > 2 else 16}
for (H, W), bmax in limits.items():
    for B in (1, 2, 4, 8, 16, 32, 64):
        if B > bmax:
            continue
        images, _, depth_gt, _ = synth.synth_batch(B, H, W, seed=5)
        x, gt = images.cuda(), depth_gt.cuda()
        with torch.no_grad():
            for _ in range(2):
                out = net(x)
                m = ops.depth_metrics(out["pred_depth"][3][:, 0].contiguous(), gt[:, 0].contiguous())
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            iters = 5
            e0.record()
            for _ in range(iters):
                out = net(x)
                m = ops.depth_metrics(out["pred_depth"][3][:, 0].contiguous(), gt[:, 0].contiguous())
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        print(json.dumps({"size": [H, W], "batch": B, "ms_per_step": round(ms, 3), "images_per_s": round(B / ms * 1000, 1),
                          "abs_rel_mean": float(m[:, 1].mean()), "finite": bool(torch.isfinite(m).all())}), flush=True)
        del out, x, gt
        torch.cuda.empty_cache()
