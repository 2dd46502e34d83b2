"""Times the DENSE-TAIL training step (train_tail.DenseTail: point_based_pred2 + pyramid (K = 80) + dense head + three losses,
forward + backward + one-norm clip + AdamW) at batch B x 480x640 with CUDA events on the launching stream; optional
per-entry-point breakdown from events around every C-ABI call of one step.
    python tools/bench_train_tail.py [--batch 16] [--steps 10] [--breakdown] [--json out.json]
With GWD_PROFILE_ONE=1 it runs one warm-up step and one step only (for `ncu --metrics gpu__time_duration.sum`)."""
import argparse
import json
import os
import sys
import time
from collections import defaultdict

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def build(B, H=480, W=640, seed=3):
    from helpers import synth_weights
    import gwdepth_b200  # noqa: F401
    from gwdepth_b200.engine import sine_table
    from gwdepth_b200.train_tail import DenseTail
    g = torch.Generator().manual_seed(seed)
    H4, W4 = H // 4, W // 4
    buf = torch.zeros(B, H4, W4, 256)
    buf[..., :192] = torch.randn(B, H4, W4, 192, generator=g)
    depth2 = torch.rand(B, H4 // 2, W4 // 2, generator=g) * 0.9 + 0.05
    coords = torch.rand(B, 80, 2, generator=g) * 2 - 1
    depth_gt = torch.rand(B, 1, H, W, generator=g) * 9.5 + 0.3
    seg_gt = (torch.rand(B, 1, H, W, generator=g) > 0.5).long()
    sd = {k: v.cuda() for k, v in synth_weights().items()
          if k.startswith("depth_decoder.") or k.startswith("dense_encoder.point_based_pred2.")}
    tail = DenseTail(sd)
    args = (buf.bfloat16().cuda(), depth2.cuda(), coords.cuda(), sine_table(H4, W4, 32, False, "cuda"), depth_gt.cuda(), seg_gt.cuda())
    return tail, args


def timed(fn, n, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, (time.perf_counter() - t0) * 1000 / n


def measure(B, steps, H=480, W=640):
    from gwdepth_b200 import capi
    tail, args = build(B, H, W)
    if os.environ.get("GWD_PROFILE_ONE"):
        for _ in range(2):
            tail.train_step(*args)
            torch.cuda.synchronize()
        return None, tail, args
    lg = timed(lambda: tail.loss_and_grads(*args), steps)
    opt = timed(tail.step, steps)
    capi.reset_launch_count()
    losses = tail.train_step(*args)
    torch.cuda.synchronize()
    launches = capi.launch_count()
    full = timed(lambda: tail.train_step(*args), steps)
    ms = max(full)
    res = {"metric": "images_per_sec_train_dense_tail_480x640_bf16", "value": B / (ms / 1000.0), "unit": "images/s", "batch": B,
           "ms_per_step": ms, "device_ms_per_step": full[0], "host_ms_per_step": full[1],
           "breakdown_ms": {"forward_losses_backward": lg[0], "allreduce_clip_adamw": opt[0]},
           "gpu_launches_per_step": launches, "losses": [float(v) for v in losses.tolist()],
           "params": int(sum(m.numel for m in tail.modules())), "peak_memory_gb": torch.cuda.max_memory_allocated() / 2 ** 30,
           "scope": "dense tail only: point_based_pred2 (+ PyramidLayer, K=80) at 1/4 scale, DensePrediction head to full resolution, "
                    "silog(depth_pred3) x 0.25 + silog(depth) + 2 x seg CE; gradients returned for the stage buffer and depth_pred2 "
                    "(Swin-stage / backbone backward not built); lr 1e-4, weight decay 1e-4, clip 0.1 as the reference"}
    return res, tail, args


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--breakdown", action="store_true")
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    from gwdepth_b200 import capi
    res, tail, args = measure(a.batch, a.steps)
    if res is None:
        return
    print(json.dumps(res))
    if a.json:
        with open(a.json, "w") as f:
            json.dump(res, f, indent=1)
    if a.breakdown:
        lib = capi.lib()
        rec = []
        names = [n for n in capi.SIGNATURES if n not in ("gwd_last_error", "gwd_version", "gwd_launch_count", "gwd_reset_launch_count")]
        orig = {n: getattr(lib, n) for n in names}

        class Wrapped:
            def __getattr__(self, n):
                f = orig.get(n)
                if f is None:
                    return getattr(lib, n)

                def call(*cargs):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    r = f(*cargs)
                    e1.record()
                    rec.append((n, e0, e1))
                    return r
                return call
        capi._lib = Wrapped()
        tail.train_step(*args)
        torch.cuda.synchronize()
        capi._lib = lib
        agg = defaultdict(lambda: [0.0, 0])
        for n, e0, e1 in rec:
            agg[n][0] += e0.elapsed_time(e1)
            agg[n][1] += 1
        tot = sum(v[0] for v in agg.values())
        print("per entry point (events around each call, %d calls, %.2f ms):" % (len(rec), tot))
        for n, (ms, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
            print("  %8.3f ms %5.1f%% x%-3d %s" % (ms, 100 * ms / tot, c, n))
        print("slowest calls:")
        for i, (n, e0, e1) in sorted(enumerate(rec), key=lambda r: -r[1][1].elapsed_time(r[1][2]))[:16]:
            print("  %8.3f ms %s (call #%d)" % (e0.elapsed_time(e1), n, i))


if __name__ == "__main__":
    main()
