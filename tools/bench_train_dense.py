"""Times the dense-prediction-head training step (train_dense.DenseHead: forward + SilogLoss + SegLoss + backward + clip +
AdamW) at batch B x 480x640 with CUDA events on the launching stream, and prints a per-kernel breakdown from events
around every C-ABI call of one step.
    python tools/bench_train_dense.py [--batch 16] [--steps 10] [--breakdown]
With GWD_PROFILE_ONE=1 it runs one warm-up step and one step only (for `ncu --metrics gpu__time_duration.sum`)."""
import argparse
import os
import sys
import time
from collections import defaultdict

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--breakdown", action="store_true")
    a = ap.parse_args()
    from helpers import synth_weights
    import gwdepth_b200  # noqa: F401
    from gwdepth_b200 import capi
    from gwdepth_b200.train_dense import DenseHead
    B, H, W = a.batch, 480, 640
    g = torch.Generator().manual_seed(3)
    buf = torch.zeros(B, H // 4, W // 4, 256)
    buf[..., :193] = torch.randn(B, H // 4, W // 4, 193, generator=g)
    buf[..., 192] = torch.rand(B, H // 4, W // 4, generator=g)
    buf = buf.bfloat16().cuda()
    depth_gt = (torch.rand(B, 1, H, W, generator=g) * 9.5 + 0.3).cuda()
    seg_gt = (torch.rand(B, 1, H, W, generator=g) > 0.5).long().cuda()
    head = DenseHead({k: v.cuda() for k, v in synth_weights().items() if k.startswith("depth_decoder.")})

    def timed(fn, n=a.steps, warm=3):
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

    if os.environ.get("GWD_PROFILE_ONE"):
        head.train_step(buf, depth_gt, seg_gt)
        torch.cuda.synchronize()
        head.train_step(buf, depth_gt, seg_gt)
        torch.cuda.synchronize()
        return
    fwd = timed(lambda: head.forward(buf, H, W))
    lg = timed(lambda: head.loss_and_grads(buf, depth_gt, seg_gt))
    opt = timed(head.step)
    capi.reset_launch_count()
    head.train_step(buf, depth_gt, seg_gt)
    torch.cuda.synchronize()
    launches = capi.launch_count()
    full = timed(lambda: head.train_step(buf, depth_gt, seg_gt))
    print("dense head, batch %d x %dx%d: forward %.2f ms | forward + losses + backward %.2f ms (host %.2f) | clip+AdamW %.3f ms | "
          "train_step %.2f ms (host %.2f) | %d launches per step | %.0f images/s | peak memory %.1f GB"
          % (B, H, W, fwd[0], lg[0], lg[1], opt[0], full[0], full[1], launches, B / (max(full) / 1000),
             torch.cuda.max_memory_allocated() / 2 ** 30))
    if a.breakdown:
        # events around every C-ABI call of one step
        lib = capi.lib()
        rec = []
        names = [n for n in capi.SIGNATURES if n not in ("gwd_last_error", "gwd_version", "gwd_launch_count", "gwd_reset_launch_count")]
        orig = {n: getattr(lib, n) for n in names}

        class Wrapped:
            def __getattr__(self, n):
                f = orig.get(n)
                if f is None:
                    return getattr(lib, n)

                def call(*args):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    r = f(*args)
                    e1.record()
                    rec.append((n, e0, e1))
                    return r
                return call
        capi._lib = Wrapped()
        head.train_step(buf, depth_gt, seg_gt)
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
        for n, e0, e1 in sorted(rec, key=lambda r: -r[1].elapsed_time(r[2]))[:14]:
            print("  %8.3f ms %s (call #%d)" % (e0.elapsed_time(e1), n, rec.index((n, e0, e1))))


if __name__ == "__main__":
    main()
