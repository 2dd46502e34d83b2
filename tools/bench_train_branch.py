"""Times the DENSE-BRANCH training step behind the 1/32 stage (train_branch.DenseBranch: three class-window Swin stages with
their entries, coarse depth head, both point-based predictions + uncertainty sampling, dense head, five losses; forward +
backward + one-norm clip + AdamW) at batch B x 480x640 with CUDA events on the launching stream; optional per-entry-point
breakdown from events around every C-ABI call of one step.
    python tools/bench_train_branch.py [--batch 16] [--steps 10] [--breakdown] [--json out.json]"""
import argparse
import json
import os
import sys
from collections import defaultdict

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def init_dist():
    """one process per GPU under torchrun (NCCL over NVLink); returns (rank, world)"""
    if int(os.environ.get("WORLD_SIZE", "1")) <= 1:
        return 0, 1
    import torch.distributed as dist
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
    return dist.get_rank(), dist.get_world_size()


def build(B, H=480, W=640, seed=5):
    from helpers import synth_weights
    import gwdepth_b200  # noqa: F401
    from gwdepth_b200.train_branch import DenseBranch
    g = torch.Generator().manual_seed(seed)
    h5, w5 = H // 32, W // 32
    x32 = torch.randn(B, h5, w5, 512, generator=g).bfloat16().cuda()
    depth0 = (torch.rand(B, h5, w5, generator=g) * 0.9 + 0.05).cuda()
    feats = [torch.randn(B, (2 ** k) * h5, (2 ** k) * w5, c, generator=g).bfloat16().cuda() for k, c in ((1, 1024), (2, 512), (3, 256))]
    depth_gt = (torch.rand(B, 1, H, W, generator=g) * 9.5 + 0.3).cuda()
    seg_gt = (torch.rand(B, 1, H, W, generator=g) > 0.5).long().cuda()
    live = ("dense_encoder.class_transformer", "dense_encoder.point_based_pred", "dense_encoder.proj_", "dense_encoder.old_",
            "dense_encoder.depth_pred16", "dense_encoder.depth_token", "dense_encoder.seg_token", "depth_decoder.")
    sd = {k: v.cuda() for k, v in synth_weights().items() if k.startswith(live)}
    return DenseBranch(sd, cuda_graph=os.environ.get("GWD_CUDA_GRAPH", "1") != "0"), (x32, depth0, feats, depth_gt, seg_gt)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--breakdown", action="store_true")
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    from bench_train_tail import timed
    rank, world = init_dist()
    from gwdepth_b200 import capi
    br, args = build(a.batch, seed=5 + rank)          # same weights on every rank, a different batch per rank
    graphed = br.use_cuda_graph
    br.use_cuda_graph = False               # launch count and the eager time first
    capi.reset_launch_count()
    losses = br.train_step(*args)
    torch.cuda.synchronize()
    launches = capi.launch_count()
    eager = timed(lambda: br.train_step(*args), a.steps)
    br.use_cuda_graph = graphed
    lg = timed(lambda: br.loss_and_grads(*args), a.steps)
    opt = timed(br.step, a.steps)
    full = timed(lambda: br.train_step(*args), a.steps)
    losses = br.train_step(*args).clone()
    ms = max(full)
    replicas_equal = None
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        # data-parallel invariant: after the same number of steps every rank holds bit-identical parameters
        replicas_equal = True
        for m in br.modules():
            ref = m.P.clone()
            dist.broadcast(ref, 0)
            replicas_equal = replicas_equal and bool(torch.equal(ref, m.P))
        flag = torch.tensor([1.0 if replicas_equal else 0.0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        replicas_equal = bool(flag.item() == 1.0)
    res = {"metric": "images_per_sec_train_dense_branch_480x640_bf16", "value": world * a.batch / (ms / 1000.0), "unit": "images/s",
           "n_gpus": world, "replicas_bit_identical_after_training": replicas_equal,
           "allreduce_bytes_per_step": int(sum(m.numel for m in br.modules())) * 4 if world > 1 else 0,
           "batch": a.batch, "ms_per_step": ms, "device_ms_per_step": full[0], "host_ms_per_step": full[1],
           "breakdown_ms": {"forward_losses_backward": lg[0], "allreduce_clip_adamw": opt[0]}, "gpu_launches_per_step": launches,
           "cuda_graph": bool(graphed), "eager_ms_per_step": max(eager), "losses": [float(v) for v in losses.tolist()], "params": int(sum(m.numel for m in br.modules())),
           "peak_memory_gb": torch.cuda.max_memory_allocated() / 2 ** 30,
           "scope": "dense branch behind the 1/32 line-window stage: entries + class-window Swin stages at 1/16, 1/8, 1/4, depth_pred16, "
                    "point_based_pred1/2 (+ PyramidLayer K=30 / 80), CertainSample, DensePrediction head, 4 silog losses + seg CE; "
                    "gradients returned for x32, C4, C3 (line-window stage / backbone backward not built)"}
    if rank == 0:
        print(json.dumps(res))
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
        return
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
        br.use_cuda_graph = False           # the per-call events need the eager launch sequence
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        br.train_step(*args)
        e1.record()
        torch.cuda.synchronize()
        capi._lib = lib
        agg = defaultdict(lambda: [0.0, 0])
        for n, s0, s1 in rec:
            agg[n][0] += s0.elapsed_time(s1)
            agg[n][1] += 1
        tot = sum(v[0] for v in agg.values())
        print("per entry point (events around each call, %d calls, %.2f ms inside C-ABI calls of %.2f ms step):" % (len(rec), tot, e0.elapsed_time(e1)))
        for n, (t, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
            print("  %8.3f ms %5.1f%% x%-3d %s" % (t, 100 * t / tot, c, n))
        print("slowest calls:")
        for i, (n, s0, s1) in sorted(enumerate(rec), key=lambda r: -r[1][1].elapsed_time(r[1][2]))[:40]:
            print("  %8.3f ms %s (call #%d)" % (s0.elapsed_time(s1), n, i))


if __name__ == "__main__":
    main()
