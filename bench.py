#!/usr/bin/env python
"""bench.py -- images/sec of the GW-Depth forward hot path at 480x640, bf16, batch 16 per GPU
(BASELINE.json configs[1]: "stage-1 ResNet-50 inference bf16 batch 16 synthetic GlassRGBD-shaped 480x640 on 1xB200").

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torch.distributed.run, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W

A step = one forward of the model over one batch of 16 synthetic images per GPU.
  value     : images/s with the input batches already resident in HBM (4 distinct batches are rotated: 236 MB of
              inputs > the 126 MB L2, and the step's activation working set is several GB)
  e2e       : the same metric through the public API (`model(samples)`) from PINNED HOST buffers: the H2D copy of the
              step's images and the D2H read of its outputs (line logits / end points, full-resolution depth and
              segmentation) are inside the timed region
  roofline  : the tcgen05 implicit-GEMM kernel (gwd_tapgemm_kernel, every launch of the step): algorithmic FLOPs
              (2*M*N*K*taps on logical dims) / sum of per-launch CUDA-event durations, against the measured bf16 peak
  cpu_baseline : the CPU oracle (oracle/gwdepth_oracle.py, a port of the reference's PyTorch forward) on the host cores
The reference arm (--impl reference) times that same CPU oracle; the reference is Python and is not on the GPU box.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

METRIC = "images_per_sec_fwd_480x640_bf16"
UNIT = "images/s"
BATCH, H, W = 16, 480, 640
WORKLOAD = "GW-Depth stage-1 ResNet-50 line+depth model, inference forward, batch 16 per GPU, 480x640 (BASELINE configs[1])"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=("b200", "reference"))
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the line-branch training-step measurement")
    return ap.parse_args()


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons sampled DURING the timed region through NVML (B200_PROFILING.md clocks line)"""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], 0, False, None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            while not self.stop_flag:
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                self.reasons |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
                time.sleep(0.005)
        except Exception as e:  # noqa: BLE001
            self.error = repr(e)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml unavailable: %s" % getattr(self, "error", "")]}
        bits = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
        sm = sorted(self.samples)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.max_mhz, "reasons": [n for b, n in bits.items() if self.reasons & b],
                "samples": len(sm)}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return float(p.get("bf16_tflops_sustained", p.get("bf16_tflops"))), "measured (MEASURED_PEAKS.json, sustained bf16)"
    return 1590.0, "fallback (B200_PROFILING.md, 1.59 PFLOP/s)"


def use_all_host_threads():
    """torch.distributed.run exports OMP_NUM_THREADS=1; the CPU arm is entitled to every core of the box"""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    torch.set_num_threads(max(n, 1))
    return torch.get_num_threads()


def cpu_oracle_rate(max_seconds=20.0, min_iters=2):
    """images/s of the CPU oracle at B=1, 480x640 (a bounded sample of the batch-16 workload)"""
    from helpers import oracle, synth, synth_weights
    use_all_host_threads()
    sd = synth_weights()
    images, _, _, _ = synth.synth_batch(1, H, W, seed=0)
    oracle.forward(sd, images)  # warm
    t0, n = time.time(), 0
    while n < min_iters or (time.time() - t0 < max_seconds and n < 12):
        oracle.forward(sd, images)
        n += 1
    dt = time.time() - t0
    return n / dt, n


def run_reference(args, rank):
    """the reference arm: the CPU port of the reference forward (oracle), all host threads, rank 0 only"""
    if rank != 0:
        return
    from helpers import oracle, synth, synth_weights
    use_all_host_threads()
    sd = synth_weights()
    images, _, _, _ = synth.synth_batch(1, H, W, seed=0)
    for _ in range(min(args.warmup, 2)):
        oracle.forward(sd, images)
    t0 = time.time()
    for _ in range(args.steps):
        oracle.forward(sd, images)
    dt = time.time() - t0
    v = args.steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample": "1 image of the 16-image batch per step (CPU)"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                             "sample": "%d forwards of 1x3x480x640 through oracle/gwdepth_oracle.py" % args.steps},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_train(args, net, resident, rank, world, dev, barrier, reduce_max_ms):
    """Data-parallel training step of the LINE BRANCH (gw-depth_b200/train.py): backbone forward (no gradient: the
    backbone / dense-branch backward is not built), input_proj -> encoder -> decoder -> heads forward with saved
    activations, SetCriterion with its 6 Hungarian matchings (scipy, host), backward kernels, ONE NCCL all-reduce of the
    flat gradient buffer, fused clip + AdamW.  Reported as images/s over all ranks, device-timed, max over ranks."""
    from helpers import synth, synth_weights
    from gwdepth_b200 import capi, model as M, train
    B = args.batch
    _, crit, _ = M.build_model(M.default_args(device="cuda", dropout=0.0))
    criterion = crit[0].to(dev)
    lb = train.LineBranch(synth_weights(), net.cfg, device=dev)
    targets = [[{k: v.to(dev) for k, v in t.items()} for t in synth.synth_batch(B, H, W, seed=100 + 7 * rank + i)[1]]
               for i in range(len(resident))]
    plan = net.plan()

    backbone = lambda images: plan.backbone(images)[3]      # noqa: E731  (no gradient: captured into the forward graph)

    def step(i):
        return lb.train_step(resident[i % len(resident)], targets[i % len(resident)], criterion, producer=backbone)

    for i in range(max(args.warmup, 3)):
        step(i)
    barrier()
    capi.reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for i in range(args.steps):
        total, _ = step(i)
    e1.record()
    barrier()
    ms = reduce_max_ms(max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1000.0))
    # where the time goes on this rank (separate, un-timed-above passes)
    def timed(fn, n=5):
        fn()
        torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(n):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t) * 1000.0 / n
    with torch.no_grad():
        c5 = plan.backbone(resident[0])[3].clone()
        ms_backbone = timed(lambda: plan.backbone(resident[0]))
    st = lb._captured(c5) if lb.use_cuda_graph else None
    ms_graphs = timed(lambda: (st["fwd"].replay(), st["bwd"].replay())) if st else None
    lo, li = (st["logits"], st["lines"]) if st else lb.forward(c5)
    lo, li = lo.detach().clone(), li.detach().clone()
    ms_crit = timed(lambda: criterion.forward_backward_stacked(lo, li, targets[0]))
    return {"metric": "images_per_sec_train_line_branch_480x640_bf16", "value": world * B * args.steps / (ms / 1000.0), "unit": UNIT,
            "ms_per_step": ms / args.steps, "n_gpus": world, "global_batch": world * B, "loss": float(total),
            "params": lb.numel, "allreduce_bytes_per_step": lb.numel * 4 if world > 1 else 0,
            "breakdown_ms": {"backbone_forward_no_grad": ms_backbone, "branch_forward_plus_backward_graph_replays": ms_graphs,
                             "set_criterion_6_hungarian_host": ms_crit},
            "scope": "line branch only: gradients stop at the C5 map (backbone / dense-branch backward not built); "
                     "dropout 0; lr 1e-4, weight decay 1e-4, clip 0.1 as the reference"}


def main():
    args = parse()
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    import torch.distributed as dist
    from helpers import synth, synth_weights
    import gwdepth_b200  # noqa: F401
    from gwdepth_b200 import capi, model as M, ops

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch

    net, _, _ = M.build_model(M.default_args(device="cuda"))
    net.load_state_dict(synth_weights())
    net.to(dev).eval()
    NB = 4
    host = [synth.synth_batch(B, H, W, seed=100 + 7 * rank + i)[0].pin_memory() for i in range(NB)]
    resident = [h.to(dev) for h in host]
    plan = net.plan()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max_ms(ms):
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    # ------------------------------------------------------------ device-resident throughput
    with torch.no_grad():
        step = plan.forward_graphed if net.use_cuda_graph else plan.forward
        for i in range(max(args.warmup, 3)):
            step(resident[i % NB])
        barrier()
        sampler = ClockSampler(local_rank)
        sampler.start()
        capi.reset_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for i in range(args.steps):
            step(resident[i % NB])
        e1.record()
        barrier()
        launches = capi.launch_count()
        if net.use_cuda_graph:   # replayed launches are not re-issued through the C ABI: count one eager step instead
            capi.reset_launch_count()
            plan.forward(resident[0])
            torch.cuda.synchronize()
            launches = capi.launch_count() * args.steps
        sampler.stop_flag = True
        ms = reduce_max_ms(e0.elapsed_time(e1))
        value = world * B * args.steps / (ms / 1000.0)

        # -------------------------------------------------------- end to end through the public API, host buffers
        # net.infer_stream: the serving loop of the public API (H2D of batch i+1 and D2H of batch i-1 overlap the forward
        # of batch i on three streams); every step's upload and read-back are inside the timed region
        def e2e_run(n):
            last = None
            for last in net.infer_stream(host[i % NB] for i in range(n)):
                pass
            torch.cuda.synchronize()
            return last
        out_host = list(e2e_run(3).values())
        barrier()
        t0 = time.perf_counter()
        e0.record()
        e2e_run(args.steps)
        e1.record()
        barrier()
        ms_e2e = reduce_max_ms(max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1000.0))
        e2e_value = world * B * args.steps / (ms_e2e / 1000.0)
        # the same loop fed with raw uint8 HWC images (a quarter of the PCIe bytes; normalised on the GPU)
        host_u8 = [torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(i)).pin_memory()
                   for i in range(NB)]

        def e2e_u8(n):
            for _ in net.infer_stream(host_u8[i % NB] for i in range(n)):
                pass
            torch.cuda.synchronize()
        e2e_u8(3)
        barrier()
        t0 = time.perf_counter()
        e0.record()
        e2e_u8(args.steps)
        e1.record()
        barrier()
        ms_u8 = reduce_max_ms(max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1000.0))
        e2e_uint8 = {"value": world * B * args.steps / (ms_u8 / 1000.0), "unit": UNIT, "ms_per_step": ms_u8 / args.steps,
                     "h2d_bytes_per_step": host_u8[0].numel(), "input": "uint8 [B,H,W,3] pinned host images, ToTensor + Normalize on the GPU (gwd_images_to_batch)"}
        h2d = host[0].numel() * host[0].element_size()
        d2h = sum(t.numel() * t.element_size() for t in out_host)

        # -------------------------------------------------------- roofline of the dominant kernel (instrumented pass)
        roof = None
        if rank == 0:
            ops.PROFILE = []
            # keep the GPU busy while the host enqueues the eager pass, so that the event pairs bracket back-to-back
            # kernels and not host launch gaps (the ~120 small DETR Linears would otherwise be charged ~10 us each)
            torch.cuda._sleep(int(0.15 * 1.9e9))
            plan.forward(resident[0])
            torch.cuda.synchronize()
            recs, ops.PROFILE = ops.PROFILE, None
            tot_ms = sum(a.elapsed_time(b) for a, b, _, _ in recs)
            tot_flop = sum(f for _, _, f, _ in recs)
            big = max(recs, key=lambda r: r[2])
            peak, peak_src = measured_peak()
            ach = tot_flop / (tot_ms / 1000.0) / 1e12
            traffic, traffic_src = None, None
            try:      # DRAM bytes of the same launches from the committed ncu pass (profiles/README.md); bench.py cannot run ncu
                with open(os.path.join(ROOT, "profiles", "r1_tapgemm_dram.json")) as f:
                    tj = json.load(f)
                traffic = tj["dram_bytes_per_launch"]
                traffic_src = "profiles/r1_tapgemm_dram.json: mean dram__bytes_read+write per launch over the %d launches of a step" % tj["launches"]
            except (OSError, KeyError, ValueError):
                pass
            roof = {"bound": "tensor", "kernel": "gwd_tapgemm_kernel (tcgen05 implicit GEMM, all %d launches of a step)" % len(recs),
                    "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": traffic,
                    "traffic_source": traffic_src, "algorithmic_flop_per_launch": tot_flop / len(recs), "peak_source": peak_src,
                    "flop_per_step": tot_flop, "kernel_ms_per_step": tot_ms, "kernel_share_of_step": tot_ms / (ms / args.steps),
                    "largest_launch": {"desc": big[3], "tflops": big[2] / (big[0].elapsed_time(big[1]) / 1000.0) / 1e12}}

    # ------------------------------------------------------------ line-branch training step (extra key, not the headline)
    train_line = None
    if not args.no_train:
        train_line = run_train(args, net, resident, rank, world, dev, barrier, reduce_max_ms)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # ------------------------------------------------------------ dense-branch training step (extra key, rank 0, one GPU's batch)
    # (everything behind the 1/32 stage: tools/bench_train_branch.py) measured in a child process: its 13 GB of activations
    # and any fault stay out of the headline measurement
    train_tail = None
    if not args.no_train:
        try:
            import subprocess
            import tempfile
            with tempfile.TemporaryDirectory() as td_:
                out = os.path.join(td_, "tail.json")
                env = dict(os.environ, CUDA_VISIBLE_DEVICES=str(dev.index if os.environ.get("CUDA_VISIBLE_DEVICES") is None else
                                                                os.environ["CUDA_VISIBLE_DEVICES"].split(",")[dev.index]))
                for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT"):
                    env.pop(k, None)
                subprocess.run([sys.executable, os.path.join(ROOT, "tools", "bench_train_branch.py"), "--batch", str(B), "--steps",
                                str(max(5, min(args.steps, 10))), "--json", out], check=True, timeout=600, env=env,
                               stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
                with open(out) as f:
                    train_tail = json.load(f)
        except Exception as e:  # noqa: BLE001  (an extra key must not take the headline line down)
            train_tail = {"value": None, "error": repr(e)[:300]}
    cpu = None
    eager = None
    if world == 1 and not args.no_cpu_baseline:
        # context only (not the reference arm): the same oracle restatement run as eager fp32 PyTorch ON THE GPU, i.e.
        # what the reference's op-by-op structure (library kernels, per-image Python loops, host syncs) costs on a B200
        try:
            from helpers import oracle
            sd_gpu = {k: v.to(dev) for k, v in synth_weights().items()}
            with torch.no_grad():
                oracle.forward(sd_gpu, resident[0])
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for i in range(3):
                    oracle.forward(sd_gpu, resident[i % NB])
                torch.cuda.synchronize()
            eager = {"value": 3 * B / (time.perf_counter() - t0), "unit": UNIT,
                     "what": "oracle/gwdepth_oracle.py (port of the reference forward) as eager fp32 PyTorch on the same GPU, batch 16"}
            del sd_gpu
        except Exception as e:  # noqa: BLE001
            eager = {"value": None, "error": repr(e)[:200]}
        rate, n = cpu_oracle_rate()
        cpu = {"value": rate, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
               "sample": "%d forwards of 1x3x480x640 (1/16 of a step) through oracle/gwdepth_oracle.py" % n}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch": world * B, "image": [H, W], "parallelism": "dp%d (replicas, no data-path collective)" % world,
                       "l2": "4 rotating input batches (236 MB > L2); per-step activations are several GB",
                       "cuda_graph": bool(net.use_cuda_graph)},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / args.steps,
                    "pipeline": "model.infer_stream: 3 streams, double-buffered H2D / forward / D2H"},
            "gpu_launches": launches, "clocks": sampler.summary(), "roofline": roof, "cpu_baseline": cpu,
            "gpu_eager_port": eager, "e2e_uint8_inputs": e2e_uint8, "train_line_branch": train_line, "train_dense_branch": train_tail}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
