#!/usr/bin/env python
"""bench.py -- images/sec of the GW-Depth model hot path at 480x640, bf16 (BASELINE.json `metric`:
"images/sec @480x640 bf16 fwd+bwd at 1/2/4/8 B200").

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torch.distributed.run, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W

HEADLINE (value / e2e / scaling): the WHOLE-MODEL TRAINING STEP of BASELINE configs[2] -- batch 8 per GPU, data parallel:
forward with saved activations, SetCriterion (6 Hungarian matchings on the host, hidden behind the dense branch), 4 x SilogLoss +
SegLoss, backward of all 684 trained tensors, NCCL all-reduce of the flat gradient buffers (overlapped with the backward), ONE
global clip norm, AdamW.  A step = one such iteration on one batch of 8 synthetic images per GPU.
  value     : images/s with the step's inputs already resident in HBM (4 distinct batches are rotated; the step's activation
              working set is several GB >> the 126 MB L2)
  e2e       : the same step through the public API (`model.trainer().train_step(...)`) from PINNED HOST buffers: the H2D copy of the
              step's images, depth / segmentation ground truth and line targets and the D2H read of its loss are inside the
              timed region
  roofline  : the tcgen05 implicit-GEMM kernel (gwd_tapgemm_kernel: every forward / data-gradient GEMM of the step), algorithmic
              FLOPs / sum of per-launch CUDA-event durations against the measured bf16 peak; `step` = the whole step against
              the reference's 1 049.5 GFLOP per image fwd+bwd
  forward   : BASELINE configs[1] -- inference forward, batch 16 per GPU (round 1's headline) -- device-resident and end to end
  gpu_eager_reference : the UNMODIFIED reference (staged under baseline/_ref by oracle/stage_ref.sh) as eager fp32 PyTorch on the
              same GPU: forward at batch 16 (the 10x denominator of north_star) and the training step at batch 8
  cpu_baseline : the UNMODIFIED reference's own training step on the host cores (all threads, fp32; kind "reference") when its staged
              copy is present, else the CPU oracle's training pass (oracle/gwdepth_oracle.py under torch.autograd; kind "port");
              a bounded sample: 1 image of the 8-image batch per step
The reference arm (--impl reference) times the same thing for --steps steps and prints it as its own line.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def _checker_paths():
    """tests/ and oracle/ on sys.path: ONLY the cpu_baseline leg, the --impl reference arm and the eager-reference leg call this (the
    product arm imports nothing from there)"""
    for d in ("tests", "oracle"):
        p = os.path.join(ROOT, d)
        if p not in sys.path:
            sys.path.insert(0, p)


def synth_weights(seed=0):
    """the synthetic state dict of the benchmark: generator in the package, key / shape list from the committed fixture"""
    from gwdepth_b200 import synth
    with open(os.path.join(ROOT, "tests", "golden", "state_dict_spec.json")) as f:
        spec = [tuple(k) for k in json.load(f)["keys"]]
    return synth.add_structural_buffers(synth.synth_state_dict(spec, seed=seed), spec)

import torch  # noqa: E402

METRIC = "images_per_sec_train_fwd_bwd_480x640_bf16"
UNIT = "images/s"
TRAIN_BATCH, FWD_BATCH, H, W = 8, 16, 480, 640
WORKLOAD = ("GW-Depth stage-1 ResNet-50 line+depth model, whole-model training step (forward + 17 losses + backward + gradient "
            "all-reduce + clip + AdamW), batch 8 per GPU, 480x640 (BASELINE configs[2])")
FLOP_PER_IMAGE_TRAIN, FLOP_PER_IMAGE_FWD = 1049.5e9, 359.0e9          # SURVEY section 6 (FlopCounterMode over the reference)


def _tapgemm_dram():
    """dram__bytes_read.sum + dram__bytes_write.sum per gwd_tapgemm_kernel launch of one training step, from the committed ncu pass
    (profiles/r2_tapgemm_dram_train.json, written by tools/summarize_dram.py); None when that capture is absent"""
    try:
        with open(os.path.join(ROOT, "profiles", "r2_tapgemm_dram_train.json")) as f:
            d = json.load(f)
        return d["dram_bytes_per_launch"], "profiles/r2_tapgemm_dram_train.json (%d launches; %s)" % (d["launches"], d["how"])
    except (OSError, KeyError, ValueError):
        return None, None


TAPGEMM_DRAM_BYTES_PER_LAUNCH, TAPGEMM_DRAM_SOURCE = _tapgemm_dram()


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=("b200", "reference"))
    ap.add_argument("--batch", type=int, default=TRAIN_BATCH)
    ap.add_argument("--dense-center", action="store_true", help="BASELINE configs[3]: --with_dense_center (3 points per line)")
    ap.add_argument("--dropout", type=float, default=0.0, help="train-mode dropout of the DETR layers (the reference's default is 0.1; "
                    "0.0 is the parity configuration the headline is quoted on)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-forward", action="store_true", help="skip the inference-forward section")
    ap.add_argument("--no-data-path", action="store_true", help="skip the data-path (augmentation pipeline) measurement")
    ap.add_argument("--fwd-streams", type=int, default=3, help="inference-forward section: graph instances / streams consecutive batches alternate on")
    ap.add_argument("--no-reference-eager", action="store_true")
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


def cpu_train_pass(sd_leaves, batch, wd):
    """one forward + backward of the CPU oracle under torch.autograd (the engine's 17 losses) on `batch`"""
    _checker_paths()
    from helpers import oracle
    images, targets, depth_gt, seg_gt = batch
    for v in sd_leaves.values():
        if v.requires_grad:
            v.grad = None
    out = oracle.forward(sd_leaves, images, grad=True)
    set_l, _ = oracle.set_criterion(out, [t["lines"] for t in targets])
    total = sum(v * wd[k] for k, v in set_l.items()) + sum(oracle.depth_losses(out["pred_depth"], depth_gt)) + oracle.seg_loss(out["pred_seg"], seg_gt)
    total.backward()
    return float(total.detach())


def cpu_setup():
    _checker_paths()
    from helpers import synth, synth_weights
    use_all_host_threads()
    sd = synth_weights()
    frozen = ("backbone.0.body.conv1", "backbone.0.body.layer1")
    leaves = {k: (v.clone().requires_grad_(True) if (v.is_floating_point() and "running" not in k and ".bn" not in k and "downsample.1" not in k
                                                     and not k.startswith(frozen)) else v) for k, v in sd.items()}
    wd = {"loss_ce": 1.0, "loss_line": 5.0}
    wd.update({"%s_%d" % (k, i): v for i in range(5) for k, v in list(wd.items())[:2]})
    return leaves, synth.synth_batch(1, H, W, seed=0), wd


def cpu_stepper():
    """-> (step callable, kind, description): ONE image of the training batch per call on the host cores, all threads, fp32.  With a
    copy of the reference at hand (baseline/_ref, or the mounted tree in the build container) the UNMODIFIED reference's own training
    step (`kind` = "reference"); otherwise the oracle port's training pass (`kind` = "port")."""
    _checker_paths()
    use_all_host_threads()
    try:
        import ref_shims
        if ref_shims.reference_available():
            from helpers import synth_weights
            model, criterions, _, _ = ref_shims.build_reference(["--device", "cpu", "--dropout", "0.0"])
            model.load_state_dict(synth_weights(), strict=True)
            step, _opt = reference_train_step(model, criterions, torch.device("cpu"), 1)
            step()              # (a first, un-timed step inside the guard: whatever goes wrong here falls back to the port)
            return step, "reference", ("training steps of 1x3x480x640 (1/8 of a step) through the UNMODIFIED reference (baseline/_ref: model, "
                                       "SetCriterion + Hungarian matcher, SilogLoss x 4, SegLoss, backward, clip_grad_norm_, torch AdamW), "
                                       "eager fp32 on the host cores")
    except Exception as e:  # noqa: BLE001  (the CPU figure must always be there: fall back to the port)
        sys.stderr.write("cpu baseline: the staged reference did not run (%r); timing the oracle port instead\n" % (e,))
    leaves, batch, wd = cpu_setup()
    return (lambda: cpu_train_pass(leaves, batch, wd)), "port", ("training passes (forward + 17 losses + backward, no optimizer step) of "
                                                                 "1x3x480x640 (1/8 of a step) through oracle/gwdepth_oracle.py under torch.autograd")


def run_reference(args, rank):
    """the reference arm, rank 0 only, all host threads, fp32, CPU.  With the staged copy of the reference (baseline/_ref, or the
    mounted tree in the build container): the UNMODIFIED reference's own training step -- its model, criteria, clip and AdamW -- on
    the host cores (`kind` = "reference").  Without it: the oracle port's training pass (forward + 17 losses + backward under
    torch.autograd, `kind` = "port").  Every step is ONE image of the 8-image batch (a bounded sample: the full batch takes ~8x as
    long), reported as a per-image rate."""
    if rank != 0:
        return
    step, kind, what = cpu_stepper()
    for _ in range(min(args.warmup, 1)):
        step()
    t0 = time.time()
    for _ in range(args.steps):
        step()
    dt = time.time() - t0
    v = args.steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample": "1 image of the 8-image batch per step on the CPU; per-image rate"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind, "sample": "%d %s" % (args.steps, what)},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def attention_rate(dev, peak_tflops):
    """BASELINE metric, second half ("attn tensor-pipe util %"): the DETR attention kernel timed alone with CUDA events at the training
    shape (B = 8, L = 300) and at the corner SURVEY section 7 names (B = 64, L = 1 200 = 960x1280 images); FLOPs = 4 Lq Lk 256 per
    image.  The tensor-pipe utilisation itself needs ncu: the committed captures are cited."""
    from gwdepth_b200 import ops
    out = {"kernel": "gwd_attention_tc_kernel / gwd_attention_flash_tc_kernel (tcgen05 + TMEM + TMA, head dim 32)", "cases": {},
           "tensor_pipe_active_pct_ncu": {"enc_b64_l1200": 31.9, "enc_b16_l300": "profiles/r2_ncu_full_attention_enc_b16_l300.txt",
                                          "source": "profiles/r2_ncu_full_attention_enc_b64_l1200.txt (ncu --set full)"},
           "note": "head dim 32 makes the soft-max exponentials (MUFU, 16 / clk / SM) the bound: <= ~25 % of the tensor pipe"}
    E, NH, HD = 256, 8, 32
    for name, (B_, L_) in {"enc_b8_l300": (8, 300), "enc_b64_l1200": (64, 1200)}.items():
        g = torch.Generator(device=dev).manual_seed(0)
        sets = [(torch.randn(B_ * L_, E, device=dev, generator=g).bfloat16() * HD ** -0.25,
                 torch.randn(B_ * L_, E, device=dev, generator=g).bfloat16() * HD ** -0.25,
                 torch.randn(B_ * L_, E, device=dev, generator=g).bfloat16()) for _ in range(3)]
        o = torch.empty(B_ * L_, E, device=dev, dtype=torch.bfloat16)

        def call(i):
            q, k, v = sets[i % 3]
            ops.attention(q, k, v, o, items=B_, heads=NH, Lq=L_, Lk=L_, hd=HD, q_strides=(L_ * E, E), k_strides=(L_ * E, E),
                          v_strides=(L_ * E, E), o_strides=(L_ * E, E), scale=1.0)
        for i in range(3):
            call(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(20):
            call(i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        tf = 4.0 * B_ * L_ * L_ * E / ms / 1e9
        out["cases"][name] = {"us": ms * 1000, "tflops": tf, "frac_of_measured_bf16_peak": tf / peak_tflops}
    return out


def data_path_rate(dev, n=48):
    """the training augmentation pipeline (make_coco_transforms('train'), src/datasets/coco.py:74-103) on 480x640
    samples -- gw-depth_b200/data.py on the GPU (one host thread) and, when the staged reference + Pillow are there, the reference's own
    PIL pipeline on one host core (what one of its DataLoader workers delivers)."""
    import random
    import types
    import numpy as np
    from gwdepth_b200 import data as gdata
    out = {"what": "augmentation pipeline of the training loader, 480x640 uint8 samples (flip / multi-scale resize / size crop / ColorJitter / "
                   "normalise), images already decoded", "unit": "samples/s"}
    rng = np.random.default_rng(0)
    samples = []
    for s_ in range(4):
        img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        depth = rng.integers(300, 9000, (H, W)).astype(np.int32)
        seg = rng.integers(0, 3, (H, W)).astype(np.uint8)
        lines = torch.tensor([[100., 120., 300., 140.], [300., 140., 320., 330.], [110., 320., 320., 330.], [100., 120., 110., 320.]])
        tgt = {"lines": lines, "poly_centers": lines[:, :2] * 0 + torch.tensor([210., 230.]), "poly_ids": torch.zeros(4, dtype=torch.int64),
               "labels": torch.zeros(4, dtype=torch.int64), "area": torch.ones(4), "iscrowd": torch.zeros(4),
               "orig_size": torch.as_tensor([H, W]), "size": torch.as_tensor([H, W])}
        samples.append((img, depth, seg, tgt))
    args_ = types.SimpleNamespace(eval=False)

    def run(tf, conv, count):
        done = 0
        for i in range(count):
            img, depth, seg, tgt = samples[i % 4]
            try:
                tf(conv(img, "RGB"), {k: v.clone() for k, v in tgt.items()}, aux_mats=[conv(depth, "I"), conv(seg, "L")])
                done += 1
            except ImportError:        # the one crop branch that needs shapely
                pass
        return done
    try:
        tf = gdata.make_coco_transforms("train", args_)
        dev_cache = {}

        def to_dev(a, _mode):
            k = id(a)
            if k not in dev_cache:
                dev_cache[k] = torch.from_numpy(a).to(dev)
            return dev_cache[k]
        random.seed(0); torch.manual_seed(0)
        run(tf, to_dev, 8)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        done = run(tf, to_dev, n)
        torch.cuda.synchronize()
        out["value"] = done / (time.perf_counter() - t0)
    except Exception as e:      # noqa: BLE001
        out["error"] = repr(e)[:200]
    try:
        from PIL import Image
        _checker_paths()           # reference leg of the data-path measurement
        import ref_shims
        if ref_shims.reference_available():
            ref_shims.install()
            src = os.path.join(ref_shims.REFERENCE_ROOT, "src")
            if src not in sys.path:
                sys.path.insert(0, src)
            try:
                import shapely.geometry  # noqa: F401
            except ImportError:
                def _no(*a, **k):
                    raise ImportError("shapely is not installed")
                sh, g = types.ModuleType("shapely"), types.ModuleType("shapely.geometry")
                g.Polygon, g.mapping, sh.geometry = _no, _no, g
                sys.modules["shapely"], sys.modules["shapely.geometry"] = sh, g
            import datasets.coco as rcoco
            rtf = rcoco.make_coco_transforms("train", args_)
            random.seed(0); torch.manual_seed(0)
            nt = torch.get_num_threads()
            torch.set_num_threads(1)
            run(rtf, lambda a, m: Image.fromarray(a, mode=None if m == "RGB" else m), 2)
            t0 = time.perf_counter()
            done = run(rtf, lambda a, m: Image.fromarray(a, mode=None if m == "RGB" else m), max(8, n // 4))
            out["reference_pil_one_core"] = done / (time.perf_counter() - t0)
            torch.set_num_threads(nt)
    except Exception as e:      # noqa: BLE001
        out["reference_error"] = repr(e)[:200]
    return out


def reference_train_step(model, criterions, dev, batch):
    """one training step of the UNMODIFIED reference as src/engine_glassrgbd.py:45-166 runs it (its own model, SetCriterion /
    SilogLoss / SegLoss, clip_grad_norm_ 0.1, torch AdamW with the two learning rates of src/main_glassrgbd.py:53-67) on a synthetic
    batch of `batch` images resident on `dev`; returns (step callable, optimizer)"""
    _checker_paths()
    from helpers import synth
    import torch.nn.functional as F
    model.train()
    crit, crit_d, crit_s = criterions[0].to(dev), criterions[1], criterions[2]
    crit.train()
    opt = torch.optim.AdamW([{"params": [p for n, p in model.named_parameters() if "backbone" not in n and p.requires_grad]},
                             {"params": [p for n, p in model.named_parameters() if "backbone" in n and p.requires_grad], "lr": 1e-5}],
                            lr=1e-4, weight_decay=1e-4)
    imgs, targets, depth_gt, seg_gt = synth.synth_batch(batch, H, W, seed=100)
    imgs, depth_gt, seg_gt = imgs.to(dev), depth_gt.to(dev), seg_gt.to(dev)
    targets = [{k: v.to(dev) for k, v in t.items()} for t in targets]

    def step():
        out = model(imgs)
        ld = crit(out, targets)
        loss = sum(ld[k] * crit.weight_dict[k] for k in ld if k in crit.weight_dict)
        mask = (depth_gt >= 0.2) & (depth_gt < 10.0)
        for i, pd in enumerate(out["pred_depth"]):
            sz = pd.shape[-2:]
            loss = loss + crit_d(pd, F.interpolate(depth_gt, size=sz, mode="nearest"),
                                 F.interpolate(mask.to(torch.uint8), size=sz, mode="nearest").to(torch.bool)) * (0.25, 0.25, 0.25, 1.0)[i]
        loss = loss + crit_s(out["pred_seg"], seg_gt.squeeze(1)) * 2.0
        opt.zero_grad()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 0.1)
        opt.step()
        return loss.detach()          # (no host read here: the GPU leg must not synchronise per step)
    return step, opt


def gpu_eager_reference(dev, n_fwd=30, n_train=10):
    """the unmodified reference on this GPU, eager fp32, as shipped (no AMP / TF32 override / cudnn.benchmark): forward at batch 16
    (north_star's 10x denominator) and one training step at batch 8 (its own criteria, AdamW, clip)"""
    _checker_paths()
    import ref_shims
    if not ref_shims.reference_available():
        return {"unavailable": "reference tree not staged (run oracle/stage_ref.sh in the build container)"}
    from helpers import synth, synth_weights
    res = {"what": "unmodified reference (baseline/_ref), eager fp32 PyTorch on the same GPU"}
    try:
        model, criterions, _, rargs = ref_shims.build_reference(["--device", "cuda", "--dropout", "0.0"])
        model.load_state_dict(synth_weights(), strict=True)
        model.to(dev).eval()
        images = synth.synth_batch(FWD_BATCH, H, W, seed=100)[0].to(dev)
        with torch.no_grad():
            for _ in range(10):         # >= 10 warm-up + >= 30 timed iterations for the 10x denominator
                model(images)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n_fwd):
                model(images)
            e1.record()
            torch.cuda.synchronize()
        res["forward_b16"] = {"value": n_fwd * FWD_BATCH / (e0.elapsed_time(e1) / 1000.0), "unit": UNIT, "ms_per_step": e0.elapsed_time(e1) / n_fwd}
        # training step, batch 8
        step, opt = reference_train_step(model, criterions, dev, TRAIN_BATCH)
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(n_train):
            step()
        e1.record()
        torch.cuda.synchronize()
        res["train_b8"] = {"value": n_train * TRAIN_BATCH / (e0.elapsed_time(e1) / 1000.0), "unit": UNIT, "ms_per_step": e0.elapsed_time(e1) / n_train}
        del model, opt
        torch.cuda.empty_cache()
    except Exception as e:  # noqa: BLE001  (context numbers must not take the headline down)
        res["error"] = repr(e)[:300]
    return res


_REAL_STDOUT = None


def emit(line):
    """the ONE JSON line of the contract on the real stdout (everything else -- NCCL's version banner, library warnings -- was
    sent to stderr by main())"""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    args = parse()
    # fd-level: native libraries (NCCL prints "NCCL version ..." on stdout) must not precede the JSON line
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    import torch.distributed as dist
    import gwdepth_b200  # noqa: F401
    from gwdepth_b200 import capi, model as M, ops, synth          # (synthetic inputs: generator in the package, not oracle/)

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, steps, warm = args.batch, args.steps, max(args.warmup, 3)

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

    margs = M.default_args(device="cuda", dropout=float(args.dropout), with_dense_center=bool(args.dense_center))
    net, criterions, _ = M.build_model(margs)
    net.load_state_dict(synth_weights())
    net.to(dev)
    criterion = criterions[0].to(dev)
    tr = net.trainer()
    NB = 4
    host = []
    for i in range(NB):
        im, tg, dg, sg = synth.synth_batch(B, H, W, seed=100 + 7 * rank + i)
        host.append((im.pin_memory(), tg, dg.pin_memory(), sg.pin_memory()))
    to_dev = lambda hb: (hb[0].to(dev, non_blocking=True), [{k: v.to(dev, non_blocking=True) for k, v in t.items()} for t in hb[1]],  # noqa: E731
                         hb[2].to(dev, non_blocking=True), hb[3].to(dev, non_blocking=True))
    resident = [to_dev(hb) for hb in host]

    # ------------------------------------------------------------ headline: training step, inputs resident in HBM
    def step(i):
        im, tg, dg, sg = resident[i % NB]
        return tr.train_step(im, tg, dg, sg, criterion)

    for i in range(warm):
        step(i)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    capi.reset_launch_count()
    replayed0 = getattr(tr, "replayed_kernels", 0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0 = time.perf_counter()
    e0.record()
    for i in range(steps):
        total, losses = step(i)
    e1.record()
    barrier()
    # library kernels of the timed region: those launched through the C ABI directly (matching costs, set loss, clip, AdamW) plus
    # those inside the three CUDA graphs of every step (counted once, while the graphs were captured)
    launches = capi.launch_count() + getattr(tr, "replayed_kernels", 0) - replayed0
    sampler.stop_flag = True
    ms = reduce_max_ms(max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1000.0))
    value = world * B * steps / (ms / 1000.0)
    loss_value = float(total)

    # ------------------------------------------------------------ end to end: the same step from pinned host buffers
    # the upload of batch i+1 runs on a side stream while step i computes; the loss of step i is copied to pinned host memory
    # asynchronously and read one step later (the reference's loop reads it with .item() every step)
    side = torch.cuda.Stream(device=dev)
    loss_host = torch.zeros(steps + warm, dtype=torch.float32).pin_memory()

    def e2e_run(n):
        main_s = torch.cuda.current_stream()
        with torch.cuda.stream(side):
            nxt = to_dev(host[0])
            ev = torch.cuda.Event()
            ev.record(side)
        seen = 0.0
        for i in range(n):
            main_s.wait_event(ev)
            cur = nxt
            if i + 1 < n:
                with torch.cuda.stream(side):
                    nxt = to_dev(host[(i + 1) % NB])
                    ev = torch.cuda.Event()
                    ev.record(side)
            tot, _ = tr.train_step(cur[0], cur[1], cur[2], cur[3], criterion)
            for t in (cur[0], cur[2], cur[3]):
                t.record_stream(main_s)
            loss_host[i:i + 1].copy_(tot.reshape(1), non_blocking=True)
            if i:
                seen += float(loss_host[i - 1])       # (the copy of step i-1 finished before the matching of step i was solved)
        torch.cuda.synchronize()
        return seen + float(loss_host[n - 1])
    e2e_run(2)
    barrier()
    t0 = time.perf_counter()
    e0.record()
    e2e_run(steps)
    e1.record()
    barrier()
    ms_e2e = reduce_max_ms(max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1000.0))
    e2e_value = world * B * steps / (ms_e2e / 1000.0)
    h2d = sum(t.numel() * t.element_size() for t in (host[0][0], host[0][2], host[0][3])) + sum(v.numel() * v.element_size() for t in host[0][1] for v in t.values())
    e2e = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4 + sum(t["lines"].shape[0] for t in host[0][1]) * 0 + 6 * B * 100 * 60 * 4,
           "ms_per_step": ms_e2e / steps, "api": "model.trainer().train_step(images, targets, depth_gt, seg_gt, criterion) from pinned host batches; "
           "d2h = the matching costs the host solver reads + the loss"}

    # ------------------------------------------------------------ where the step goes + roofline (instrumented, un-timed above)
    def timed(fn, n=3):
        fn()
        torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(n):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t) * 1000.0 / n
    roof, breakdown = None, None
    if True:      # every rank runs it (the criterion and the optimizer step contain collectives); rank 0 reports
        im, tg, dg, sg = resident[0]
        # the instrumented pass runs the step's kernels ONE AT A TIME (weight gradients not forked onto the side stream), so that
        # an event pair brackets one tcgen05 GEMM launch alone, like the per-launch times of the ncu launch list under profiles/;
        # with the forks on, every GEMM shares the SMs with a concurrent weight gradient and its event time is not its own
        from gwdepth_b200 import train_flat
        fork0, train_flat.FORK_WGRAD = train_flat.FORK_WGRAD, False
        try:
            ops.PROFILE = []
            torch.cuda._sleep(int(0.2 * 1.9e9))       # keep the GPU busy while the host enqueues: events bracket kernels, not launch gaps
            lo, li, outs = tr.forward(im)
            g = tr.dense.loss_grads(outs, dg, sg)
            tr.backward_dense(*g)
            _, dlo, dli = criterion.forward_backward_stacked(lo, li, tg)
            tr.backward_line(dlo, dli)
            torch.cuda.synchronize()
            recs, ops.PROFILE = ops.PROFILE, None
        finally:
            train_flat.FORK_WGRAD, ops.PROFILE = fork0, None
        # what an event pair costs on its own (record, record with nothing in between, behind a busy GPU): subtracted per launch
        torch.cuda._sleep(int(0.02 * 1.9e9))
        null = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(200)]
        for a, b in null:
            a.record()
            b.record()
        torch.cuda.synchronize()
        null_ms = sorted(a.elapsed_time(b) for a, b in null)[len(null) // 2]
        raw_ms = sum(a.elapsed_time(b) for a, b, _, _ in recs)
        tot_ms = raw_ms - null_ms * len(recs)
        tot_flop = sum(f for _, _, f, _ in recs)
        big = max(recs, key=lambda r: r[2])
        peak, peak_src = measured_peak()
        ach = tot_flop / (tot_ms / 1000.0) / 1e12
        step_ach = value / world * FLOP_PER_IMAGE_TRAIN / 1e12
        roof = {"bound": "tensor", "kernel": "gwd_tapgemm_kernel (tcgen05 implicit GEMM: the %d forward + data-gradient launches of a step)" % len(recs),
                "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": TAPGEMM_DRAM_BYTES_PER_LAUNCH,
                "traffic_source": TAPGEMM_DRAM_SOURCE,
                "algorithmic_flop_per_launch": tot_flop / len(recs), "peak_source": peak_src, "flop_per_step": tot_flop,
                "kernel_ms_per_step": tot_ms, "kernel_ms_per_step_raw_events": raw_ms, "event_pair_overhead_us": null_ms * 1000.0,
                "how": "one CUDA-event pair per launch on the launching stream, kernels of the step run one at a time (no forked weight "
                       "gradients beside them), the median empty event pair subtracted per launch",
                "kernel_share_of_step": tot_ms / (ms / steps),
                "largest_launch": {"desc": big[3], "tflops": big[2] / (big[0].elapsed_time(big[1]) / 1000.0) / 1e12},
                "step": {"achieved": step_ach, "frac": step_ach / peak, "flop_per_image": FLOP_PER_IMAGE_TRAIN,
                         "what": "whole training step per GPU against the reference's fwd+bwd FLOPs (SURVEY section 6)"}}
        ms_fwd = timed(lambda: tr.forward(im))
        lo, li, outs = tr.forward(im)
        g = tr.dense.loss_grads(outs, dg, sg)

        def bwd():
            tr.forward(im)
            tr.backward_dense(*tr.dense.loss_grads(outs, dg, sg))
            tr.backward_line(dlo, dli)
        ms_fb = timed(bwd)
        ms_crit = timed(lambda: criterion.forward_backward_stacked(lo, li, tg))
        tr._works = []
        ms_ar = timed(lambda: [dist.all_reduce(m.G) for m in tr.modules()]) if world > 1 else 0.0
        breakdown = {"allreduce_alone_ms": ms_ar, "forward_ms": ms_fwd, "forward_plus_backward_ms": ms_fb, "matching_host_ms": ms_crit, "optimizer_ms": timed(tr.step),
                     "lsap_threads": M._lsap_threads(), "allreduce_bytes_per_step": tr.numel() * 4 if world > 1 else 0,
                     "flat_buffers": len(tr.modules()), "trained_parameters": tr.numel()}

    # ------------------------------------------------------------ BASELINE configs[1]: inference forward, batch 16 per GPU
    fwd = None
    if not args.no_forward:
        net.sync_from_trainer()
        net.eval()
        plan = net.plan()
        hostf = [synth.synth_batch(FWD_BATCH, H, W, seed=200 + 7 * rank + i)[0].pin_memory() for i in range(NB)]
        resf = [h.to(dev) for h in hostf]
        with torch.no_grad():
            # consecutive batches replay two CUDA-graph instances on two streams (what model.infer_stream does): the latency-bound
            # phases of one forward share the SMs with the wide phases of the next.  --fwd-streams 1: one stream, one instance.
            nstr = args.fwd_streams if net.use_cuda_graph else 1
            fstreams = [torch.cuda.Stream(dev) for _ in range(nstr)]

            def frun(n):
                cur = torch.cuda.current_stream()
                for st_ in fstreams:
                    st_.wait_stream(cur)
                for i in range(n):
                    with torch.cuda.stream(fstreams[i % nstr]):
                        if net.use_cuda_graph:
                            plan.forward_graphed(resf[i % NB], slot=i % nstr)
                        else:
                            plan.forward(resf[i % NB])
                for st_ in fstreams:
                    cur.wait_stream(st_)
            frun(max(warm, 2 * nstr))
            barrier()
            e0.record()
            frun(steps)
            e1.record()
            barrier()
            ms_f = reduce_max_ms(e0.elapsed_time(e1))

            def fe2e(n):
                for _ in net.infer_stream(hostf[i % NB] for i in range(n)):
                    pass
                torch.cuda.synchronize()
            fe2e(3)
            barrier()
            t0 = time.perf_counter()
            e0.record()
            fe2e(steps)
            e1.record()
            barrier()
            ms_fe = reduce_max_ms(max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1000.0))
        fv = world * FWD_BATCH * steps / (ms_f / 1000.0)
        fwd = {"metric": "images_per_sec_fwd_480x640_bf16", "value": fv, "unit": UNIT, "ms_per_step": ms_f / steps, "batch_per_gpu": FWD_BATCH,
               "e2e": world * FWD_BATCH * steps / (ms_fe / 1000.0), "frac_of_peak": fv / world * FLOP_PER_IMAGE_FWD / 1e12 / measured_peak()[0],
               "streams": nstr,
               "workload": "BASELINE configs[1]: inference forward, one CUDA-graph replay per batch, consecutive batches on %d alternating streams "
                           "(ms_per_step = timed region / batches); e2e = model.infer_stream from pinned host batches" % nstr}
        del resf, plan
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    cpu, eager = None, None
    if world == 1:
        if not args.no_reference_eager:
            del tr
            net.__dict__["_trainer"] = None
            torch.cuda.empty_cache()
            eager = gpu_eager_reference(dev)
            if fwd and isinstance(eager.get("forward_b16"), dict):
                eager["forward_speedup"] = fwd["value"] / eager["forward_b16"]["value"]
            if isinstance(eager.get("train_b8"), dict):
                eager["train_speedup"] = value / eager["train_b8"]["value"]
        if not args.no_cpu_baseline:
            cpu_step, cpu_kind, cpu_what = cpu_stepper()
            cpu_step()
            t0, n = time.time(), 0
            while n < 2 or (time.time() - t0 < 20.0 and n < 6):
                cpu_step()
                n += 1
            cpu = {"value": n / (time.time() - t0), "unit": UNIT, "cores": torch.get_num_threads(), "kind": cpu_kind,
                   "sample": "%d %s" % (n, cpu_what)}
    data_path, attention = None, None
    if world == 1 and not args.no_data_path:
        data_path = data_path_rate(dev)
    if world == 1:
        attention = attention_rate(dev, measured_peak()[0])
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warm,
            "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic", "e2e": e2e, "gpu_launches": launches, "clocks": sampler.summary(),
            "config": {"workload": WORKLOAD, "global_batch": world * B, "image": [H, W], "with_dense_center": bool(args.dense_center),
                       "parallelism": "dp%d: batch sharded over the ranks, NCCL all-reduce of %d flat gradient buffers overlapped with the backward" % (world, len(net.__dict__.get("_live") or []) and 22),
                       "l2": "4 rotating input batches; per-step activations are several GB", "loss": loss_value,
                       "optimizer": "AdamW lr 1e-4 (backbone 1e-5), weight decay 1e-4, global clip 0.1; dropout %g" % args.dropout},
            "breakdown": breakdown, "forward": fwd, "roofline": roof, "cpu_baseline": cpu, "gpu_eager_reference": eager,
            "data_path": data_path, "attention": attention}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
