"""clip + AdamW step over ALL trainable parameters of the model (the groups of src/main_glassrgbd.py:59-66):
optim.FlatAdamW (gwd_sumsq + gwd_adamw_step on flat buffers) vs clip_grad_norm_ + torch.optim.AdamW (foreach / fused)."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gwdepth_b200  # noqa: F401,E402
from gwdepth_b200 import model as M, optim  # noqa: E402


def groups(m):
    return [{"params": [p for n, p in m.named_parameters() if "backbone" not in n and p.requires_grad]},
            {"params": [p for n, p in m.named_parameters() if "backbone" in n and p.requires_grad], "lr": 1e-5}]


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t = time.perf_counter()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, (time.perf_counter() - t) * 1000 / n


net, _, _ = M.build_model(M.default_args(device="cuda"))
net.cuda()
ref, _, _ = M.build_model(M.default_args(device="cuda"))
ref.cuda()
nparam = sum(p.numel() for g in groups(net) for p in g["params"])
ntens = sum(len(g["params"]) for g in groups(net))
opt = optim.FlatAdamW(groups(net), lr=1e-4, weight_decay=1e-4, max_norm=0.1)
for g in opt.groups:
    g["G"].normal_(std=1e-3)
for p in ref.parameters():
    if p.requires_grad:
        p.grad = torch.randn_like(p) * 1e-3
res = {"flat": timed(opt.step)}
for name, kw in (("torch_foreach", dict(foreach=True)), ("torch_fused", dict(fused=True))):
    ro = torch.optim.AdamW(groups(ref), lr=1e-4, weight_decay=1e-4, **kw)

    def step():
        torch.nn.utils.clip_grad_norm_([p for p in ref.parameters() if p.requires_grad], 0.1)
        ro.step()
    res[name] = timed(step)
bytes_ = nparam * 30       # p, g, m, v read (16 B) + p, m, v written (12 B) + the norm pass (4 B) - mirror off
print("%d tensors, %.1f M parameters" % (ntens, nparam / 1e6))
for k, (dev_ms, host_ms) in res.items():
    print("%-14s device %.3f ms, host %.3f ms per step%s" % (k, dev_ms, host_ms, ("  (%.0f GB/s of %d MB algorithmic)" % (bytes_ / dev_ms / 1e6, bytes_ / 1e6)) if k == "flat" else ""))
