"""Times gwd_ref_diffuse (conv + instance-norm + GELU + residual) at the benchmark shape: B=16, 16 heads, P=441, R=40."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import gwdepth_b200  # noqa: F401
from gwdepth_b200 import ops

B, nh, P, R = 16, 16, 441, 40
g = torch.Generator(device="cuda").manual_seed(0)
a0 = torch.randn(B, nh, P, R, device="cuda", generator=g)
a1 = torch.empty_like(a0)
raw = torch.empty_like(a0)
stats = torch.empty(B * nh * 2, dtype=torch.float64, device="cuda")
w = (torch.randn(nh, nh, 3, 3, generator=torch.Generator().manual_seed(1)) * 0.1).contiguous()
b = torch.zeros(nh)
for _ in range(3):
    ops.ref_diffuse(a0, a1, w, b, raw, stats, B, nh, P, R)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    ops.ref_diffuse(a0, a1, w, b, raw, stats, B, nh, P, R)
e1.record()
torch.cuda.synchronize()
print("ref_diffuse (conv + norm): %.1f us" % (e0.elapsed_time(e1) / 20 * 1000))
