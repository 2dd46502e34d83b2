"""one eager forward + backward + step of the line branch (batch 16, 480x640) for `ncu --metrics gpu__time_duration.sum`"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import synth_weights  # noqa: E402
import gwdepth_b200  # noqa: F401,E402
from gwdepth_b200 import model as M, train  # noqa: E402

B = int(os.environ.get("B", 16))
cfg = M.GlassRGBD(M.default_args(device="cuda", dropout=0.0)).cfg
lb = train.LineBranch(synth_weights(), cfg)
c5 = (torch.randn(B, 15, 20, 2048, device="cuda").relu() * 0.5).bfloat16()
lo, li = lb.forward(c5)
torch.cuda.synchronize()
torch.cuda.nvtx.range_push("train_step")
lo, li = lb.forward(c5)
lb.backward(torch.randn_like(lo) * 1e-3, torch.randn_like(li) * 1e-3)
lb.step()
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
