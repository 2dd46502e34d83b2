"""one eager WHOLE-MODEL training step (train_model.Trainer, batch B x 480x640) after one warm-up step, for
    ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "train_step/" ...
(the step of interest is the NVTX range `train_step`)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import synth, synth_weights  # noqa: E402
import gwdepth_b200  # noqa: F401,E402
from gwdepth_b200 import model as M  # noqa: E402

B, H, W = int(os.environ.get("B", 8)), int(os.environ.get("H", 480)), int(os.environ.get("W", 640))
net, crit, _ = M.build_model(M.default_args(device="cuda", dropout=0.0))
net.load_state_dict(synth_weights())
net.cuda()
criterion = crit[0].cuda()
tr = net.trainer()
im, tg, dg, sg = synth.synth_batch(B, H, W, seed=1)
im, dg, sg = im.cuda(), dg.cuda(), sg.cuda()
tg = [{k: v.cuda() for k, v in t.items()} for t in tg]
for _ in range(int(os.environ.get("WARM", 1))):
    tr.train_step(im, tg, dg, sg, criterion)
torch.cuda.synchronize()
torch.cuda.nvtx.range_push("train_step")
tr.train_step(im, tg, dg, sg, criterion)
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
