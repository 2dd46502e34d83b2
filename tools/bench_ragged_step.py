"""training step on a RAGGED batch (8 images, padded size 480x640, half of them 448x576 with a padding mask): CUDA-graph replay vs kernel
by kernel.  usage: python tools/bench_ragged_step.py"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import gwdepth_b200  # noqa: F401,E402
from gwdepth_b200 import model as M, synth  # noqa: E402
from gwdepth_b200.train_model import Trainer  # noqa: E402
import json  # noqa: E402

spec = [tuple(k) for k in json.load(open(os.path.join(ROOT, "tests", "golden", "state_dict_spec.json")))["keys"]]
sd = synth.add_structural_buffers(synth.synth_state_dict(spec, seed=0), spec)
B, H, W = 8, 480, 640
images, targets, depth_gt, seg_gt = synth.synth_batch(B, H, W, seed=1)
mask = torch.zeros(B, H, W, dtype=torch.bool)
mask[::2, 448:, :] = True
mask[::2, :, 576:] = True
images = images.masked_fill(mask[:, None], 0.0).cuda()
depth_gt = depth_gt.masked_fill(mask[:, None], 0.0).cuda()
seg_gt = seg_gt.masked_fill(mask[:, None], 0).cuda()
mask = mask.cuda()
tg = [{k: v.cuda() for k, v in t.items()} for t in targets]
_, crit, _ = M.build_model(M.default_args(device="cuda", dropout=0.0))
criterion = crit[0].cuda()
for graph in (True, False):
    tr = Trainer(sd)
    tr.use_cuda_graph = graph
    for _ in range(3):
        tr.train_step(images, tg, depth_gt, seg_gt, criterion, mask=mask)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 10
    for _ in range(n):
        total, _ = tr.train_step(images, tg, depth_gt, seg_gt, criterion, mask=mask)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / n * 1000
    print("ragged batch, %s: %.2f ms per step (%.1f images/s), loss %.4f" % ("CUDA graphs" if graph else "kernel by kernel", ms, B / ms * 1000, float(total)))
    del tr
    torch.cuda.empty_cache()
