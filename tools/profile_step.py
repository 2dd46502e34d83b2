"""One forward step bracketed by cudaProfilerStart/Stop, for `ncu --profile-from-start off` (see profiles/README.md).
usage: python tools/profile_step.py [batch] [H] [W]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
from helpers import synth, synth_weights  # noqa: E402
import gwdepth_b200  # noqa: F401,E402
from gwdepth_b200 import model as M  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
H = int(sys.argv[2]) if len(sys.argv) > 2 else 480
W = int(sys.argv[3]) if len(sys.argv) > 3 else 640
net, _, _ = M.build_model(M.default_args(device="cuda"))
net.load_state_dict(synth_weights())
net.cuda().eval()
x = synth.synth_batch(B, H, W, seed=100)[0].cuda()
with torch.no_grad():
    for _ in range(2):
        net(x)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    net(x)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
print("profiled one forward of", (B, 3, H, W))
