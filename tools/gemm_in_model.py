"""Per-launch CUDA-event timing of every gwd_conv_gemm launch inside one model forward."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["GWD_CUDA_GRAPH"] = "0"      # eager launches: the per-launch events need them
import torch  # noqa: E402
from helpers import synth, synth_weights  # noqa: E402
import gwdepth_b200  # noqa: F401,E402
from gwdepth_b200 import model as M, ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
net, _, _ = M.build_model(M.default_args(device="cuda"))
net.load_state_dict(synth_weights())
net.cuda().eval()
x = synth.synth_batch(B, 480, 640, seed=100)[0].cuda()
with torch.no_grad():
    for _ in range(2):
        net(x)
    torch.cuda.synchronize()
    ops.PROFILE = []
    net(x)
    torch.cuda.synchronize()
recs, ops.PROFILE = ops.PROFILE, None
rows = [(a.elapsed_time(b) * 1000, f, d) for a, b, f, d in recs]
print("total %.1f us, %.1f TFLOP/s" % (sum(r[0] for r in rows), sum(r[1] for r in rows) / sum(r[0] for r in rows) / 1e6))
agg = {}
for us, f, d in rows:
    a = agg.setdefault(d, [0, 0.0, 0.0])
    a[0] += 1; a[1] += us; a[2] += f
print("grouped by shape:")
for d, (n, us, f) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:60]:
    print("%8.1f us x%-3d %7.1f us each %7.1f TFLOP/s  %s" % (us, n, us / n, f / us / 1e6, d))
