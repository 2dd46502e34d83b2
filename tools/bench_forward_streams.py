"""inference forward throughput with K graph instances replayed on K alternating streams (batch 16 x 480x640)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from helpers import synth, synth_weights
import gwdepth_b200  # noqa: F401
from gwdepth_b200 import model as M
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
net, _, _ = M.build_model(M.default_args(device="cuda"))
net.load_state_dict(synth_weights()); net.cuda().eval()
plan = net.plan()
xs = [synth.synth_batch(B, 480, 640, seed=200 + i)[0].cuda() for i in range(4)]
for K in (1, 2, 3):
    streams = [torch.cuda.Stream() for _ in range(K)]
    with torch.no_grad():
        for i in range(2 * K):
            with torch.cuda.stream(streams[i % K]):
                plan.forward_graphed(xs[i % 4], slot=i % K)
        torch.cuda.synchronize()
        steps = 30
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in streams:
            s.wait_event(e0)
        for i in range(steps):
            with torch.cuda.stream(streams[i % K]):
                plan.forward_graphed(xs[i % 4], slot=i % K)
        for s in streams:
            torch.cuda.current_stream().wait_stream(s)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print("K=%d streams: %.3f ms per batch of %d -> %.1f images/s" % (K, ms, B, B / ms * 1000), flush=True)
