"""Times the line-branch training step piece by piece (CUDA events on the launching stream), batch 16 x 480x640.
    python tools/bench_train.py [--batch 16] [--steps 10]"""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--steps", type=int, default=10)
    a = ap.parse_args()
    from helpers import synth, synth_weights
    import gwdepth_b200  # noqa: F401
    from gwdepth_b200 import capi, model as M, train
    net, crit, _ = M.build_model(M.default_args(device="cuda", dropout=0.0))
    net.load_state_dict(synth_weights())
    net.cuda().eval()
    images, targets, _, _ = synth.synth_batch(a.batch, 480, 640, seed=5)
    images = images.cuda().float().contiguous()
    targets = [{k: v.cuda() for k, v in t.items()} for t in targets]
    with torch.no_grad():
        c5 = net.plan().backbone(images)[3].contiguous()
    lb = train.LineBranch(synth_weights(), net.cfg)
    criterion = crit[0]

    def timed(fn, n=a.steps, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n, (time.perf_counter() - t0) * 1000 / n

    fwd = timed(lambda: lb.forward(c5))
    logits, lines = lb.forward(c5)
    dlo, dli = torch.randn_like(logits) * 1e-3, torch.randn_like(lines) * 1e-3

    def fb():
        lb.forward(c5)
        lb.backward(dlo, dli)
    fwd_bwd = timed(fb)
    opt = timed(lb.step)
    lb.use_cuda_graph = False
    eager_full = timed(lambda: lb.train_step(c5, targets, criterion))
    lb.use_cuda_graph = True
    lg = timed(lambda: lb.loss_and_grads(c5, targets, criterion))
    st = lb._captured(c5)
    gr = timed(lambda: (st["fwd"].replay(), st["bwd"].replay()))
    lo_, li_ = st["logits"].detach().clone(), st["lines"].detach().clone()
    cr = timed(lambda: criterion.forward_stacked(lo_, li_, targets))
    mt = timed(lambda: criterion.matcher.forward_stacked(lo_, li_, targets))
    print("graph replays fwd+bwd %.2f ms; stacked criterion %.2f ms host (of which matching %.2f ms)" % (gr[0], cr[1], mt[1]))
    print("eager train_step %.2f ms; graphed loss_and_grads %.2f ms (host %.2f)" % (eager_full[0], lg[0], lg[1]))
    capi.reset_launch_count()
    fb()
    lb.step()
    torch.cuda.synchronize()
    launches = capi.launch_count()
    full = timed(lambda: lb.train_step(c5, targets, criterion))
    print("batch %d: forward %.2f ms | forward+backward %.2f ms (host %.2f) | clip+AdamW+transposes %.2f ms | "
          "train_step incl. criterion + 6 Hungarian matchings %.2f ms (host %.2f) | %d launches per step | %.0f images/s"
          % (a.batch, fwd[0], fwd_bwd[0], fwd_bwd[1], opt[0], full[0], full[1], launches, a.batch / (max(full) / 1000)))


if __name__ == "__main__":
    main()
