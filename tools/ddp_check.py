"""Data-parallel check of the whole-model training step (run under torchrun, one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/ddp_check.py
Every rank trains on ITS OWN batches for a few fused steps (train_model.Trainer: asynchronous all-reduce of the flat gradient
buffers, one global clip norm, AdamW).  Checks: the replicas' parameters stay BIT-IDENTICAL (the exchange covers every trained
tensor), the all-reduced gradient equals the mean of the ranks' local gradients, and the loss is finite.  Prints one JSON line."""
import json, os, sys
import torch
import torch.distributed as dist
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
from helpers import synth, synth_weights
import gwdepth_b200  # noqa: F401
from gwdepth_b200 import model as M

rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
B, H, W = 2, 224, 320
net, crit, _ = M.build_model(M.default_args(device="cuda", dropout=0.0))
net.load_state_dict(synth_weights())
net.to(dev)
criterion = crit[0].to(dev)
tr = net.trainer()
losses = []
for s in range(3):
    im, tg, dg, sg = synth.synth_batch(B, H, W, seed=10 * s + rank)
    tg = [{k: v.to(dev) for k, v in t.items()} for t in tg]
    if s == 0 and world > 1:       # local gradients first (no exchange), to compare with the all-reduced ones
        tr.exchange_grads = False
        lo, li, outs = tr.forward(im.to(dev))
        tr.backward_dense(*tr.dense.loss_grads(outs, dg.to(dev), sg.to(dev)))
        _, dlo, dli = criterion.forward_backward_stacked(lo, li, tg)
        tr.backward_line(dlo, dli)
        local_g = [m.G.clone() for m in tr.modules()]
        tr.exchange_grads = True
    total, _ = tr.train_step(im.to(dev), tg, dg.to(dev), sg.to(dev), criterion)
    if s == 0 and world > 1:
        worst = 0.0
        for m, g in zip(tr.modules(), local_g):
            dist.all_reduce(g)
            worst = max(worst, float((m.G - g).abs().max() / g.abs().max().clamp_min(1e-20)))
    losses.append(float(total))
ok_same = True
if world > 1:
    for m in tr.modules():
        ref = m.P.clone()
        dist.broadcast(ref, 0)
        ok_same &= bool(torch.equal(ref, m.P))
    flag = torch.tensor([int(ok_same)], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    ok_same = bool(flag.item())
if rank == 0:
    print(json.dumps({"world": world, "losses_rank0": losses, "replicas_bit_identical": ok_same,
                      "allreduced_vs_sum_of_local_max_rel": worst if world > 1 else None, "buffers": len(tr.modules()), "parameters": tr.numel()}))
if world > 1:
    dist.destroy_process_group()
