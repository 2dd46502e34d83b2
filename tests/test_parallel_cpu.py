"""CPU, world_size 2 over gloo: the host-side data-parallel logic (sharding, gradient all-reduce on a flat buffer, metric
and timing reductions) that the N > 1 paths of bench.py / train.LineBranch.step / tools/eval_sweep.py rely on."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import gwdepth_b200  # noqa: F401
from gwdepth_b200 import parallel


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rk, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rk), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rk, world_size=world)
    try:
        # 1. sharding: the 7 evaluation images split 4 + 3, every image owned exactly once
        mine = list(parallel.shard(7))
        # 2. flat gradient buffer: rank r holds (r + 1) * base; after the sum all-reduce and the 1 / world scale every rank
        #    has the mean gradient
        base = torch.arange(1000, dtype=torch.float32)
        G = base * (rk + 1)
        w = parallel.allreduce_sum_(G)
        # 3. evaluation metrics: per-image rows -> mean over all images of all ranks
        per_image = torch.stack([torch.full((9,), float(i), dtype=torch.float64) for i in mine])
        mean = parallel.mean_depth_metrics(per_image)
        # 4. timing: max over ranks; 5. the criterion's normaliser
        t = parallel.max_over_ranks(10.0 + rk)
        n = parallel.global_num_items(12 + 5 * rk)
        # 6. the whole-model step's exchange: several flat buffers started asynchronously in backward order, one wait, then ONE
        #    clip norm over all of them (what Trainer._exchange / Trainer.step do with gwd_sumsq / gwd_adamw_step on the GPU)
        bufs = [torch.full((n_,), float(rk + 1 + i), dtype=torch.float32) for i, n_ in enumerate((5, 1000, 33))]
        works = parallel.allreduce_async(bufs[:2]) + parallel.allreduce_async(bufs[2:])
        wa = parallel.wait_all(works)
        sumsq = sum(float((b / wa).double().pow(2).sum()) for b in bufs)
        torch.save({"mine": mine, "G": G / w, "mean": mean, "t": t, "n": n, "w": w, "bufs": [b / wa for b in bufs], "sumsq": sumsq},
                   os.path.join(out, "r%d.pt" % rk))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r = [torch.load(os.path.join(tmp_path, "r%d.pt" % i)) for i in range(world)]
    assert r[0]["mine"] == [0, 1, 2, 3] and r[1]["mine"] == [4, 5, 6]
    base = torch.arange(1000, dtype=torch.float32)
    for x in r:
        assert x["w"] == 2 and torch.equal(x["G"], base * 1.5)
        assert torch.allclose(x["mean"], torch.full((9,), 3.0, dtype=torch.float64))
        assert x["t"] == 11.0 and x["n"] == (12 + 17) / 2
        for i, b in enumerate(x["bufs"]):                       # mean over the two ranks of (rank + 1 + i)
            assert torch.equal(b, torch.full_like(b, 1.5 + i))
        assert abs(x["sumsq"] - (5 * 1.5 ** 2 + 1000 * 2.5 ** 2 + 33 * 3.5 ** 2)) < 1e-6


def test_single_process_is_a_no_op():
    assert parallel.world_size() == 1 and list(parallel.shard(5)) == [0, 1, 2, 3, 4]
    g = torch.ones(4)
    assert parallel.allreduce_sum_(g) == 1 and torch.equal(g, torch.ones(4))
    assert parallel.max_over_ranks(3.5) == 3.5 and parallel.global_num_items(0) == 1.0
    assert parallel.allreduce_async([g]) == [] and parallel.wait_all([]) == 1
    m = parallel.mean_depth_metrics(torch.tensor([[1.0] * 9, [3.0] * 9], dtype=torch.float64))
    assert torch.allclose(m, torch.full((9,), 2.0, dtype=torch.float64))
