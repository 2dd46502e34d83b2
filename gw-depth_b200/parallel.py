"""Host-side data-parallel plumbing (one process per GPU, torch.distributed; NCCL on the box, gloo in the CPU tests).

The path shards over the batch with no data-path collective.  What does cross ranks:
  * training: one sum all-reduce per flat gradient buffer per step (15 buffers for the whole model, started asynchronously as
    their backward completes: allreduce_async / wait_all), and the scalar `num_items` of SetCriterion
    (src/models/glassrgbd.py:324-326);
  * evaluation: one all-reduce of [9 metric sums, image count] (src/engine_glassrgbd.py:243-264 averages per image);
  * benchmarking: the max over ranks of the device-timed interval.
Every helper is a no-op for a single process.
"""
import torch
import torch.distributed as dist


def world_size():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank():
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def shard(n_items, rk=None, world=None):
    """contiguous, balanced slice of range(n_items) owned by rank `rk` (earlier ranks take the remainder)"""
    rk = rank() if rk is None else rk
    world = world_size() if world is None else world
    base, rem = divmod(n_items, world)
    start = rk * base + min(rk, rem)
    return range(start, start + base + (1 if rk < rem else 0))


def allreduce_sum_(flat):
    """in-place sum over ranks of a flat buffer (the gradient all-reduce); returns the world size to divide by"""
    w = world_size()
    if w > 1:
        dist.all_reduce(flat)
    return w


def allreduce_async(buffers):
    """start the sum all-reduce of several flat gradient buffers and return the work handles (empty for one process): the
    whole-model training step issues one per stage module as soon as its backward has been enqueued, so the exchange overlaps
    the rest of the backward (train_model.Trainer._exchange)"""
    if world_size() == 1:
        return []
    return [dist.all_reduce(b, async_op=True) for b in buffers]


def wait_all(works):
    """block the current stream (CUDA) / the caller (CPU) until every started all-reduce has finished; returns the world size
    the summed gradients have to be divided by"""
    for w in works:
        w.wait()
    return world_size()


def max_over_ranks(value, device="cpu"):
    if world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def mean_depth_metrics(per_image):
    """per_image: fp64 [n_local, 9] (gwd_depth_metrics rows of this rank's images) -> fp64 [9] mean over ALL images of all
    ranks, exactly what the reference's batch-1 evaluation loop averages"""
    acc = torch.cat([per_image.sum(0), per_image.new_tensor([float(per_image.shape[0])])])
    if world_size() > 1:
        dist.all_reduce(acc)
    return acc[:9] / acc[9].clamp_min(1.0)


def global_num_items(n_local, device="cpu"):
    """SetCriterion's normaliser: clamp(sum over ranks / world, min 1) (src/models/glassrgbd.py:322-326)"""
    t = torch.tensor([float(n_local)], dtype=torch.float32, device=device)
    if world_size() > 1:
        dist.all_reduce(t)
    return float(torch.clamp(t / world_size(), min=1).item())
