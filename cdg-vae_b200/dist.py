"""Data-parallel gradient exchange: one process per GPU, batch rows sharded,
parameters replicated, one all-reduce(sum) of the gradient arena per step over NCCL
(gloo in the CPU tests).  The reference has no multi-GPU path; this is new surface."""
import torch
import torch.distributed as dist


def world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank():
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def allreduce_arena(grads, ranges, bucket_floats=8 << 20):
    """Sum `grads[o:o+n]` for every (o, n) in `ranges` across ranks, in buckets issued in reverse
    (backward) order.  Returns the scale that turns the sum into the mean-of-shards gradient."""
    w = world()
    if w == 1:
        return 1.0
    handles = []
    for o, n in reversed(list(ranges)):
        for s in range(0, n, bucket_floats):
            e = min(n, s + bucket_floats)
            handles.append(dist.all_reduce(grads[o + s:o + e], op=dist.ReduceOp.SUM, async_op=True))
    for h in handles:
        h.wait()
    return 1.0 / w


def allreduce_mean_(t):
    w = world()
    if w > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        t.div_(w)
    return t


def shard_rows(n_rows):
    """Contiguous row shard [lo, hi) of this rank for a global batch of n_rows (equal shards)."""
    w, r = world(), rank()
    if n_rows % w != 0:
        raise ValueError(f"global batch {n_rows} must be divisible by world size {w}")
    per = n_rows // w
    return r * per, (r + 1) * per
