"""Data-parallel gradient exchange: one process per GPU, batch rows sharded,
parameters replicated, one all-reduce(sum) of the gradient arena per step over NCCL
(gloo in the CPU tests).  The reference has no multi-GPU path; this is new surface."""
import torch
import torch.distributed as dist


def world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank():
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


class OneShot:
    """Symmetric peer buffer for `cdg_allreduce_oneshot` (csrc/oneshot.cu): small gradient arenas are summed by one kernel
    per rank over NVLink peer loads instead of an NCCL call (~40 us inside the step's graph at 8 GPUs for 2.4 KB)."""
    MAX_FLOATS = 16384
    _inst = {}

    def __init__(self, device):
        import ctypes as C
        import torch.distributed._symmetric_memory as symm
        g = dist.group.WORLD
        try:
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                symm.enable_symm_mem_for_group(g.group_name)
        except Exception:
            pass
        self.buf = symm.empty(2 * self.MAX_FLOATS + 64, dtype=torch.float32, device=device)
        self.buf.zero_()
        hdl = symm.rendezvous(self.buf, g)
        self.hdl = hdl
        self.world, self.rank = hdl.world_size, hdl.rank
        self.ptrs = (C.c_uint64 * self.world)(*[int(p) for p in hdl.buffer_ptrs])
        self.step = torch.zeros(1, dtype=torch.int32, device=device)
        torch.cuda.synchronize(device)
        dist.barrier()                                  # every rank's flags are zero before anyone raises one

    @classmethod
    def get(cls, device):
        """The process-wide instance for `device`, or None where peer-mapped symmetric memory is not available (or switched
        off with CDG_ONESHOT=0): the caller then uses NCCL.  Collective: every rank reaches this at its first exchange."""
        import os
        key = (device.type, device.index)
        if key not in cls._inst:
            inst = None
            if device.type == "cuda" and os.environ.get("CDG_ONESHOT", "1") != "0" and dist.get_backend() == "nccl":
                try:
                    inst = cls(device)
                except Exception as e:                  # no P2P / fabric handles on this box
                    import warnings
                    warnings.warn(f"one-shot all-reduce unavailable ({type(e).__name__}: {e}); using NCCL")
                # all ranks must agree, or some would wait in the kernel for flags that never come
                ok = torch.tensor([1 if inst is not None else 0], device=device)
                dist.all_reduce(ok, op=dist.ReduceOp.MIN)
                if int(ok.item()) == 0:
                    inst = None
            cls._inst[key] = inst
        return cls._inst[key]

    def allreduce_(self, grads):
        import ctypes as C
        from . import _lib
        stream = torch.cuda.current_stream(grads.device).cuda_stream
        with torch.cuda.device(grads.device):
            _lib.check(_lib.lib().cdg_allreduce_oneshot(C.c_void_p(grads.data_ptr()), grads.numel(), self.MAX_FLOATS, self.ptrs,
                                                        self.world, self.rank, C.c_void_p(self.step.data_ptr()),
                                                        C.c_void_p(stream)))


def allreduce_arena(grads, ranges, bucket_floats=8 << 20):
    """Sum `grads[o:o+n]` for every (o, n) in `ranges` across ranks, in buckets issued in reverse
    (backward) order.  Returns the scale that turns the sum into the mean-of-shards gradient."""
    w = world()
    if w == 1:
        return 1.0
    if grads.is_cuda and grads.numel() <= OneShot.MAX_FLOATS and grads.dtype == torch.float32 and grads.is_contiguous():
        one = OneShot.get(grads.device)
        if one is not None:
            one.allreduce_(grads)                       # the whole arena (padding included) in one launch
            return 1.0 / w
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
