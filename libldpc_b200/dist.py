"""Multi-GPU host: one process per GPU (torchrun), frames of every round sharded over ranks, the
per-point counters {fec, bec, frames, iters} combined by ONE tiny all-reduce per round
(NCCL over NVLink on GPUs, gloo in CPU tests).  No other data-path collective exists: frames are
independent and the code tables are replicated per GPU."""
import ctypes as ct

import numpy as np
import torch
import torch.distributed as dist


def shard_range(frame0, n_frames, rank, world):
    """Contiguous split of [frame0, frame0+n) used by the sweep driver (sim_driver.cpp)."""
    lo = frame0 + n_frames * rank // world
    hi = frame0 + n_frames * (rank + 1) // world
    return lo, hi


def make_allreduce(device=None, group=None):
    """Returns a Python callable (values_ptr, n, user) -> None summing a uint64 array over ranks."""
    buf = {}

    def _allreduce(ptr, n, _user):
        # the sweep driver's rounds are pipelined (the next round is already running on the GPU while this is called), so this
        # small host round trip is off the device's critical path; the staging tensors are allocated once
        arr = np.ctypeslib.as_array(ptr, shape=(n,)).view(np.int64)
        if n not in buf:
            host = torch.empty(n, dtype=torch.int64, pin_memory=device is not None)
            buf[n] = (host, torch.empty(n, dtype=torch.int64, device=device) if device is not None else host)
        host, t = buf[n]
        host.numpy()[:] = arr
        if device is not None:
            t.copy_(host, non_blocking=True)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        if device is not None:
            host.copy_(t)
        arr[:] = host.numpy()
    return _allreduce


def simulate_distributed(ctx, snr, **kw):
    """Runs Context.simulate with this process' rank/world from torch.distributed (must be initialised)."""
    rank, world = dist.get_rank(), dist.get_world_size()
    device = None
    if dist.get_backend() == "nccl":
        device = torch.device("cuda", torch.cuda.current_device())
    return ctx.simulate(snr, rank=rank, world=world, allreduce=make_allreduce(device), **kw)
