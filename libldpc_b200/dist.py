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
    def _allreduce(ptr, n, _user):
        arr = np.ctypeslib.as_array(ptr, shape=(n,))
        t = torch.from_numpy(arr.astype(np.int64))
        if device is not None:
            t = t.to(device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        arr[:] = t.cpu().numpy().astype(np.uint64)
    return _allreduce


def simulate_distributed(ctx, snr, **kw):
    """Runs Context.simulate with this process' rank/world from torch.distributed (must be initialised)."""
    rank, world = dist.get_rank(), dist.get_world_size()
    device = None
    if dist.get_backend() == "nccl":
        device = torch.device("cuda", torch.cuda.current_device())
    return ctx.simulate(snr, rank=rank, world=world, allreduce=make_allreduce(device), **kw)
