"""Multi-GPU plumbing: one process per GPU, members sharded, ONE collective at the end.

Ensemble members never interact (all state is re-initialised per member: RRMODEL/dyna_initialise_run.f90:27-36,
dyna_init_satzone.f90:194-222; DTA/dta_route_processing.f90:279-292), so every rank runs its own
contiguous block of members independently.  The host samples the whole parameter table once with the
sequential RANMAR stream and hands each rank its slice, which makes the results independent of the
number of GPUs.  The only message is the final gather of per-member performance measures and status
words to rank 0 (torch.distributed: NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np


def shard_range(n_member: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of rank `rank`: blocks differ by at most one member."""
    base, rem = divmod(int(n_member), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_to_rank0(local, n_total: int, group=None):
    """Gathers ragged member-slowest tensors (first dim = local members) to rank 0.
    `local` is a torch tensor (CUDA for NCCL, CPU for gloo).  Returns the (n_total, ...) tensor on
    rank 0 and None elsewhere."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [shard_range(n_total, r, world)[1] - shard_range(n_total, r, world)[0] for r in range(world)]
    assert local.shape[0] == sizes[rank], (local.shape, sizes, rank)
    pad = max(sizes)
    buf = torch.zeros((pad,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    buf[: sizes[rank]] = local
    if rank == 0:
        outs = [torch.empty_like(buf) for _ in range(world)]
        dist.gather(buf, outs, dst=0, group=group)
        return torch.cat([o[: sizes[r]] for r, o in enumerate(outs)], dim=0)
    dist.gather(buf, None, dst=0, group=group)
    return None


class PerfGather:
    """The end-of-run gather of the per-member measures for EQUAL member blocks: receive buffers are
    allocated once on rank 0 (`out`: (world, n_local, ...), rank r's block at out[r]) and reused by every call, so a
    call is exactly one `dist.gather` -- no allocation, no padding copy, no concatenation."""

    def __init__(self, n_local: int, tail_shape, world: int, rank: int, device, dtype=None, group=None):
        import torch
        self.group, self.rank = group, rank
        self.out = None
        self.outs = None
        if rank == 0:
            self.out = torch.empty((world, n_local) + tuple(tail_shape), dtype=dtype or torch.float64, device=device)
            self.outs = list(self.out.unbind(0))

    def __call__(self, local):
        import torch.distributed as dist
        dist.gather(local, self.outs, dst=0, group=self.group)
        return self.out


def lpt_assign(costs, world: int):
    """Longest-processing-time-first assignment of catchments to ranks (multi-catchment runs,
    SURVEY §8e): returns a list of rank ids, one per catchment."""
    order = np.argsort(-np.asarray(costs, dtype=np.float64), kind="stable")
    load = np.zeros(world)
    owner = np.zeros(len(costs), dtype=np.int64)
    for i in order:
        r = int(np.argmin(load))
        owner[i] = r
        load[r] += costs[i]
    return owner.tolist()
