"""Multi-GPU driver logic: the batch is an array of independent trajectories, so N GPUs = N contiguous index ranges,
one process (one rk_handle) per GPU, no data-path collective. torch.distributed is used for exactly two
things: a barrier around the timed region and a MAX-reduction of the per-rank device times.
"""
from __future__ import annotations


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced [lo, hi) of trajectory indices owned by `rank`; the ranges tile [0, n) exactly."""
    if world < 1 or not (0 <= rank < world) or n < 0:
        raise ValueError("bad shard arguments")
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def max_over_ranks(value: float, device=None) -> float:
    """MAX of a per-rank scalar over the default process group (identity when not initialised)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_host(local, n_total: int, rank: int, world: int):
    """The "final host gather" of a sharded result: every rank contributes its slice of a [.., n] host array; rank 0
    returns the assembled array (others None). Only used for results that must leave the GPUs; never on the hot path."""
    import numpy as np
    import torch.distributed as dist
    if world == 1:
        return local
    parts = [None] * world
    dist.all_gather_object(parts, local)
    if rank != 0:
        return None
    return np.concatenate(parts, axis=-1)
