"""Host-side multi-GPU plumbing for batched worlds.

Worlds are independent units: rank r of R owns a contiguous block of worlds and never exchanges
body data.  The only collective on the path is the SUM all-reduce of the 8-double statistics vector
(PgStats, include/physics_gpu.h) that the stats kernel writes straight into a caller-owned device
buffer.  These helpers are backend-agnostic so the same code runs under NCCL on GPUs and under gloo
in the CPU tests.
"""
from __future__ import annotations

import numpy as np

STATS_LEN = 8


def shard_range(total_worlds: int, rank: int, world_size: int) -> tuple[int, int]:
    """[first, first+count) of the worlds rank owns when `total_worlds` are split as evenly as
    possible in contiguous blocks (strong scaling); lower ranks take the remainder."""
    if not (0 <= rank < world_size) or total_worlds < 0:
        raise ValueError("bad rank / world_size / total")
    base, rem = divmod(total_worlds, world_size)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def weak_range(worlds_per_rank: int, rank: int) -> tuple[int, int]:
    """Weak scaling: every rank owns `worlds_per_rank` worlds, seeds continue across ranks."""
    return rank * worlds_per_rank, worlds_per_rank


def all_reduce_stats(vec, group=None):
    """SUM-reduce a length-8 float64 torch tensor (host or device) over the process group, in place."""
    import torch.distributed as dist
    if vec.numel() != STATS_LEN:
        raise ValueError("statistics vector must have 8 entries")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(vec, op=dist.ReduceOp.SUM, group=group)
    return vec


def stats_from_state(vel: np.ndarray, at_rest: np.ndarray | None = None) -> np.ndarray:
    """The same 8 aggregates the stats kernel produces, from host arrays (used to cross-check it)."""
    v = np.asarray(vel, np.float64).reshape(-1, 3)
    out = np.zeros(STATS_LEN, np.float64)
    out[0] = len(v)
    out[1] = 0.5 * float((v * v).sum())
    out[2:5] = v.sum(axis=0)
    out[6] = float(np.count_nonzero(at_rest)) if at_rest is not None else 0.0
    out[7] = float(np.count_nonzero(~np.isfinite(v).all(axis=1)))
    return out
