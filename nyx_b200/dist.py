"""Multi-GPU ensemble sharding: one process per GPU, contiguous index ranges, ONE all-gather.

The reference fans runs out over a rayon pool and sorts results by run index
(mc/montecarlo.rs:233-253, 266-267).  Runs are independent, so across GPUs the ensemble is
split into contiguous index ranges [g*N/G, (g+1)*N/G) — draw order == run index, hence no
sort — each rank integrates its shard with zero communication, and the final states are
exchanged with a single `all_gather` (NCCL over NVLink on GPUs; gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np


def shard_bounds(n: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of rank `rank`; sizes differ by at most one."""
    base, rem = divmod(n, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_soa(arr: np.ndarray, world_size: int, rank: int) -> np.ndarray:
    lo, hi = shard_bounds(arr.shape[-1], world_size, rank)
    return np.ascontiguousarray(arr[..., lo:hi])


def all_gather_final_states(local, n_total: int, group=None):
    """Gather per-rank [rows, n_local] tensors (torch, any device) into [rows, n_total] on every rank.

    Shards may differ by one column, so each rank pads to the maximum shard width; a single
    `all_gather_into_tensor` moves the data, then the padding is dropped.
    """
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    rows = local.shape[0]
    width = max(shard_bounds(n_total, world, r)[1] - shard_bounds(n_total, world, r)[0] for r in range(world))
    send = local
    if local.shape[1] != width:
        send = torch.zeros((rows, width), dtype=local.dtype, device=local.device)
        send[:, : local.shape[1]] = local
    recv = torch.empty((world, rows, width), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(recv.view(-1), send.contiguous().view(-1), group=group)
    out = torch.empty((rows, n_total), dtype=local.dtype, device=local.device)
    for r in range(world):
        lo, hi = shard_bounds(n_total, world, r)
        out[:, lo:hi] = recv[r, :, : hi - lo]
    return out
