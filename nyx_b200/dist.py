"""Multi-GPU ensemble sharding: one process per GPU, contiguous index ranges, ONE all-gather.

The reference fans runs out over a rayon pool and sorts results by run index
(mc/montecarlo.rs:233-253, 266-267).  Runs are independent, so across GPUs the ensemble is
split into contiguous index ranges [g*N/G, (g+1)*N/G) — draw order == run index, hence no
sort — each rank integrates its shard with zero communication, and the final states are
exchanged with a single `all_gather` (NCCL over NVLink on GPUs; gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Tuple

import numpy as np


def shard_bounds(n: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of rank `rank`; sizes differ by at most one."""
    base, rem = divmod(n, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def propagate_batch_multi(engines, state_soa, consts_soa, epoch0_ns, end_epoch_ns, step_ns=None):
    """ONE ensemble over several GPUs of this process through the C ABI (`nyxb_propagate_batch_multi`): `engines` = one
    `nyx_b200.Engine` per device, all built from the same propagator; contiguous run-index shards, every device integrates
    concurrently, results land directly in the returned host arrays.  Same return value as `Engine.propagate_batch`."""
    import ctypes as C

    from . import abi
    from .propagator import PropagationError

    lib = abi.load_library()
    state_soa = np.ascontiguousarray(state_soa, dtype=np.float64)
    consts_soa = np.ascontiguousarray(consts_soa, dtype=np.float64)
    epoch0_ns = np.ascontiguousarray(epoch0_ns, dtype=np.int64)
    n = state_soa.shape[1]
    out = np.empty((9, n)); out_ep = np.empty(n, dtype=np.int64)
    det = np.zeros(n, dtype=abi.DETAILS_DTYPE); status = np.zeros(n, dtype=np.int32)
    handles = (C.c_void_p * len(engines))(*[e.handle for e in engines])
    rc = lib.nyxb_propagate_batch_multi(handles, len(engines), n, state_soa.ctypes.data, consts_soa.ctypes.data, epoch0_ns.ctypes.data,
                                        int(end_epoch_ns), step_ns.ctypes.data if step_ns is not None else None, out.ctypes.data,
                                        out_ep.ctypes.data, det.ctypes.data, status.ctypes.data)
    if rc != 0:
        raise PropagationError(f"nyxb_propagate_batch_multi failed ({rc}): {abi.last_error()}")
    return out, out_ep, det, status


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


def shard_tracking_arc(arc, world_size: int, rank: int):
    """The observation sets of this rank's filters: one tracking schedule, obs[m][2][lo:hi] (od/process/mod.rs:128-497 runs are
    independent, so an ensemble of filters shards exactly like an ensemble of propagations)."""
    from .od import TrackingDataArc

    lo, hi = shard_bounds(arc.n, world_size, rank)
    return TrackingDataArc(arc.epoch_ns, list(arc.tracker), np.ascontiguousarray(arc.obs[:, :, lo:hi]))


def sharded_process_arcs(odp, initial_estimates, arc, group=None, device=None):
    """n filters split over the ranks by contiguous index; every rank runs `process_arcs` on its shard (ONE launch), then
    ONE all-gather of [9 state + 81 covariance] per filter gives every rank the full, index-ordered set of final estimates.
    Returns (local ODSolution, final_state[9][n], covar[n][9][9])."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n = len(initial_estimates)
    lo, hi = shard_bounds(n, world, rank)
    sol = odp.process_arcs(initial_estimates[lo:hi], shard_tracking_arc(arc, world, rank))
    local = np.concatenate([sol.final_state_soa, sol.covar.reshape(hi - lo, 81).T], axis=0)  # [90][n_local]
    t = torch.from_numpy(np.ascontiguousarray(local))
    if device is not None:
        t = t.to(device)
    full = all_gather_final_states(t, n, group).cpu().numpy()
    return sol, full[:9], np.ascontiguousarray(full[9:].T).reshape(n, 9, 9)
