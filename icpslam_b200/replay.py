"""Offline replay of recorded sweeps sharded one block of scans per GPU (BASELINE
configs[3]).  Each (scan i-1, scan i) pair is an independent unit — the reference's per-callback state
is only prev_cloud_ (reference src/icpslam/icp_odometer.cpp:179-182,209) — so ranks get contiguous
blocks of pairs, run them with b2icp_align_batch and exchange only the fixed-size per-scan records:
one collective per batch (NCCL on GPUs, gloo in the CPU tests).  Rank 0 then does the serial SE(3)
prefix composition pose_i = pose_{i-1} o T_i (icp_odometer.cpp:111-113).
"""
from __future__ import annotations

import numpy as np

from . import pose6dof

RECORD = 20  # T[16], converged, iterations, n_corr, mse


def shard_range(n_items: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous block [lo, hi) of rank `rank`; sizes differ by at most one."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def pack_results(results) -> np.ndarray:
    out = np.zeros((len(results), RECORD), dtype=np.float64)
    for i, r in enumerate(results):
        out[i, :16] = list(r.T)
        out[i, 16:] = (r.converged, r.iterations, r.n_corr_last, r.mse_last)
    return out


def gather_records(local: np.ndarray, n_total: int, device=None) -> np.ndarray:
    """All ranks contribute their [n_local, RECORD] block; returns the [n_total, RECORD] table in scan
    order on every rank.  One all_gather of equal-sized (padded) blocks."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    per = -(-n_total // world)  # the largest shard
    buf = torch.zeros((per, RECORD), dtype=torch.float64, device=device)
    buf[: len(local)] = torch.from_numpy(local).to(buf.device)
    out = torch.empty((world * per, RECORD), dtype=torch.float64, device=device)
    dist.all_gather_into_tensor(out, buf)
    out = out.cpu().numpy().reshape(world, per, RECORD)
    rows = []
    for r in range(world):
        lo, hi = shard_range(n_total, world, r)
        rows.append(out[r, : hi - lo])
    return np.concatenate(rows)


def compose_odometry(records: np.ndarray, initial_pose=None, fitness=None, fitness_accept: float = 20.0) -> np.ndarray:
    """pose_i = pose_{i-1} o T_i for accepted scans (converged, and fitness < 20 when given:
    icp_odometer.cpp:201); a rejected scan keeps the previous pose, as the reference drops it."""
    pose = pose6dof.identity() if initial_pose is None else np.asarray(initial_pose, np.float64)
    poses = [pose]
    for i, rec in enumerate(records):
        ok = rec[16] > 0 and (fitness is None or fitness[i] < fitness_accept)
        if ok:
            pose = pose6dof.compose(pose, pose6dof.from_matrix(rec[:16]))
        poses.append(pose)
    return np.stack(poses)


def replay_pairs(sweeps, registration, rank: int = 0, world: int = 1, device=None):
    """Register sweep i against sweep i-1 for every i >= 1, this rank's block through
    b2icp_align_batch (consecutive mode), then gather.  Returns (records[n-1, RECORD], poses[n, 7])."""
    n_pairs = len(sweeps) - 1
    lo, hi = shard_range(n_pairs, world, rank)
    if hi > lo:
        srcs = [sweeps[i + 1] for i in range(lo, hi)]
        tgts = [sweeps[lo]] + [None] * (hi - lo - 1)
        rc, res = registration.alignBatch(srcs, tgts)
        local = pack_results(res)
    else:
        local = np.zeros((0, RECORD))
    records = gather_records(local, n_pairs, device)
    return records, compose_odometry(records)
