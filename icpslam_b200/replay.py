"""Offline replay of recorded sweeps sharded one block of scans per GPU (BASELINE
configs[3]).  Each (scan i-1, scan i) pair is an independent unit — the reference's per-callback state
is only prev_cloud_ (reference src/icpslam/icp_odometer.cpp:179-182,209) — so ranks get contiguous
blocks of pairs, run them with b2icp_align_batch and exchange only the fixed-size per-scan records:
one collective per batch (NCCL on GPUs, gloo in the CPU tests).  After the gather comes the serial part: pairs
that follow a REJECTED scan are re-registered against the last accepted sweep (the reference keeps prev_cloud_
then, icp_odometer.cpp:201-209), and the SE(3) prefix composition pose_i = pose_{i-1} o T_i
(icp_odometer.cpp:111-113).
"""
from __future__ import annotations

import numpy as np

from . import pose6dof

RECORD = 22  # T[16], converged, iterations, n_corr, mse, fitness, index of the sweep registered against
FITNESS_ACCEPT = 20.0  # the literal of icp_odometer.cpp:201
HARD_ERRORS = (-1, -5, -7)  # INVALID_ARG, SOLVER_FAILED, CUDA: anything else is a per-pair status kept in the record


def shard_range(n_items: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous block [lo, hi) of rank `rank`; sizes differ by at most one."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def pack_results(results, first_target: int = 0) -> np.ndarray:
    """One record per pair; pair k of a block registers sweep first_target+k+1 against sweep first_target+k."""
    out = np.zeros((len(results), RECORD), dtype=np.float64)
    for i, r in enumerate(results):
        out[i, :16] = list(r.T)
        out[i, 16:20] = (r.converged, r.iterations, r.n_corr_last, r.mse_last)
        out[i, 20] = r.fitness
        out[i, 21] = first_target + i
    return out


def accepted(rec, fitness_accept: float = FITNESS_ACCEPT) -> bool:
    """`icp.hasConverged() && icp.getFitnessScore() < 20` (icp_odometer.cpp:201).  A record without a fitness
    (NaN: the batch did not ask for it) is judged on convergence alone."""
    fit = rec[20]
    return bool(rec[16] > 0 and (np.isnan(fit) or fit < fitness_accept))


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


def compose_odometry(records: np.ndarray, initial_pose=None, fitness_accept: float = FITNESS_ACCEPT) -> np.ndarray:
    """pose_i = pose_{i-1} o T_i for accepted scans (icp_odometer.cpp:111-113, 201-209); a rejected scan keeps
    the previous pose, as the reference drops it.  Only valid on records whose targets follow the reference's
    rule (every pair registered against the LAST ACCEPTED sweep): fixup_rejected() establishes that."""
    pose = pose6dof.identity() if initial_pose is None else np.asarray(initial_pose, np.float64)
    poses = [pose]
    last_ok = 0
    for i, rec in enumerate(records):
        if accepted(rec, fitness_accept):
            if int(rec[21]) != last_ok:
                raise ValueError(f"pair {i} was registered against sweep {int(rec[21])} but the last accepted sweep is "
                                 f"{last_ok}: run fixup_rejected() first")
            pose = pose6dof.compose(pose, pose6dof.from_matrix(rec[:16]))
            last_ok = i + 1
        poses.append(pose)
    return np.stack(poses)


def fixup_rejected(records: np.ndarray, sweeps, registration, fitness_accept: float = FITNESS_ACCEPT) -> int:
    """The reference only does `*prev_cloud_ = *curr_cloud_` after an ACCEPTED scan (icp_odometer.cpp:201-209): after
    a rejected scan i the next cloud is registered against the older prev_cloud_, not against sweep i.  The batch
    registered every sweep against its predecessor, which is the same thing while every pair is accepted; this
    serial pass re-registers, in order, every pair whose target was not the last accepted sweep.  Returns the
    number of pairs it re-registered.  (A chain of rejections stays serial — as it is in the reference.)"""
    redone = 0
    last_ok = 0
    for i in range(len(records)):
        if int(records[i, 21]) != last_ok:
            rc, res = registration.alignBatch([sweeps[i + 1]], [sweeps[last_ok]], with_fitness=True)
            if rc in HARD_ERRORS:
                raise RuntimeError(f"b2icp_align_batch failed with status {rc} on pair {i}")
            records[i] = pack_results(res, last_ok)[0]
            redone += 1
        if accepted(records[i], fitness_accept):
            last_ok = i + 1
    return redone


def replay_pairs(sweeps, registration, rank: int = 0, world: int = 1, device=None):
    """Register sweep i against sweep i-1 for every i >= 1, this rank's block through b2icp_align_batch
    (consecutive mode, with getFitnessScore so that the reference's accept test can be applied), gather, then
    the serial fix-up of the pairs that follow a rejected scan.  Returns (records[n-1, RECORD], poses[n, 7])."""
    n_pairs = len(sweeps) - 1
    lo, hi = shard_range(n_pairs, world, rank)
    if hi > lo:
        srcs = [sweeps[i + 1] for i in range(lo, hi)]
        tgts = [sweeps[lo]] + [None] * (hi - lo - 1)
        rc, res = registration.alignBatch(srcs, tgts, with_fitness=True)
        if rc in HARD_ERRORS:
            raise RuntimeError(f"b2icp_align_batch failed with status {rc}")
        local = pack_results(res, lo)
    else:
        local = np.zeros((0, RECORD))
    records = gather_records(local, n_pairs, device)
    fixup_rejected(records, sweeps, registration)  # every rank does the same serial pass: identical tables
    return records, compose_odometry(records)
