"""Relocalisation scoring (BASELINE.json configs[4]): many pose hypotheses per frame, each scored by the ICP residual
reduction of the tracker (slam_odom_score_poses), sharded over the ranks of a torch.distributed group; the best pose is
chosen with ONE min-allreduce of a packed 64-bit key (error bits << 32 | hypothesis index).

The score of a hypothesis is the reference's acceptance statistic lastICPError = sqrt(residual) / count
(RGBDOdometryef.cpp:505-507), +inf when fewer than `min_inliers` pixels associate (lc/Ferns.cpp:262 rejects on the same
two numbers).  Non-negative IEEE-754 floats order like their bit patterns, so the integer minimum of the keys is the
minimum error, ties broken by the smaller index: the result does not depend on the number of ranks.
"""
from __future__ import annotations

import numpy as np


def icp_error(residual: np.ndarray, count: np.ndarray, min_inliers: float = 1.0) -> np.ndarray:
    residual = np.asarray(residual, np.float32)
    count = np.asarray(count, np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        err = np.sqrt(residual) / count
    err = np.where(count >= min_inliers, err, np.float32(np.inf)).astype(np.float32)
    return np.where(np.isnan(err), np.float32(np.inf), err).astype(np.float32)


def pack_keys(err: np.ndarray, index: np.ndarray) -> np.ndarray:
    """(error >= 0 as float32, global hypothesis index < 2^31) -> int64 keys whose integer order is (error, index)."""
    err = np.ascontiguousarray(err, np.float32)
    assert not np.any(err < 0) and not np.any(np.isnan(err))
    bits = err.view(np.uint32).astype(np.int64)
    return (bits << 32) | np.asarray(index, np.int64)


def unpack_key(key: int):
    key = int(key)
    err = np.array([key >> 32], np.uint32).view(np.float32)[0]
    return float(err), int(key & 0xFFFFFFFF)


def shard_range(n: int, rank: int, world: int):
    """Contiguous block of hypotheses owned by `rank` (sizes differ by at most one)."""
    return n * rank // world, n * (rank + 1) // world


def best_key(local_keys: np.ndarray, group=None, device=None) -> int:
    """min over all ranks of the packed keys: the only collective of the path."""
    import torch
    import torch.distributed as dist
    local = int(local_keys.min()) if len(local_keys) else np.iinfo(np.int64).max
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    t = torch.tensor([local], dtype=torch.int64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    return int(t.item())


def score_sharded(odo, level, prev_pose, trans_n, rot_n, rank=0, world=1, min_inliers=1.0, group=None, device=None):
    """Each rank scores its block of the n hypotheses on its own GPU (every rank holds the same frame); returns
    (best global index, its error, local errors of this rank's block)."""
    trans_n = np.asarray(trans_n, np.float32).reshape(-1, 3)
    rot_n = np.asarray(rot_n, np.float32).reshape(-1, 3, 3)
    lo, hi = shard_range(len(trans_n), rank, world)
    if hi > lo:
        res, cnt = odo.score_poses(level, prev_pose, trans_n[lo:hi], rot_n[lo:hi])
        err = icp_error(res, cnt, min_inliers)
    else:
        err = np.zeros(0, np.float32)
    keys = pack_keys(err, np.arange(lo, hi))
    e, i = unpack_key(best_key(keys, group, device))
    return i, e, err
