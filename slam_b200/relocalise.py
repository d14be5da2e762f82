"""Relocalisation scoring (BASELINE.json configs[4]): many pose hypotheses per frame, each scored by the ICP residual
reduction of the tracker (slam_odom_score_poses), sharded over the ranks of a torch.distributed group; the best pose is
chosen with ONE min-allreduce of a packed 64-bit key (error bits << 32 | hypothesis index).

The score of a hypothesis is the reference's acceptance statistic lastICPError = sqrt(residual) / count
(RGBDOdometryef.cpp:505-507), +inf when fewer than `min_inliers` pixels associate (lc/Ferns.cpp:262 rejects on the same
two numbers).  Non-negative IEEE-754 floats order like their bit patterns, so the integer minimum of the keys is the
minimum error, ties broken by the smaller index: the result does not depend on the number of ranks.

Two forms: score_sharded() reads the per-hypothesis sums back and packs the keys on the host (parity tests look at every
hypothesis); score_sharded_device() is the product path -- the scoring launch itself folds its block's best key into one device
word (atomicMin, slam_odom_score_poses_best), that word is min-all-reduced in place by NCCL, and only the winning key (8 bytes)
ever reaches the host.  broadcast_frame() replicates rank 0's frame (depth + the two predicted maps) once per frame.
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
    # world == 1: this rank scores everything alone, whatever process group may exist (the single-GPU check of a sharded run)
    e, i = unpack_key(int(keys.min()) if world == 1 else best_key(keys, group, device))
    return i, e, err


INT64_MAX = np.iinfo(np.int64).max


def perturbed_hypotheses(gt_pose, n, sigma_t=0.05, sigma_r_deg=3.0, seed=0xBEEF):
    """SURVEY 8(d) configs[4]: n hypotheses = gt_pose o random SE3 perturbation (translation sigma 5 cm, rotation sigma 3 deg about a
    random axis), hypothesis 0 unperturbed.  -> (trans[n,3], rot[n,3,3]) float32; identical on every rank for the same seed."""
    rng = np.random.default_rng(seed)
    gt = np.asarray(gt_pose, np.float64)
    T = np.empty((n, 3), np.float32)
    R = np.empty((n, 3, 3), np.float32)
    for i in range(n):
        if i == 0:
            dR, dt = np.eye(3), np.zeros(3)
        else:
            w = rng.normal(scale=np.deg2rad(sigma_r_deg), size=3)
            th = np.linalg.norm(w)
            k = w / th if th > 0 else np.array([1.0, 0, 0])
            K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
            dR = np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K
            dt = rng.normal(scale=sigma_t, size=3)
        R[i] = (gt[:3, :3] @ dR).astype(np.float32)
        T[i] = (gt[:3, 3] + gt[:3, :3] @ dt).astype(np.float32)
    return T, R


def broadcast_frame(tensors, src=0, group=None):
    """Replicate the frame of rank `src` (device tensors: depth, predicted vertices / normals, ...) on every rank, in place."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        import torch
        for t in tensors:
            # NCCL has no 16-bit integer type: the u16 depth image travels as the bytes it is
            dist.broadcast(t.view(torch.uint8) if t.dtype in (torch.int16, torch.uint16) else t, src=src, group=group)


def score_sharded_device(odo, level, prev_pose, trans_n, rot_n, key_tensor, rank=0, world=1, min_inliers=1.0, group=None, stream=None):
    """Product path: this rank's block of hypotheses is scored on its GPU, the launch leaves the block's best packed key in
    key_tensor (one int64 on the device), one NCCL min-all-reduce of that word picks the winner of all ranks.  Nothing but the
    final 8 bytes is read back.  stream: torch stream wrapping the handle's CUDA stream (the collective is ordered behind the
    scoring launch on it).  -> (best global index, its error)."""
    import torch
    import torch.distributed as dist
    trans_n = np.asarray(trans_n, np.float32).reshape(-1, 3)
    rot_n = np.asarray(rot_n, np.float32).reshape(-1, 3, 3)
    lo, hi = shard_range(len(trans_n), rank, world)
    ctx = torch.cuda.stream(stream) if stream is not None else _null_context()
    with ctx:
        key_tensor.fill_(INT64_MAX)
        if hi > lo:
            odo.score_poses_best(level, prev_pose, trans_n[lo:hi], rot_n[lo:hi], key_tensor, index_base=lo, min_inliers=min_inliers)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(key_tensor, op=dist.ReduceOp.MIN, group=group)
        key = int(key_tensor.item())
    e, i = unpack_key(key)
    return i, e


def connect_peers(odo, rank=None, world=None, group=None) -> bool:
    """Exchange the CUDA IPC handles of the ranks' slot arrays (once per handle) so that score_sharded_peers can take the minimum
    over NVLink peer memory.  Returns False (and leaves the handle unconnected) when the processes cannot map each other's memory --
    the caller then stays on score_sharded_device (NCCL)."""
    import torch.distributed as dist
    rank = dist.get_rank(group) if rank is None else rank
    world = dist.get_world_size(group) if world is None else world
    mine = odo.peer_export()
    handles = [None] * world
    if world > 1:
        dist.all_gather_object(handles, mine, group=group)
    else:
        handles[0] = mine
    ok = True
    try:
        odo.peer_connect(rank, world, b"".join(handles))
    except Exception:
        ok = False
    if world > 1:   # all or none
        flags = [None] * world
        dist.all_gather_object(flags, ok, group=group)
        ok = all(flags)
    return ok


def score_sharded_peers(odo, level, prev_pose, trans_n, rot_n, rank=0, world=1, min_inliers=1.0):
    """Product path on one NVLink node: this rank's block of hypotheses is scored, one warp writes the block's best packed key into
    every peer's memory and takes the minimum of what the peers wrote here (slam_odom_score_poses_best_peers) -- no collective
    library, no host round trip between scoring and reduction.  Collective: every rank calls it.  -> (best global index, its error)."""
    trans_n = np.asarray(trans_n, np.float32).reshape(-1, 3)
    rot_n = np.asarray(rot_n, np.float32).reshape(-1, 3, 3)
    lo, hi = shard_range(len(trans_n), rank, world)
    key = odo.score_poses_best_peers(level, prev_pose, trans_n[lo:hi], rot_n[lo:hi], index_base=lo, min_inliers=min_inliers)
    e, i = unpack_key(key)
    return i, e


class _null_context:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False
