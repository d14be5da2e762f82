"""Trajectory evaluation: the reference's benchmark metrics for ``tick tx ty tz qx qy qz qw`` pose logs.

Python 3 restatement of what the reference ships as Python 2 scripts under ``benchmark/``:

  * ``read_stamped`` / ``associate``      benchmark/associate.py:50-105   greedy closest-stamp matching, each stamp used once
  * ``horn_align`` / ``absolute_error``   benchmark/evaluate_ate.py:47-79,116-150   ATE after a closed-form rigid alignment
  * ``pose_from_row`` / ``read_poses``    benchmark/evaluate_rpe.py:46-108
  * ``relative_errors``                   benchmark/evaluate_rpe.py:110-297 RPE over fixed-delta or all pairs, delta in s / m / rad / deg / frames

Pinned against golden values produced by the reference's own functions (tests/golden/make_eval_golden.py -> eval_golden.json).
Command line:  ``python -m slam_b200.evaluate ate GT EST [--verbose]``  /  ``python -m slam_b200.evaluate rpe GT EST --fixed_delta --delta_unit f``.
"""
from __future__ import annotations

import argparse
import random
import sys

import numpy as np

_EPS = np.finfo(float).eps * 4.0          # evaluate_rpe.py:44


# ---------------------------------------------------------------------------------------------- files
def _rows(path):
    """Non-comment lines split on blanks, commas and tabs."""
    with open(path) as f:
        for line in f.read().replace(",", " ").replace("\t", " ").split("\n"):
            if line and line[0] != "#":
                fields = [v for v in (w.strip() for w in line.split(" ")) if v != ""]
                if fields:
                    yield fields


def read_stamped(path) -> dict:
    """{stamp: [remaining fields as strings]} -- lines with a stamp only are dropped (associate.py:50-72)."""
    return {float(f[0]): f[1:] for f in _rows(path) if len(f) > 1}


def pose_from_row(row) -> np.ndarray:
    """(stamp, tx, ty, tz, qx, qy, qz, qw) -> 4x4; a (near-)zero quaternion gives the identity rotation (evaluate_rpe.py:46-74)."""
    t = row[1:4]
    q = np.array(row[4:8], dtype=np.float64)
    T = np.eye(4)
    T[:3, 3] = t
    nq = float(q @ q)
    if nq < _EPS:
        return T
    q = q * np.sqrt(2.0 / nq)
    o = np.outer(q, q)
    T[:3, :3] = [[1.0 - o[1, 1] - o[2, 2], o[0, 1] - o[2, 3], o[0, 2] + o[1, 3]],
                 [o[0, 1] + o[2, 3], 1.0 - o[0, 0] - o[2, 2], o[1, 2] - o[0, 3]],
                 [o[0, 2] - o[1, 3], o[1, 2] + o[0, 3], 1.0 - o[0, 0] - o[1, 1]]]
    return T


def read_poses(path) -> dict:
    """{stamp: 4x4}; rows with an all-zero quaternion or a NaN are skipped (evaluate_rpe.py:76-108)."""
    out = {}
    for k, f in enumerate(_rows(path)):
        v = [float(x) for x in f]
        if v[4:8] == [0, 0, 0, 0]:
            continue
        if any(np.isnan(x) for x in v):
            sys.stderr.write("Warning: line %d of file '%s' has NaNs, skipping line\n" % (k, path))
            continue
        out[v[0]] = pose_from_row(v)
    return out


# ---------------------------------------------------------------------------------------------- association
def associate(first: dict, second: dict, offset: float = 0.0, max_difference: float = 0.02) -> list:
    """Pairs (a, b) of stamps with |a - (b + offset)| < max_difference, best differences first, every stamp used at most once; sorted."""
    a_left, b_left = set(first.keys()), set(second.keys())
    candidates = sorted((abs(a - (b + offset)), a, b) for a in first for b in second if abs(a - (b + offset)) < max_difference)
    pairs = []
    for _, a, b in candidates:
        if a in a_left and b in b_left:
            a_left.discard(a)
            b_left.discard(b)
            pairs.append((a, b))
    return sorted(pairs)


# ---------------------------------------------------------------------------------------------- ATE
def horn_align(model: np.ndarray, data: np.ndarray):
    """Rigid (rot, trans) taking `model` (3 x n) onto `data` in the least-squares sense, and the per-point residual norms."""
    model, data = np.asarray(model, np.float64), np.asarray(data, np.float64)
    mc, dc = model.mean(1, keepdims=True), data.mean(1, keepdims=True)
    W = np.zeros((3, 3))
    for k in range(model.shape[1]):              # same accumulation order as the reference's loop (bit-for-bit sums)
        W += np.outer(model[:, k] - mc[:, 0], data[:, k] - dc[:, 0])
    U, _, Vh = np.linalg.svd(W.T)
    S = np.eye(3)
    if np.linalg.det(U) * np.linalg.det(Vh) < 0:
        S[2, 2] = -1
    rot = U @ S @ Vh
    trans = dc - rot @ mc
    residual = rot @ model + trans - data
    return rot, trans, np.sqrt((residual * residual).sum(0))


def absolute_error(gt_path, est_path, offset=0.0, scale=1.0, max_difference=0.02) -> dict:
    """evaluate_ate.py's __main__: associate, align the estimate onto the ground truth, statistics of the translational error."""
    first, second = read_stamped(gt_path), read_stamped(est_path)
    matches = associate(first, second, float(offset), float(max_difference))
    if len(matches) < 2:
        raise ValueError("Couldn't find matching timestamp pairs between groundtruth and estimated trajectory! Did you choose the correct sequence?")
    gt = np.array([[float(v) for v in first[a][0:3]] for a, _ in matches]).T
    est = np.array([[float(v) * float(scale) for v in second[b][0:3]] for _, b in matches]).T
    rot, trans, err = horn_align(est, gt)
    return dict(pairs=len(err), rmse=float(np.sqrt(err @ err / len(err))), mean=float(err.mean()), median=float(np.median(err)), std=float(err.std()),
                min=float(err.min()), max=float(err.max()), rot=rot, trans=trans, trans_error=err, matches=matches)


# ---------------------------------------------------------------------------------------------- RPE
def closest_index(values, t) -> int:
    """Index of the entry of the sorted list closest to t among those a bisection for t visits (evaluate_rpe.py:110-136; it is the
    overall closest except when the bisection passes over it, which the reference's results inherit)."""
    lo, hi = 0, len(values)
    best, gap = 0, abs(values[0] - t)
    while lo < hi:
        mid = (lo + hi) // 2
        d = abs(values[mid] - t)
        if d < gap:
            best, gap = mid, d
        if values[mid] == t:
            return mid
        if values[mid] > t:
            hi = mid
        else:
            lo = mid + 1
    return best


def _between(a, b):
    return np.linalg.inv(a) @ b                     # ominus


def _angle(T):
    return float(np.arccos(min(1.0, max(-1.0, (np.trace(T[:3, :3]) - 1.0) / 2.0))))


def _path_index(traj: dict, unit: str):
    """The quantity `delta` is measured in, per pose of the estimated trajectory (evaluate_rpe.py:246-259)."""
    stamps = sorted(traj)
    if unit == "s":
        return stamps
    if unit == "f":
        return list(range(len(stamps)))
    if unit not in ("m", "rad", "deg"):
        raise ValueError("Unknown unit for delta: '%s'" % unit)
    steps = [_between(traj[stamps[k + 1]], traj[stamps[k]]) for k in range(len(stamps) - 1)]
    factor = 180.0 / np.pi if unit == "deg" else 1.0
    acc, total = [0], 0
    for T in steps:
        total += float(np.linalg.norm(T[:3, 3])) if unit == "m" else _angle(T) * factor
        acc.append(total)
    return acc


def relative_errors(traj_gt: dict, traj_est: dict, max_pairs=10000, fixed_delta=False, delta=1.0, delta_unit="s", offset=0.0, scale=1.0, rng=None) -> list:
    """Rows [stamp_est_0, stamp_est_1, stamp_gt_0, stamp_gt_1, translational error, rotational error] (evaluate_rpe.py:204-297).

    With fixed_delta the second pose of a pair is the one whose path index is closest to index[i] + delta, and pairs ending on the LAST pose
    are dropped (the reference's rule); without it all pairs (or max_pairs random ones) are compared.  Ground-truth poses are the closest in
    time, rejected beyond twice the median ground-truth interval.  `rng` (random.Random) replaces the module-level generator the reference seeds
    with 0 when pairs have to be sampled."""
    rng = rng or random.Random(0)
    gt_stamps, est_stamps = sorted(traj_gt), sorted(traj_est)
    overlap = []
    for t in est_stamps:
        g = gt_stamps[closest_index(gt_stamps, t + offset)]
        back = est_stamps[closest_index(est_stamps, g - offset)]
        if back not in overlap:
            overlap.append(back)
    if len(overlap) < 2:
        raise ValueError("Number of overlap in the timestamps is too small. Did you run the evaluation on the right files?")
    n = len(traj_est)
    index = _path_index(traj_est, delta_unit)
    if not fixed_delta:
        if max_pairs == 0 or n < np.sqrt(max_pairs):
            pairs = [(i, j) for i in range(n) for j in range(n)]
        else:
            pairs = [(rng.randint(0, n - 1), rng.randint(0, n - 1)) for _ in range(max_pairs)]
    else:
        pairs = []
        for i in range(n):
            j = closest_index(index, index[i] + delta)
            if j != n - 1:
                pairs.append((i, j))
        if max_pairs != 0 and len(pairs) > max_pairs:
            pairs = rng.sample(pairs, max_pairs)
    tolerance = 2 * np.median(np.diff(gt_stamps))
    rows = []
    for i, j in pairs:
        e0, e1 = est_stamps[i], est_stamps[j]
        g0 = gt_stamps[closest_index(gt_stamps, e0 + offset)]
        g1 = gt_stamps[closest_index(gt_stamps, e1 + offset)]
        if abs(g0 - (e0 + offset)) > tolerance or abs(g1 - (e1 + offset)) > tolerance:
            continue
        step_est = _between(traj_est[e1], traj_est[e0]).copy()
        step_est[:3, 3] *= scale
        err = _between(step_est, _between(traj_gt[g1], traj_gt[g0]))
        rows.append([e0, e1, g0, g1, float(np.linalg.norm(err[:3, 3])), _angle(err)])
    if len(rows) < 2:
        raise ValueError("Couldn't find matching timestamp pairs between groundtruth and estimated trajectory!")
    return rows


def percentile(seq, q):
    """evaluate_rpe.py:299-305"""
    s = sorted(seq)
    return s[int((len(s) - 1) * q)]


def relative_error_stats(rows) -> dict:
    r = np.array(rows)
    te, re = r[:, 4], r[:, 5]
    deg = 180.0 / np.pi
    return dict(pairs=len(te), trans_rmse=float(np.sqrt(te @ te / len(te))), trans_mean=float(te.mean()), trans_median=float(np.median(te)), trans_std=float(te.std()),
                trans_min=float(te.min()), trans_max=float(te.max()), rot_rmse_deg=float(np.sqrt(re @ re / len(re)) * deg), rot_mean_deg=float(re.mean() * deg),
                rot_median_deg=float(np.median(re) * deg), rot_std_deg=float(re.std() * deg), rot_min_deg=float(re.min() * deg), rot_max_deg=float(re.max() * deg))


# ---------------------------------------------------------------------------------------------- command line
def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description="ATE / RPE of an estimated trajectory (formats and options of the reference's benchmark scripts)")
    sub = ap.add_subparsers(dest="metric", required=True)
    a = sub.add_parser("ate")
    r = sub.add_parser("rpe")
    for p in (a, r):
        p.add_argument("groundtruth_file")
        p.add_argument("estimated_file")
        p.add_argument("--offset", type=float, default=0.0)
        p.add_argument("--scale", type=float, default=1.0)
        p.add_argument("--verbose", action="store_true")
        p.add_argument("--save")
    a.add_argument("--max_difference", type=float, default=0.02)
    r.add_argument("--max_pairs", type=int, default=10000)
    r.add_argument("--fixed_delta", action="store_true")
    r.add_argument("--delta", type=float, default=1.0)
    r.add_argument("--delta_unit", default="s")
    args = ap.parse_args(argv)
    if args.metric == "ate":
        res = absolute_error(args.groundtruth_file, args.estimated_file, args.offset, args.scale, args.max_difference)
        if args.verbose:
            print("compared_pose_pairs %d pairs" % res["pairs"])
            for k in ("rmse", "mean", "median", "std", "min", "max"):
                print("absolute_translational_error.%s %f m" % (k, res[k]))
        else:
            print("%f" % res["rmse"])
        if args.save:
            second = read_stamped(args.estimated_file)
            stamps = sorted(second)
            xyz = np.array([[float(v) * args.scale for v in second[s][0:3]] for s in stamps]).T
            aligned = res["rot"] @ xyz + res["trans"]
            with open(args.save, "w") as f:
                f.write("\n".join("%f " % s + " ".join("%f" % d for d in col) for s, col in zip(stamps, aligned.T)))
        return 0
    rows = relative_errors(read_poses(args.groundtruth_file), read_poses(args.estimated_file), args.max_pairs, args.fixed_delta, args.delta, args.delta_unit,
                           args.offset, args.scale)
    st = relative_error_stats(rows)
    if args.save:
        with open(args.save, "w") as f:
            f.write("\n".join(" ".join("%f" % v for v in row) for row in rows))
    if args.verbose:
        print("compared_pose_pairs %d pairs" % st["pairs"])
        for k in ("rmse", "mean", "median", "std", "min", "max"):
            print("translational_error.%s %f m" % (k, st["trans_" + k]))
        for k in ("rmse", "mean", "median", "std", "min", "max"):
            print("rotational_error.%s %f deg" % (k, st["rot_%s_deg" % k]))
    else:
        print(st["trans_mean"])
    return 0


if __name__ == "__main__":
    sys.exit(main())
