"""Trajectory evaluation (ATE / RPE of pose logs) against golden values computed by the REFERENCE's own benchmark scripts
(benchmark/associate.py, evaluate_ate.py, evaluate_rpe.py; generator: tests/golden/make_eval_golden.py, which executes their functions under
Python 3).  Parity pinned: stamps associate identically, alignment and every error statistic agree to 1e-12."""
import json
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

GOLD = Path(__file__).resolve().parent / "golden"


@pytest.fixture(scope="module")
def gold():
    return json.loads((GOLD / "eval_golden.json").read_text())


def test_association_matches_the_reference(gold):
    from slam_b200 import evaluate as ev
    first, second = ev.read_stamped(GOLD / "eval_gt.txt"), ev.read_stamped(GOLD / "eval_est.txt")
    assert len(first) == 120 and len(second) == 115
    assert [list(p) for p in ev.associate(first, second, 0.0, 0.02)] == gold["matches"]
    tight = [list(p) for p in ev.associate(first, second, 0.001, 0.003)]
    assert tight == gold["matches_offset_0.001_maxdiff_0.003"] and 0 < len(tight) < len(gold["matches"])


def test_ate_matches_the_reference(gold):
    from slam_b200 import evaluate as ev
    res = ev.absolute_error(GOLD / "eval_gt.txt", GOLD / "eval_est.txt")
    g = gold["ate"]
    assert res["pairs"] == len(g["trans_error"])
    assert np.allclose(res["rot"], g["rot"], rtol=0, atol=1e-12) and np.allclose(res["trans"].reshape(-1), g["trans"], rtol=0, atol=1e-12)
    assert np.allclose(res["trans_error"], g["trans_error"], rtol=0, atol=1e-12)
    for k in ("rmse", "mean", "median", "std", "min", "max"):
        assert abs(res[k] - g[k]) < 1e-12, k
    # the estimate is the ground truth moved rigidly + millimetre noise and drift: the alignment must remove the rigid part
    assert 5e-4 < res["rmse"] < 1e-2
    # the same metric as the helper the GPU tests use (positions already associated)
    from slam_b200.synth import ate_rmse
    first, second = ev.read_stamped(GOLD / "eval_gt.txt"), ev.read_stamped(GOLD / "eval_est.txt")
    gt = np.array([[float(v) for v in first[a][:3]] for a, _ in res["matches"]])
    est = np.array([[float(v) for v in second[b][:3]] for _, b in res["matches"]])
    assert abs(ate_rmse(gt, est) - res["rmse"]) < 1e-9


@pytest.mark.parametrize("tag,kw", [
    ("frames_1", dict(max_pairs=0, fixed_delta=True, delta=1.0, delta_unit="f")),
    ("frames_5", dict(max_pairs=0, fixed_delta=True, delta=5.0, delta_unit="f")),
    ("seconds_1", dict(max_pairs=0, fixed_delta=True, delta=1.0, delta_unit="s")),
    ("metres_0.05", dict(max_pairs=0, fixed_delta=True, delta=0.05, delta_unit="m")),
    ("degrees_1", dict(max_pairs=0, fixed_delta=True, delta=1.0, delta_unit="deg")),
    ("all_pairs_scaled", dict(max_pairs=0, fixed_delta=False, scale=1.1, offset=0.002)),
])
def test_rpe_matches_the_reference(gold, tag, kw):
    from slam_b200 import evaluate as ev
    rows = ev.relative_errors(ev.read_poses(GOLD / "eval_gt.txt"), ev.read_poses(GOLD / "eval_est.txt"), **kw)
    g = gold["rpe_" + tag]
    assert len(rows) == g["n"]
    assert np.allclose(np.array(rows[:5]), np.array(g["first_rows"]), rtol=0, atol=1e-12)
    st = ev.relative_error_stats(rows)
    for k in ("trans_rmse", "trans_mean", "trans_median", "trans_std", "trans_min", "trans_max", "rot_rmse_deg", "rot_mean_deg"):
        assert abs(st[k] - g[k]) < 1e-12, (tag, k, st[k], g[k])
    assert abs(float(np.array(rows).sum()) - g["checksum"]) < 1e-6 * max(1.0, abs(g["checksum"]))


def test_small_pieces_and_the_command_line(gold):
    from slam_b200 import evaluate as ev
    assert np.allclose(ev.pose_from_row([0.0, 1.0, 2.0, 3.0, 0.1, -0.2, 0.3, 0.9]), np.array(gold["transform44"]), rtol=0, atol=1e-15)
    assert np.array_equal(ev.pose_from_row([0, 1, 2, 3, 0, 0, 0, 0]), np.array([[1, 0, 0, 1], [0, 1, 0, 2], [0, 0, 1, 3], [0, 0, 0, 1.0]]))
    assert [ev.percentile([5, 1, 4, 2, 3, 9, 7], q) for q in (0.0, 0.5, 0.9, 1.0)] == gold["percentile"]
    assert ev.closest_index([0.0, 1.0, 2.0, 4.0], 2.9) == 2 and ev.closest_index([0.0, 1.0, 2.0, 4.0], 3.1) == 3 and ev.closest_index([5.0], -1.0) == 0
    out = subprocess.run([sys.executable, "-m", "slam_b200.evaluate", "ate", str(GOLD / "eval_gt.txt"), str(GOLD / "eval_est.txt")], capture_output=True,
                         text=True, cwd=str(GOLD.parent.parent))
    assert out.returncode == 0 and out.stdout.strip() == "%f" % gold["ate"]["rmse"], (out.stdout, out.stderr)
    out = subprocess.run([sys.executable, "-m", "slam_b200.evaluate", "rpe", str(GOLD / "eval_gt.txt"), str(GOLD / "eval_est.txt"), "--fixed_delta", "--delta_unit", "f",
                          "--max_pairs", "0", "--verbose"], capture_output=True, text=True, cwd=str(GOLD.parent.parent))
    assert out.returncode == 0 and "compared_pose_pairs %d pairs" % gold["rpe_frames_1"]["n"] in out.stdout
    assert "translational_error.rmse %f m" % gold["rpe_frames_1"]["trans_rmse"] in out.stdout


def test_pose_log_writer_round_trips_through_the_evaluation(tmp_path):
    """The pose log the tracker side writes (io.PoseLogWriter, the reference's `tick tx ty tz qx qy qz qw`) reads back as the same poses."""
    from slam_b200 import evaluate as ev
    from slam_b200.io import PoseLogWriter
    from slam_b200.synth import Scene
    poses = Scene(seed=0x51A7).trajectory(30)
    path = tmp_path / "poses.txt"
    with PoseLogWriter(path) as w:
        for k, T in enumerate(poses):
            w.write(k, T[:3, 3], T[:3, :3])
    back = ev.read_poses(path)
    assert sorted(back) == [float(k) for k in range(30)]
    for k, T in enumerate(poses):
        assert np.abs(back[float(k)] - T).max() < 2e-5
    res = ev.absolute_error(path, path)
    assert res["pairs"] == 30 and res["rmse"] < 1e-9
