"""Generates tests/golden/eval_*.txt and eval_golden.json by running the REFERENCE's own trajectory-evaluation code
(/root/reference/benchmark/associate.py, evaluate_ate.py, evaluate_rpe.py) on a small synthetic trajectory pair.

The reference scripts are Python 2.  Their functions are executed here unmodified except for the textual minimum Python 3 needs,
applied in memory (nothing is copied into the repository):
  * everything from ``if __name__`` on (argument parsing and the ``print`` statements) is cut off,
  * ``import associate`` is dropped (the module is supplied in the namespace),
  * ``dict.keys()`` results that are mutated / sorted in place are wrapped in ``list(...)``,
  * ``numpy.linalg.linalg.svd`` -> ``numpy.linalg.svd`` (the private alias is gone from NumPy 2).
Run from the repository root in the authoring container: ``python tests/golden/make_eval_golden.py``.
"""
import json
import sys
import types
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
REF = Path("/root/reference/benchmark")
sys.path.insert(0, str(ROOT))


def load(name, extra=None):
    src = (REF / f"{name}.py").read_text()
    src = src[: src.index("if __name__")]
    src = src.replace("import associate\n", "")
    src = src.replace("first_keys = first_list.keys()", "first_keys = list(first_list.keys())")
    src = src.replace("second_keys = second_list.keys()", "second_keys = list(second_list.keys())")
    src = src.replace("keys = traj.keys()", "keys = list(traj.keys())")
    src = src.replace("numpy.linalg.linalg.svd", "numpy.linalg.svd")
    mod = types.ModuleType("ref_" + name)
    if extra:
        mod.__dict__.update(extra)
    exec(compile(src, str(REF / f"{name}.py"), "exec"), mod.__dict__)
    return mod


def main():
    from slam_b200.io import quaternion_from_rotation
    from slam_b200.synth import Scene
    assoc = load("associate")
    ate = load("evaluate_ate", {"associate": assoc})
    rpe = load("evaluate_rpe")

    scene = Scene(seed=0x51A7)
    poses = scene.trajectory(1000)[100:220].astype(np.float64)
    rng = np.random.default_rng(2024)
    out = Path(__file__).resolve().parent

    def write(path, stamps, Ts):
        with open(path, "w") as f:
            f.write("# timestamp tx ty tz qx qy qz qw\n")
            for s, T in zip(stamps, Ts):
                q = quaternion_from_rotation(T[:3, :3])
                f.write("%.6f %.9f %.9f %.9f %.9f %.9f %.9f %.9f\n" % (s, *T[:3, 3], *q))

    stamps_gt = 1000.0 + np.arange(len(poses)) / 30.0
    # estimate: every ground-truth frame but five, time stamps jittered by a few ms, a rigid offset + drift + noise on the poses
    keep = np.ones(len(poses), bool)
    keep[[7, 33, 34, 80, 119]] = False
    stamps_est = (stamps_gt + rng.uniform(-0.004, 0.004, len(poses)))[keep]
    ang = np.deg2rad(4.0)
    off = np.eye(4)
    off[:3, :3] = [[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]]
    off[:3, 3] = [0.3, -0.1, 0.2]
    est = []
    for k, T in enumerate(poses):
        d = np.eye(4)
        a = np.deg2rad(0.02) * k + rng.normal(0, 2e-4)
        d[:3, :3] = [[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]]
        d[:3, 3] = rng.normal(0, 1.5e-3, 3) + 2e-5 * k
        est.append(off @ T @ d)
    est = np.array(est)[keep]
    write(out / "eval_gt.txt", stamps_gt, poses)
    write(out / "eval_est.txt", stamps_est, est)

    gold = {}
    first, second = assoc.read_file_list(str(out / "eval_gt.txt")), assoc.read_file_list(str(out / "eval_est.txt"))
    matches = assoc.associate(first, second, 0.0, 0.02)
    gold["matches"] = [[a, b] for a, b in matches]
    m2 = assoc.associate(first, second, 0.001, 0.003)
    gold["matches_offset_0.001_maxdiff_0.003"] = [[a, b] for a, b in m2]
    first_xyz = np.matrix([[float(v) for v in first[a][0:3]] for a, b in matches]).transpose()
    second_xyz = np.matrix([[float(v) for v in second[b][0:3]] for a, b in matches]).transpose()
    rot, trans, err = ate.align(second_xyz, first_xyz)
    gold["ate"] = dict(rot=np.asarray(rot).tolist(), trans=np.asarray(trans).reshape(-1).tolist(), trans_error=err.tolist(),
                       rmse=float(np.sqrt(np.dot(err, err) / len(err))), mean=float(np.mean(err)), median=float(np.median(err)), std=float(np.std(err)),
                       min=float(np.min(err)), max=float(np.max(err)))
    tg, te = rpe.read_trajectory(str(out / "eval_gt.txt")), rpe.read_trajectory(str(out / "eval_est.txt"))
    for tag, kw in (("frames_1", dict(param_max_pairs=0, param_fixed_delta=True, param_delta=1.0, param_delta_unit="f")),
                    ("frames_5", dict(param_max_pairs=0, param_fixed_delta=True, param_delta=5.0, param_delta_unit="f")),
                    ("seconds_1", dict(param_max_pairs=0, param_fixed_delta=True, param_delta=1.0, param_delta_unit="s")),
                    ("metres_0.05", dict(param_max_pairs=0, param_fixed_delta=True, param_delta=0.05, param_delta_unit="m")),
                    ("degrees_1", dict(param_max_pairs=0, param_fixed_delta=True, param_delta=1.0, param_delta_unit="deg")),
                    ("all_pairs_scaled", dict(param_max_pairs=0, param_fixed_delta=False, param_scale=1.1, param_offset=0.002))):
        res = np.array(rpe.evaluate_trajectory(tg, te, **kw))
        te_, re_ = res[:, 4], res[:, 5]
        gold["rpe_" + tag] = dict(n=int(len(res)), first_rows=res[:5].tolist(), trans_rmse=float(np.sqrt(np.dot(te_, te_) / len(te_))), trans_mean=float(np.mean(te_)),
                                  trans_median=float(np.median(te_)), trans_std=float(np.std(te_)), trans_min=float(np.min(te_)), trans_max=float(np.max(te_)),
                                  rot_rmse_deg=float(np.sqrt(np.dot(re_, re_) / len(re_)) * 180.0 / np.pi), rot_mean_deg=float(np.mean(re_) * 180.0 / np.pi),
                                  checksum=float(res.sum()))
    gold["transform44"] = rpe.transform44([0.0, 1.0, 2.0, 3.0, 0.1, -0.2, 0.3, 0.9]).tolist()
    gold["percentile"] = [rpe.percentile([5, 1, 4, 2, 3, 9, 7], q) for q in (0.0, 0.5, 0.9, 1.0)]
    (out / "eval_golden.json").write_text(json.dumps(gold))
    print("wrote", out / "eval_golden.json", {k: (v["n"] if isinstance(v, dict) and "n" in v else "") for k, v in gold.items()})


if __name__ == "__main__":
    main()
