"""Synthetic data generator invariants and the multi-rank sharding logic of bench.py (gloo, world size 2, CPU)."""
import os
import socket
import subprocess
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent


def test_scene_is_deterministic_and_well_formed(built):
    from slam_b200.synth import Scene
    a, b = Scene(), Scene()
    poses = a.trajectory(1000)
    assert np.array_equal(poses, b.trajectory(1000))
    step = np.linalg.norm(np.diff(poses[:, :3, 3], axis=0), axis=1)
    assert step.max() < 0.015, "more than 1.5 cm per frame"
    for k in (0, 333, 999):
        R = poses[k, :3, :3].astype(np.float64)
        assert np.allclose(R.T @ R, np.eye(3), atol=1e-6) and abs(np.linalg.det(R) - 1) < 1e-6
    d1, c1 = a.render_frame(poses[100])
    d2, c2 = b.render_frame(poses[100])
    assert np.array_equal(d1, d2) and np.array_equal(c1, c2)
    assert d1.dtype == np.uint16 and c1.dtype == np.uint8 and c1.shape == (480, 640, 4)
    assert c1[..., :3].min() >= 1, "intensity 0 is the tracker's 'invalid' marker"
    assert 0.0 < (d1 == 0).mean() < 0.6 and d1.max() <= 3300
    v, n, c = a.render_model(poses[100])
    valid = v[..., 2] > 0
    assert 0.5 < valid.mean() <= 1.0
    assert np.allclose(np.linalg.norm(n[valid][:, :3], axis=1), 1.0, atol=1e-5)
    both = valid & (d1 > 0)
    assert np.abs(v[..., 2][both] * 1000.0 - d1[both]).max() <= 0.51       # same surface, same camera
    # the model normal faces the camera: n . v < 0
    assert (np.einsum("ij,ij->i", n[valid][:, :3], v[valid][:, :3]) < 0).all()


def test_ate_and_rpe_definitions():
    from slam_b200.synth import ate_rmse, rpe_trans_mean
    rng = np.random.default_rng(0)
    gt = np.cumsum(rng.normal(size=(50, 3)) * 0.01, axis=0)
    # a rigidly moved copy has zero ATE (Horn alignment, benchmark/evaluate_ate.py:47-79)
    ang = 0.3
    R = np.array([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]])
    est = gt @ R.T + np.array([1.0, -2.0, 0.5])
    assert ate_rmse(gt, est) < 1e-12
    est2 = est.copy()
    est2[:, 0] += 0.01 * np.sin(np.arange(50))
    assert 0.001 < ate_rmse(gt, est2) < 0.01
    T = np.tile(np.eye(4), (10, 1, 1))
    T[:, 0, 3] = np.arange(10) * 0.1
    T2 = T.copy()
    T2[:, 0, 3] = np.arange(10) * 0.11
    assert abs(rpe_trans_mean(T, T2) - 0.01) < 1e-9


WORKER = r'''
import os, sys, json
sys.path.insert(0, sys.argv[1])
import torch, torch.distributed as dist
import numpy as np
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
import bench
# every rank owns its own sequence (weak scaling): different trajectory seeds -> different frames, no data exchange
frames, first = bench.make_frames(rank, 2)
digest = float(frames[0]["depth"].astype(np.float64).sum() + frames[1]["gt_pose"].sum())
t = torch.tensor([digest], dtype=torch.float64)
gathered = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
dist.all_gather(gathered, t)
# timing protocol: barrier, per-rank elapsed, MAX over ranks, whole-job value = world * steps / max
elapsed = torch.tensor([0.010 * (rank + 1)], dtype=torch.float64)
dist.barrier()
dist.all_reduce(elapsed, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"digests": [float(g.item()) for g in gathered], "max_elapsed": float(elapsed.item()), "value": world * 100 / float(elapsed.item())}))
dist.destroy_process_group()
'''


def test_two_rank_sharding_protocol_gloo(built, tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, OMP_NUM_THREADS="2")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port",
                          str(port), str(script), str(ROOT)], capture_output=True, text=True, timeout=280, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    res = json.loads(line)
    assert len(res["digests"]) == 2 and res["digests"][0] != res["digests"][1], "ranks must track different sequences"
    assert abs(res["max_elapsed"] - 0.020) < 1e-12 and abs(res["value"] - 2 * 100 / 0.020) < 1e-6


def test_reference_arm_prints_contract_line(built):
    import json
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "3"], capture_output=True, text=True,
                         timeout=280)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert line["impl"] == "reference" and line["unit"] == "frames/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0


def test_relocalise_key_packing_orders_like_error_then_index():
    from slam_b200.relocalise import icp_error, pack_keys, shard_range, unpack_key
    rng = np.random.default_rng(3)
    err = rng.uniform(0, 1e-2, 1000).astype(np.float32)
    err[10] = err[700]                      # a tie: the smaller index must win
    err[5] = np.inf
    keys = pack_keys(err, np.arange(1000))
    order = np.argsort(keys, kind="stable")
    want = np.lexsort((np.arange(1000), err))
    assert np.array_equal(order, want)
    e, i = unpack_key(keys.min())
    assert i == int(want[0]) and e == float(err[want[0]])
    assert unpack_key(pack_keys(np.array([np.inf], np.float32), np.array([123456]))[0]) == (float("inf"), 123456)
    # error definition: sqrt(residual) / count, +inf below the inlier threshold or for empty sets
    got = icp_error(np.array([4.0, 1.0, 0.0], np.float32), np.array([100.0, 5.0, 0.0], np.float32), min_inliers=10)
    assert got[0] == np.float32(0.02) and np.isinf(got[1]) and np.isinf(got[2])
    # shards tile the hypothesis range for any world size
    for world in (1, 2, 3, 8):
        cover = [shard_range(256, r, world) for r in range(world)]
        assert cover[0][0] == 0 and cover[-1][1] == 256 and all(cover[k][1] == cover[k + 1][0] for k in range(world - 1))


RELOC_WORKER = r'''
import json, sys
sys.path.insert(0, sys.argv[1])
import torch.distributed as dist
import numpy as np
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
from slam_b200.relocalise import score_sharded, icp_error

class FakeScorer:
    """Stands in for RGBDOdometry.score_poses on a machine without a GPU: residual / count are a deterministic function
    of the pose, identical on every rank (as the real scores are: every rank holds the same frame)."""
    def score_poses(self, level, prev_pose, trans_n, rot_n, seq=0):
        d = np.linalg.norm(np.asarray(trans_n, np.float32) - np.float32([0.1, 0.2, 0.3]), axis=1)
        count = np.where(d < 0.12, 5000.0, 3.0).astype(np.float32)      # far hypotheses lose their inliers
        return (count * (1e-3 + d) ** 2).astype(np.float32), count

rng = np.random.default_rng(11)
T = (np.float32([0.1, 0.2, 0.3]) + rng.normal(scale=0.05, size=(257, 3))).astype(np.float32)
T[200] = [0.1, 0.2, 0.3005]
T[41] = [0.1, 0.2, 0.3005]      # same error as 200: the smaller index wins on every world size
R = np.repeat(np.eye(3, dtype=np.float32)[None], 257, 0)
best, err, local = score_sharded(FakeScorer(), 2, np.eye(4, dtype=np.float32), T, R, rank, world, min_inliers=10)
res, cnt = FakeScorer().score_poses(2, None, T, R)
full = icp_error(res, cnt, 10)
if rank == 0:
    print(json.dumps({"best": best, "err": err, "want": int(np.lexsort((np.arange(257), full))[0]), "want_err": float(full.min()), "local": len(local)}))
dist.destroy_process_group()
'''


def test_two_rank_relocalisation_min_allreduce_gloo(built, tmp_path):
    import json
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "reloc_worker.py"
    script.write_text(RELOC_WORKER)
    env = dict(os.environ, OMP_NUM_THREADS="2")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port",
                          str(port), str(script), str(ROOT)], capture_output=True, text=True, timeout=280, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    res = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert res["best"] == res["want"] == 41 and res["err"] == res["want_err"]
    assert res["local"] == 128      # rank 0 scored its half of the 257 hypotheses


def test_perturbed_hypotheses_are_reproducible_and_start_with_the_ground_truth():
    """SURVEY 8(d) configs[4]: the hypothesis set is a pure function of (pose, n, seed) -- every rank builds the same one."""
    from slam_b200.relocalise import perturbed_hypotheses
    gt = np.eye(4)
    gt[:3, 3] = [1.0, -0.5, 2.0]
    T1, R1 = perturbed_hypotheses(gt, 64)
    T2, R2 = perturbed_hypotheses(gt, 64)
    assert np.array_equal(T1, T2) and np.array_equal(R1, R2)
    assert np.array_equal(T1[0], gt[:3, 3].astype(np.float32)) and np.array_equal(R1[0], np.eye(3, dtype=np.float32))
    d = np.linalg.norm(T1[1:] - T1[0], axis=1)
    assert 0.04 < d.mean() < 0.13          # sigma 5 cm per axis
    ang = np.degrees(np.arccos(np.clip((np.trace(R1[1:], axis1=1, axis2=2) - 1) / 2, -1, 1)))
    assert 2.0 < ang.mean() < 8.0          # sigma 3 deg per axis
    for r in R1:
        assert np.allclose(r @ r.T, np.eye(3), atol=1e-6)
