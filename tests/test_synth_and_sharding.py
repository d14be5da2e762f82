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
