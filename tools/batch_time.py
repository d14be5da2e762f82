"""Development aid: frames/s of a batched handle at a given batch size (which engine runs is decided by the library;
SLAM_BATCH_ENGINE_MIN=<n> moves the threshold).  usage: python tools/batch_time.py B [steps]"""
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from slam_b200 import RGBDOdometry          # noqa: E402
from slam_b200.synth import Scene   # noqa: E402

W, H = 640, 480
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
scene = Scene(width=W, height=H, fx=481.20, fy=-480.0, cx=319.5, cy=239.5)
poses = scene.trajectory(12, seed=7)
frames = []
for k in range(1, 12):
    depth, rgba = scene.render_frame(poses[k])
    mv, mn, mrgba = scene.render_model(poses[k - 1])
    frames.append(dict(depth=depth.view(np.int16), rgba=rgba, mv=mv, mn=mn, mrgba=mrgba, model_pose=poses[k - 1], gt=poses[k]))
first = torch.from_numpy(scene.render_frame(poses[0])[1]).cuda()
nf = len(frames)
keys = ("depth", "rgba", "mv", "mn", "mrgba")
odo = RGBDOdometry(W, H, 319.5, 239.5, 481.20, -480.0, device=0, batch=B)
odo.initFirstRGB(torch.stack([first] * B))
sets = []
for off in (0, 1):
    d = {k: torch.stack([torch.from_numpy(frames[(off + 3 * b) % nf][k]).cuda() for b in range(B)]) for k in keys}
    P = np.stack([frames[(off + 3 * b) % nf]["model_pose"] for b in range(B)])
    G = np.stack([frames[(off + 3 * b) % nf]["gt"][:3, 3] for b in range(B)])
    sets.append((P, odo.make_frame(d["depth"], d["rgba"], d["mv"], d["mn"], d["mrgba"], P, 3.0, 20.0), G, d))
torch.cuda.synchronize()
for i in range(3):
    P, fr, G, _ = sets[i % 2]
    out = odo.track_device(fr, P[:, :3, 3].copy(), P[:, :3, :3].copy())
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(steps):
    P, fr, G, _ = sets[(3 + i) % 2]
    out = odo.track_device(fr, P[:, :3, 3].copy(), P[:, :3, :3].copy())
torch.cuda.synchronize()
dt = time.perf_counter() - t0
err = float(np.linalg.norm(out[0].reshape(B, 3) - G, axis=1).max() * 1e3)
print(f"batch {B:3d}: {dt / steps * 1e3:8.3f} ms/step  {B * steps / dt:9.1f} frames/s  worst error {err:.3f} mm  launches/step {odo.launch_count() / (steps + 3):.1f}")
