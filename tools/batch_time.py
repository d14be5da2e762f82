"""Throughput of the tracker for batches of independent sequences on one GPU (development aid)."""
import os, sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from slam_b200 import RGBDOdometry
from tests.support import make_scene, frame_pair

scene, intr = make_scene(640, 480)
poses = scene.trajectory(1000)
NF = 8
frames = [frame_pair(scene, poses, 100 + 90 * i) for i in range(NF)]
first = scene.render_frame(poses[99])[1]
args = (intr["width"], intr["height"], intr["cx"], intr["cy"], intr["fx"], intr["fy"])

def stack(key, B, off):
    arrs = [frames[(off + b) % NF][key] for b in range(B)]
    a = np.stack([(x.view(np.int16) if x.dtype == np.uint16 else x) for x in arrs])
    return torch.from_numpy(a).to("cuda:0")

for B in [int(x) for x in (sys.argv[1:] or ["1", "2", "4", "8", "16", "32", "64"])]:
    odo = RGBDOdometry(*args, batch=B)
    sets = []
    for off in range(2):
        d = {k: stack(k, B, off) for k in ("depth", "rgba", "mv", "mn", "mrgba")}
        P = np.stack([frames[(off + b) % NF]["model_pose"] for b in range(B)])
        fr = odo.make_frame(d["depth"], d["rgba"], d["mv"], d["mn"], d["mrgba"], P, 3.0, 20.0)
        sets.append((d, P, fr))
    firstB = torch.from_numpy(np.stack([first] * B)).to("cuda:0")
    torch.cuda.synchronize()
    odo.initFirstRGB(firstB)
    def step(i):
        d, P, fr = sets[i % 2]
        m = os.environ.get('SLAM_MODE', 'full')
        kw = dict(icpWeight=100.0, so3=False) if m == 'icp' else dict(rgbOnly=True, so3=False) if m == 'rgb' else {}
        return odo.track_device(fr, P[:, :3, 3].copy(), P[:, :3, :3].copy(), **kw)
    for i in range(4):
        out = step(i)
    torch.cuda.synchronize()
    n = max(4, 64 // B)
    odo.set_profiling(True); odo.get_profile(reset=True)
    t0 = time.perf_counter()
    for i in range(n):
        out = step(i)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    ms, nl = odo.get_profile(reset=True)
    gt = np.stack([frames[((n - 1) % 2 + b) % NF]["gt_pose"][:3, 3] for b in range(B)])
    err = np.linalg.norm(out[0].reshape(B, 3) - gt, axis=1).max() * 1e3
    print(f"batch {B:3d}: {n * B / dt:9.1f} frames/s   {dt / n * 1e3:8.3f} ms per batched step   gn kernel {ms / max(nl,1):8.3f} ms  max err {err:.2f} mm", flush=True)
    odo.close()
