"""Literal element-wise relative error of the JtJ / Jtr sums against the reference replay (development aid for the tolerance
statement in DESIGN.md section 5)."""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from slam_b200 import RGBDOdometry
from oracle.ref_cuda import RefOdometry
from tests.support import make_scene, frame_pair, to_device, run_frame

scene, intr = make_scene(640, 480)
poses = scene.trajectory(1000)
args = (intr["width"], intr["height"], intr["cx"], intr["cy"], intr["fx"], intr["fy"])
worst = {"icp diag": 0, "icp offdiag": 0, "icp jtr": 0, "rgb diag": 0, "rgb offdiag": 0, "rgb jtr": 0}
idx = {}
k = 0
for i in range(7):
    for j in range(i, 7):
        idx[k] = (i, j)
        k += 1
for frame in (120, 300, 640):
    first = torch.from_numpy(scene.render_frame(poses[frame - 1])[1]).to("cuda:0")
    d = to_device(frame_pair(scene, poses, frame))
    for host_loop in (True, False):
        mine, ref = RGBDOdometry(*args, host_loop=host_loop), RefOdometry(*args)
        mine.set_trace(True); ref.set_trace(True)
        run_frame(mine, d, first_rgb=first, so3=False); run_frame(ref, d, first_rgb=first, so3=False)
        a, b = mine.get_trace()[0], ref.get_trace()[0]      # first step: bit-identical inputs
        for key in ("icp", "rgb"):
            for q in range(28):
                i, j = idx[q]
                rel = abs(float(a[key][q]) - float(b[key][q])) / max(abs(float(b[key][q])), 1e-30)
                kind = "jtr" if j == 6 and i != 6 else ("diag" if i == j else "offdiag")
                worst[f"{key} {kind}"] = max(worst[f"{key} {kind}"], rel)
        mine.close(); ref.close()
print({k: f"{v:.2e}" for k, v in worst.items()})
