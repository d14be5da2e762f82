"""Per-phase SM cycles of ONE Gauss-Newton iteration of a pyramid level (development aid): the phase accounting of the fine-level kernel
(SLAM_GN_PHASES=3) for two iteration lists that differ only at that level, subtracted.  usage: SLAM_GN_PHASES=3 python tools/level_phases.py"""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from slam_b200 import RGBDOdometry
from tests.support import make_scene, frame_pair, to_device, DEPTH_CUTOFF, MODEL_CUTOFF

scene, intr = make_scene(640, 480)
poses = scene.trajectory(1000)
NF = 8
frames = [to_device(frame_pair(scene, poses, 100 + 40 * i)) for i in range(NF)]
first = torch.from_numpy(scene.render_frame(poses[99])[1]).to("cuda:0")
torch.cuda.synchronize()


def phases(iters, so3=True):
    odo = RGBDOdometry(intr["width"], intr["height"], intr["cx"], intr["cy"], intr["fx"], intr["fy"], iterations=iters)
    odo.initFirstRGB(first)
    fr = [odo.make_frame(d["depth"], d["rgba"], d["mv"], d["mn"], d["mrgba"], d["model_pose"], DEPTH_CUTOFF, MODEL_CUTOFF) for d in frames]
    pri = [(d["model_pose"][:3, 3].copy(), d["model_pose"][:3, :3].copy()) for d in frames]
    for i in range(16):
        odo.track_device(fr[i % NF], *pri[i % NF], so3=so3)
    odo.get_phase_cycles(reset=True)
    for i in range(80):
        odo.track_device(fr[i % NF], *pri[i % NF], so3=so3)
    ph, nl = odo.get_phase_cycles(reset=True)
    odo.close()
    return {k: v / max(nl, 1) for k, v in ph.items()}


base = phases((10, 5, 4))
for name, it, n in (("level 0", (2, 5, 4), 8), ("level 1", (10, 1, 4), 4), ("level 2", (10, 5, 1), 3)):
    p = phases(it)
    keys = ["step set-up", "rgb assoc", "icp map", "icp reduce + count wait", "rgb products+reduce", "sums wait", "solve", "end barrier"]
    print(name, "cycles per iteration:", ", ".join(f"{k} {(base[k] - p[k]) / n:.0f}" for k in keys), f" (sum {sum((base[k] - p[k]) / n for k in keys):.0f})")
print("whole frame:", ", ".join(f"{k} {v:.0f}" for k, v in base.items()))
