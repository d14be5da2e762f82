"""Untraced cost of one Gauss-Newton iteration per pyramid level and of the SO3 pre-alignment (development aid): the persistent kernel is
timed with the library's CUDA events for iteration lists that differ by one iteration at a time."""
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


def gn_us(iters, so3, icpWeight=10.0):
    odo = RGBDOdometry(intr["width"], intr["height"], intr["cx"], intr["cy"], intr["fx"], intr["fy"], iterations=iters)
    odo.initFirstRGB(first)
    fr = [odo.make_frame(d["depth"], d["rgba"], d["mv"], d["mn"], d["mrgba"], d["model_pose"], DEPTH_CUTOFF, MODEL_CUTOFF) for d in frames]
    pri = [(d["model_pose"][:3, 3].copy(), d["model_pose"][:3, :3].copy()) for d in frames]
    for i in range(16):
        odo.track_device(fr[i % NF], *pri[i % NF], icpWeight=icpWeight, so3=so3)
    odo.set_profiling(True)
    for i in range(120):
        odo.track_device(fr[i % NF], *pri[i % NF], icpWeight=icpWeight, so3=so3)
    ms, nl = odo.get_profile(reset=True)
    s = odo.stats().so3_iterations
    odo.close()
    return ms / nl * 1e3, s


base, _ = gn_us((10, 5, 4), False)
print(f"icp+rgb, no so3, 10/5/4: {base:.1f} us")
for name, it in (("level 0", (9, 5, 4)), ("level 1", (10, 4, 4)), ("level 2", (10, 5, 3))):
    t, _ = gn_us(it, False)
    print(f"  one iteration less at {name}: {t:.1f} us  -> {base - t:.2f} us per iteration")
t_so3, n = gn_us((10, 5, 4), True)
print(f"with so3: {t_so3:.1f} us (+{t_so3 - base:.1f} us, {n} so3 iterations on the last frame)")
b2, _ = gn_us((10, 5, 4), False, icpWeight=100.0)
t2, _ = gn_us((9, 5, 4), False, icpWeight=100.0)
print(f"icp only 10/5/4: {b2:.1f} us, level-0 iteration {b2 - t2:.2f} us")
t1, _ = gn_us((1, 1, 1), False)
print(f"icp+rgb 1/1/1: {t1:.1f} us (fixed cost of the launch + level set-up)")
