"""Run a few tracked frames through the device-resident loop (to be wrapped in ncu)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from slam_b200 import RGBDOdometry
from tests.support import make_scene, frame_pair, to_device, run_frame

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
mode = sys.argv[2] if len(sys.argv) > 2 else "full"
scene, intr = make_scene(640, 480)
poses = scene.trajectory(1000)
frames = [to_device(frame_pair(scene, poses, 100 + 40 * i)) for i in range(4)]
first = torch.from_numpy(scene.render_frame(poses[99])[1]).to("cuda:0")
torch.cuda.synchronize()
odo = RGBDOdometry(intr["width"], intr["height"], intr["cx"], intr["cy"], intr["fx"], intr["fy"])
kw = dict(so3=True) if mode == "full" else dict(so3=False, icpWeight=100.0)
run_frame(odo, frames[0], first_rgb=first, **kw)
for i in range(n):
    run_frame(odo, frames[i % 4], **kw)
torch.cuda.synchronize()
print("done", odo.launch_count())
