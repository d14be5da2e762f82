"""Print the per-step timeline of the persistent Gauss-Newton kernel (development aid)."""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from slam_b200 import RGBDOdometry
from tests.support import make_scene, frame_pair, to_device, run_frame

mode = sys.argv[1] if len(sys.argv) > 1 else "full"
scene, intr = make_scene(640, 480)
poses = scene.trajectory(1000)
fr = to_device(frame_pair(scene, poses, 300))
first = torch.from_numpy(scene.render_frame(poses[299])[1]).to("cuda:0")
torch.cuda.synchronize()
odo = RGBDOdometry(intr["width"], intr["height"], intr["cx"], intr["cy"], intr["fx"], intr["fy"])
kw = dict(so3=True) if mode == "full" else dict(so3=False, icpWeight=100.0)
odo.set_trace(1)
run_frame(odo, fr, first_rgb=first, **kw)
for _ in range(3):
    run_frame(odo, fr, **kw)
tr = odo.get_trace()
print("kind lvl it |  begin  prep   mapA  barA  sigma  mapB  barB+fold solve (cycles, deltas)")
prev_end = None
for s in tr:
    t = s["t_cycles"]
    if s["kind"] == 0:
        continue
    d = np.diff(t)
    gap = t[0] - prev_end if prev_end is not None else 0
    prev_end = t[7]
    ts = s["t_solve"]
    sub = [ts[0] - t[6]] + list(np.diff(ts[:5])) + [t[7] - ts[4]]
    print(f"{s['kind']:4d} {s['level']:3d} {s['iteration']:2d} | gap {gap:6d} " + " ".join(f"{int(x):6d}" for x in d) + f"  total {t[7]-t[0]:7d} | solve: comb {sub[0]} elim {sub[1]} rodr {sub[2]} upd {sub[3]} prep {sub[4]} rec {sub[5]}")
