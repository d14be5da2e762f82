"""Quick per-frame timing of the tracker variants (development aid, not the benchmark)."""
import sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from slam_b200 import RGBDOdometry
from oracle.ref_cuda import RefOdometry
from tests.support import make_scene, frame_pair, to_device, run_frame

scene, intr = make_scene(640, 480)
poses = scene.trajectory(1000)
NF = 16
frames = [to_device(frame_pair(scene, poses, 100 + 40 * i)) for i in range(NF)]
first = torch.from_numpy(scene.render_frame(poses[99])[1]).to("cuda:0")
torch.cuda.synchronize()
args = (intr["width"], intr["height"], intr["cx"], intr["cy"], intr["fx"], intr["fy"])

def bench(name, odo, n=200, **kw):
    run_frame(odo, frames[0], first_rgb=first, **kw)
    for i in range(20):
        run_frame(odo, frames[i % NF], **kw)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(n):
        run_frame(odo, frames[i % NF], **kw)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n
    extra = ""
    if hasattr(odo, "get_profile") and not getattr(odo, "_host_loop", False):
        odo.set_profiling(True)
        for i in range(50):
            run_frame(odo, frames[i % NF], **kw)
        ms, nl = odo.get_profile(reset=True)
        odo.set_profiling(False)
        if nl:
            extra = f"  gn kernel {ms / nl * 1e3:8.1f} us/launch"
        odo.get_phase_cycles(reset=True)
        for i in range(50):
            run_frame(odo, frames[i % NF], **kw)
        ph, nl = odo.get_phase_cycles(reset=True)
        if nl:
            extra += "\n      cycles/frame: " + ", ".join(f"{k} {v / nl:.0f}" for k, v in ph.items()) + f"  (sum {sum(ph.values()) / nl:.0f})"
    print(f"{name:40s} {dt*1e6:9.1f} us/frame  {1/dt:9.1f} fps{extra}", flush=True)

import os
ONLY_DEV = os.environ.get("QT_DEVICE_ONLY")
for kw, tag in ((dict(so3=True), "icp+rgb+so3"), (dict(so3=False, icpWeight=100.0), "icp only")):
    bench("device loop " + tag, RGBDOdometry(*args), **kw)
    if ONLY_DEV: continue
    bench("host loop   " + tag, RGBDOdometry(*args, host_loop=True), **kw)
    bench("reference   " + tag, RefOdometry(*args), n=50, **kw)
