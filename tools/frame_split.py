"""Where a tracked frame's wall time goes (development aid): device span (first prep launch -> end of the GN kernel) from events on the
handle's stream, the GN kernel alone from the library's own events, and the wall clock per frame of the synchronous call."""
import sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from slam_b200 import RGBDOdometry
from tests.support import make_scene, frame_pair, to_device, DEPTH_CUTOFF, MODEL_CUTOFF

scene, intr = make_scene(640, 480)
poses = scene.trajectory(1000)
NF = 16
frames = [to_device(frame_pair(scene, poses, 100 + 40 * i)) for i in range(NF)]
first = torch.from_numpy(scene.render_frame(poses[99])[1]).to("cuda:0")
torch.cuda.synchronize()
odo = RGBDOdometry(intr["width"], intr["height"], intr["cx"], intr["cy"], intr["fx"], intr["fy"])
odo.initFirstRGB(first)
fr = [odo.make_frame(d["depth"], d["rgba"], d["mv"], d["mn"], d["mrgba"], d["model_pose"], DEPTH_CUTOFF, MODEL_CUTOFF) for d in frames]
pri = [(d["model_pose"][:3, 3].copy(), d["model_pose"][:3, :3].copy()) for d in frames]
st = torch.cuda.ExternalStream(odo.stream)
for i in range(30):
    odo.track_device(fr[i % NF], *pri[i % NF])
n = 300
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(n):
    ev[i][0].record(st)
    odo.track_device(fr[i % NF], *pri[i % NF])
    ev[i][1].record(st)
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / n * 1e6
span = np.array([a.elapsed_time(b) for a, b in ev]) * 1e3
gaps = np.array([ev[i][1].elapsed_time(ev[i + 1][0]) for i in range(n - 1)]) * 1e3
odo.set_profiling(True)
for i in range(100):
    odo.track_device(fr[i % NF], *pri[i % NF])
ms, nl = odo.get_profile(reset=True)
odo.set_profiling(False)
print(f"wall {wall:.1f} us/frame | device span (events around the call) median {np.median(span):.1f} us | between calls (host) median {np.median(gaps):.1f} us | "
      f"gn kernel {ms / nl * 1e3:.1f} us")
# host cost of the call alone: enqueue without waiting is not exposed for track_device; time the python+C enqueue of the prep launches
t0 = time.perf_counter()
for i in range(n):
    odo.initICP(frames[i % NF]["depth"], DEPTH_CUTOFF)
torch.cuda.synchronize()
print(f"one async prep entry point (3 launches, python + C + launches, no wait): {(time.perf_counter() - t0) / n * 1e6:.1f} us")
# the solve alone on prepared buffers: wall - kernel = launch latency of the cooperative kernel + completion flag + python
odo.set_profiling(True)
d0 = frames[0]
tr, ro = d0["model_pose"][:3, 3].copy(), d0["model_pose"][:3, :3].copy()
for i in range(20):
    odo.getIncrementalTransformation(tr.copy(), ro.copy(), False, 10.0, True, False, False)
odo.get_profile(reset=True)
t0 = time.perf_counter()
for i in range(n):
    odo.getIncrementalTransformation(tr.copy(), ro.copy(), False, 10.0, True, False, False)
w = (time.perf_counter() - t0) / n * 1e6
ms, nl = odo.get_profile(reset=True)
print(f"solve only (no so3): wall {w:.1f} us per call, kernel {ms / nl * 1e3:.1f} us -> {w - ms / nl * 1e3:.1f} us of launch + flag + python")
