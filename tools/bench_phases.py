"""Phase cycles of the fine-level kernel and the SO3 iteration counts on bench.py's own 96 frames (development aid).
usage: SLAM_GN_PHASES=3 python tools/bench_phases.py"""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import bench
from slam_b200 import RGBDOdometry

frames, first = bench.make_frames(0, bench.N_FRAMES_DISTINCT)
up = lambda a: torch.from_numpy(a.view(np.int16) if a.dtype == np.uint16 else a).to("cuda:0")
odo = RGBDOdometry(bench.W, bench.H, 319.5, 239.5, 481.20, -480.0)
odo.initFirstRGB(up(first))
dfr = [{k: (up(v) if k not in ("model_pose", "gt_pose") else v) for k, v in fr.items()} for fr in frames]
fr = [odo.make_frame(d["depth"], d["rgba"], d["mv"], d["mn"], d["mrgba"], d["model_pose"], bench.DEPTH_CUTOFF, bench.MODEL_CUTOFF) for d in dfr]
pri = [(f["model_pose"][:3, 3].copy(), f["model_pose"][:3, :3].copy()) for f in frames]
for i in range(20):
    odo.track_device(fr[i % 96], *pri[i % 96])
odo.get_phase_cycles(reset=True)
so3 = []
for i in range(96):
    odo.track_device(fr[(20 + i) % 96], *pri[(20 + i) % 96])
    so3.append(odo.stats().so3_iterations)
ph, nl = odo.get_phase_cycles(reset=True)
print("so3 iterations per frame: mean %.2f, histogram %s" % (np.mean(so3), np.bincount(so3).tolist()))
print("cycles/frame:", ", ".join(f"{k} {v / max(nl, 1):.0f}" for k, v in ph.items()))
