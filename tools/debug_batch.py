import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from slam_b200 import RGBDOdometry
from tests.support import make_scene, frame_pair, to_device, run_frame
scene, intr = make_scene(640, 480)
poses = scene.trajectory(1000)
args = (intr["width"], intr["height"], intr["cx"], intr["cy"], intr["fx"], intr["fy"])
ks = (150, 420, 810, 333, 644)
frames = [frame_pair(scene, poses, k) for k in ks]
firsts = [scene.render_frame(poses[k - 1])[1] for k in ks]
def show(tr):
    for r in tr:
        if r["kind"] == 0:
            print("   so3", r["iteration"], r["so3"][9], r["so3"][10], r["x"][:3])
        else:
            err = np.sqrt(float(r["rgb_sigma"])) / max(r["rgb_count"], 1)
            print("   gn", r["level"], r["iteration"], r["rgb_count"], r["rgb_sigma"], "%.9f" % err, r["icp"][27], r["icp"][28], r["tcurr"])
for mode in ("icp+rgb+so3",):
    kw = dict(so3=(mode != "rgb_only"), rgbOnly=(mode == "rgb_only"), icpWeight=10.0, pyramid=True, fastOdom=False)
    print("==", mode)
    singles = []
    for fr, f0 in zip(frames, firsts):
        o = RGBDOdometry(*args); o.set_trace(1); d = to_device(fr); f0d = torch.from_numpy(f0).to("cuda:0"); torch.cuda.synchronize()
        t, r = run_frame(o, d, first_rgb=f0d, **kw); st = o.stats()
        print("single", t, st.lastICPCount, st.lastRGBCount, st.lastRGBError, st.gn_iterations, st.so3_iterations)
        singles.append(o.get_trace(0))
    ob = RGBDOdometry(*args, batch=5); ob.set_trace(1)
    stack = lambda key: torch.from_numpy(np.stack([(f[key].view(np.int16) if f[key].dtype == np.uint16 else f[key]) for f in frames])).to("cuda:0")
    depth, rgba, mv, mn, mrgba = (stack(k) for k in ("depth", "rgba", "mv", "mn", "mrgba"))
    P = np.stack([f["model_pose"] for f in frames]); torch.cuda.synchronize()
    f0b = torch.from_numpy(np.stack(firsts)).to("cuda:0"); torch.cuda.synchronize(); ob.initFirstRGB(f0b); ob.initICPModel(mv, mn, 20.0, P); ob.initRGBModel(mrgba); ob.initICP(depth, 3.0); ob.initRGB(rgba)
    tb, rb = ob.getIncrementalTransformation(P[:, :3, 3].copy(), P[:, :3, :3].copy(), kw["rgbOnly"], kw["icpWeight"], True, False, kw["so3"])
    for b in range(5):
        st = ob.stats(b)
        print("batch ", tb[b], st.lastICPCount, st.lastRGBCount, st.lastRGBError, st.gn_iterations, st.so3_iterations)
    for b in range(5):
        print(" seq", b, "single:"); show(singles[b])
        print(" seq", b, "batch:"); show(ob.get_trace(b))
