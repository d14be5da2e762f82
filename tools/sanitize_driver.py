"""Small workloads for compute-sanitizer (racecheck / synccheck / memcheck); see profiles/r02_sanitizer_*.log.
usage: sanitize_driver.py gn1 | gn3 | batch8 | frame1 | frame8   (the prediction / fern kernels are run through their own tests, see tools/sanitize.sh)
frame1 / frame8: the frame-level entry point (k_prepare_frame: the whole preparation in one tiled launch, then the tracker) for one / eight sequences."""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from tests.support import make_scene, frame_pair, to_device, run_frame

mode = sys.argv[1]
scene, intr = make_scene(640, 480)
poses = scene.trajectory(1000)
args = (intr["width"], intr["height"], intr["cx"], intr["cy"], intr["fx"], intr["fy"])
if mode in ("gn1", "gn3", "batch8", "frame1", "frame8"):
    from slam_b200 import RGBDOdometry
    B = {"gn1": 1, "gn3": 3, "batch8": 8, "frame1": 1, "frame8": 8}[mode]
    ks = [150 + 40 * b for b in range(B)]
    frames = [frame_pair(scene, poses, k) for k in ks]
    first = np.stack([scene.render_frame(poses[k - 1])[1] for k in ks])
    stack = lambda key: torch.from_numpy(np.stack([(f[key].view(np.int16) if f[key].dtype == np.uint16 else f[key]) for f in frames])).to("cuda:0")
    depth, rgba, mv, mn, mrgba = (stack(k) for k in ("depth", "rgba", "mv", "mn", "mrgba"))
    P = np.stack([f["model_pose"] for f in frames])
    odo = RGBDOdometry(*args, batch=B)
    odo.initFirstRGB(torch.from_numpy(first).to("cuda:0"))
    for rep in range(2):
        if mode.startswith("frame"):
            fr = odo.make_frame(depth, rgba, mv, mn, mrgba, P if B > 1 else P[0], 3.0, 20.0)
            t, r = odo.track_device(fr, P[:, :3, 3].copy(), P[:, :3, :3].copy())
            continue
        odo.initICPModel(mv, mn, 20.0, P)
        odo.initRGBModel(mrgba)
        odo.initICP(depth, 3.0)
        odo.initRGB(rgba)
        t, r = odo.getIncrementalTransformation(P[:, :3, 3].copy(), P[:, :3, :3].copy(), False, 10.0, True, False, True)
    err = [float(np.linalg.norm(np.asarray(t).reshape(B, 3)[b] - frames[b]["gt_pose"][:3, 3]) * 1e3) for b in range(B)]
    print(mode, "tracked, error mm", [round(e, 3) for e in err])
    odo.close()
