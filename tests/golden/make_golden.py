"""Generate tests/golden/ref_track_128x96.npz from the REFERENCE's own CUDA kernels (oracle/_ref).

Run on a GPU box (the reference kernels are CUDA):
    gpurun -- 'python tests/golden/make_golden.py gpurun_out/ref_track_128x96.npz'
then copy the file into tests/golden/.  The fixture is self-contained: it stores the synthetic inputs
(so later changes of the generator cannot invalidate it), every prepared buffer of the reference tracker,
the per-step reduction sums of one ICP+RGB+SO3 frame and the resulting pose.  The `-m "not gpu"` tests pin
the CPU oracle (oracle/odom_oracle.c) against it.

128x96 keeps the fixture small and keeps 16*cols a multiple of the 512-byte cudaMallocPitch pitch at every
level (128, 64, 32 columns), which the reference's linear indexing of corresImg needs (reduce.cu:838).
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))


def main(out_path):
    import torch
    from oracle.ref_cuda import RefOdometry
    from slam_b200.odometry import Tap
    from tests.support import frame_pair, make_scene, to_device

    W, H = 128, 96
    scene, intr = make_scene(W, H)
    poses = scene.trajectory(1000)
    k = 300
    fr = frame_pair(scene, poses, k, model_k=k - 3)      # 3 frames apart: a visible motion at this resolution
    first_rgba = scene.render_frame(poses[k - 1])[1]
    d = to_device(fr)
    d0 = torch.from_numpy(first_rgba).to("cuda:0")
    torch.cuda.synchronize()

    ref = RefOdometry(W, H, intr["cx"], intr["cy"], intr["fx"], intr["fy"])
    ref.set_trace(True)
    ref.initFirstRGB(d0)
    ref.initICPModel(d["mv"], d["mn"], 20.0, d["model_pose"])
    ref.initRGBModel(d["mrgba"])
    ref.initICP(d["depth"], 3.0)
    ref.initRGB(d["rgba"])
    out = dict(width=W, height=H, fx=intr["fx"], fy=intr["fy"], cx=intr["cx"], cy=intr["cy"], depth=fr["depth"], rgba=fr["rgba"], mv=fr["mv"], mn=fr["mn"],
               mrgba=fr["mrgba"], first_rgba=first_rgba, model_pose=fr["model_pose"], gt_pose=fr["gt_pose"])
    pre = {}
    for level in range(3):
        for name, tap in (("depth_u16", Tap.DEPTH_U16), ("vmap_curr", Tap.VMAP_CURR), ("nmap_curr", Tap.NMAP_CURR), ("vmap_prev", Tap.VMAP_PREV),
                          ("nmap_prev", Tap.NMAP_PREV), ("last_depth", Tap.LAST_DEPTH), ("next_depth", Tap.NEXT_DEPTH), ("last_image", Tap.LAST_IMAGE),
                          ("next_image", Tap.NEXT_IMAGE), ("lastnext_image", Tap.LASTNEXT_IMAGE)):
            pre[f"{name}_{level}"] = ref.tap(tap, level)
    pose = fr["model_pose"]
    t, r = ref.getIncrementalTransformation(pose[:3, 3].copy(), pose[:3, :3].copy(), False, 10.0, True, False, True)
    for level in range(3):
        for name, tap in (("didx", Tap.DIDX), ("didy", Tap.DIDY), ("cloud", Tap.CLOUD), ("corres", Tap.CORRES)):
            pre[f"{name}_{level}"] = ref.tap(tap, level)
    out.update(pre)
    tr = ref.get_trace()
    out["n_steps"] = len(tr)
    for key in ("kind", "level", "iteration", "rgb_count", "rgb_sigma", "sigma_in"):
        out["step_" + key] = np.array([s[key] for s in tr])
    for key in ("so3", "icp", "rgb", "x", "Rcurr", "tcurr", "Rcurr_in", "tcurr_in", "krkinv_in", "kt_in", "so3_in"):
        out["step_" + key] = np.stack([np.asarray(s[key]) for s in tr])
    out["trans"], out["rot"] = t, r
    st = ref.stats()
    out["stats"] = np.array([st.lastICPError, st.lastICPCount, st.lastRGBError, st.lastRGBCount, st.lastSO3Error, st.lastSO3Count], np.float32)
    out["lastA"] = np.array(st.lastA[:]).reshape(6, 6)
    out["lastb"] = np.array(st.lastb[:])
    np.savez_compressed(out_path, **out)
    print("wrote", out_path, "steps", len(tr), "final t", t, "gt t", fr["gt_pose"][:3, 3], "prior t", pose[:3, 3])


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else str(Path(__file__).resolve().parent / "ref_track_128x96.npz"))
