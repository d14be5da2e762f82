"""Shared helpers of the test-suite: synthetic frames, device upload, comparison metrics."""
from __future__ import annotations

import math
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent

ICL = dict(width=640, height=480, fx=481.20, fy=-480.0, cx=319.5, cy=239.5)
DEPTH_CUTOFF = 3.0          # src/configs/ef_iclnuim.cfg:7-30 (depthCutoff)
MODEL_CUTOFF = 20.0         # maxDepthProcessed
ANGLE_THRESH = math.sin(20.0 * 3.14159254 / 180.0)


def scaled_intrinsics(width, height):
    s = width / 640.0
    return dict(width=width, height=height, fx=481.20 * s, fy=-480.0 * s, cx=(319.5 + 0.5) * s - 0.5, cy=(239.5 + 0.5) * s - 0.5)


def make_scene(width=640, height=480, **kw):
    from slam_b200.synth import Scene
    intr = scaled_intrinsics(width, height)
    return Scene(**intr, **kw), intr


def frame_pair(scene, poses, k, model_k=None):
    """Frame k tracked against the model rendered at pose model_k (default k-1)."""
    model_k = k - 1 if model_k is None else model_k
    depth, rgba = scene.render_frame(poses[k])
    mv, mn, mrgba = scene.render_model(poses[model_k])
    return dict(depth=depth, rgba=rgba, mv=mv, mn=mn, mrgba=mrgba, model_pose=poses[model_k].copy(), gt_pose=poses[k].copy())


def to_device(fr, device="cuda:0"):
    import torch
    out = {}
    for k, v in fr.items():
        if k in ("model_pose", "gt_pose"):
            out[k] = v
        elif v.dtype == np.uint16:
            out[k] = torch.from_numpy(v.view(np.int16).copy()).to(device)   # torch has no uint16 arithmetic; raw bits suffice
        else:
            out[k] = torch.from_numpy(v).to(device)
    return out


def run_frame(odo, d, so3=True, rgbOnly=False, icpWeight=10.0, pyramid=True, fastOdom=False, first_rgb=None, pose=None):
    """The reference app's per-frame call order (apps/elastic_fusion_file.cpp:359-374)."""
    if first_rgb is not None:
        odo.initFirstRGB(first_rgb)
    odo.initICPModel(d["mv"], d["mn"], MODEL_CUTOFF, d["model_pose"])
    odo.initRGBModel(d["mrgba"])
    odo.initICP(d["depth"], DEPTH_CUTOFF)
    odo.initRGB(d["rgba"])
    pose = d["model_pose"] if pose is None else pose
    trans = pose[:3, 3].astype(np.float32).copy()
    rot = pose[:3, :3].astype(np.float32).copy()
    return odo.getIncrementalTransformation(trans, rot, rgbOnly, icpWeight, pyramid, fastOdom, so3)


# ---- comparison metrics ------------------------------------------------------------------
def nan_pattern_equal(a, b):
    return np.array_equal(np.isnan(a), np.isnan(b))


def bits_equal_where(a, b, mask):
    return np.array_equal(a.view(np.uint32)[mask], b.view(np.uint32)[mask])


def planar_map_mismatch(mine, ref):
    """Compare planar [3][h][w] maps with the reference's NaN convention (validity on the x plane only).

    Returns (#pixels whose NaN-ness differs, #valid pixels whose any component differs bitwise)."""
    nan_m, nan_r = np.isnan(mine[0]), np.isnan(ref[0])
    nan_diff = int((nan_m != nan_r).sum())
    valid = ~nan_m & ~nan_r
    diff = np.zeros_like(valid)
    for c in range(3):
        diff |= (mine[c].view(np.uint32) != ref[c].view(np.uint32)) & valid
    return nan_diff, int(diff.sum())


def se3_sums_rel_err(mine29, ref29, noise_floor=True):
    """Max error of the 27 JtJ/Jtr sums + residual, each normalised by its Cauchy-Schwarz scale sqrt(S_ii * S_jj).

    Near convergence on noise-free data the residual column (Jtr) is a sum of ~1e5 signed terms that cancel to almost
    nothing, so its fp32 rounding noise (which depends on the reduction order, free by contract) is large relative to
    sqrt(S_ii * S_rr).  With noise_floor the Jtr scale additionally admits an error worth a 1e-6 change of the solved
    increment (|db_i| <= 1e-6 * S_ii), far below the 1e-5 pose-increment bar."""
    m = np.asarray(mine29, dtype=np.float64)
    r = np.asarray(ref29, dtype=np.float64)
    # rebuild the 7x7 gram matrix (gg = residual at index 27)
    G = np.zeros((7, 7))
    Gm = np.zeros((7, 7))
    k = 0
    for i in range(7):
        for j in range(i, 7):
            G[i, j] = G[j, i] = r[k]
            Gm[i, j] = Gm[j, i] = m[k]
            k += 1
    d = np.sqrt(np.maximum(np.diag(G), 1e-30))
    scale = np.outer(d, d)
    err = np.abs(G - Gm) / scale
    if noise_floor:
        for i in range(6):
            e = abs(G[i, 6] - Gm[i, 6]) / (scale[i, 6] + 1e-2 * G[i, i])   # 1e-4 * (.. + 1e-2 S_ii) = 1e-6 S_ii
            err[i, 6] = err[6, i] = e
    return float(np.max(err))


def se3_jtj_literal_rel_err(mine29, ref29):
    """Literal element-wise relative error |mine - ref| / |ref| of the 21 JtJ terms and the residual (index 27) -- the
    north-star's "1e-4 relative" read word for word.  These sums do not cancel; the 6 Jtr terms do (near convergence they are
    sums of ~1e5 signed terms of either sign) and are judged by se3_sums_rel_err instead."""
    m = np.asarray(mine29, dtype=np.float64)
    r = np.asarray(ref29, dtype=np.float64)
    worst, k = 0.0, 0
    for i in range(7):
        for j in range(i, 7):
            if j < 6 or i == 6:
                if r[k] != 0.0:
                    worst = max(worst, abs(m[k] - r[k]) / abs(r[k]))
                else:
                    worst = max(worst, abs(m[k]))
            k += 1
    return worst


def so3_sums_rel_err(mine11, ref11):
    m = np.asarray(mine11, dtype=np.float64)
    r = np.asarray(ref11, dtype=np.float64)
    G = np.zeros((4, 4))
    Gm = np.zeros((4, 4))
    k = 0
    for i in range(4):
        for j in range(i, 4):
            G[i, j] = G[j, i] = r[k]
            Gm[i, j] = Gm[j, i] = m[k]
            k += 1
    d = np.sqrt(np.maximum(np.diag(G), 1e-30))
    scale = np.outer(d, d)
    err = np.abs(G - Gm) / scale
    for i in range(3):   # Jtr column: same noise-floor argument as se3_sums_rel_err
        e = abs(G[i, 3] - Gm[i, 3]) / (scale[i, 3] + 1e-2 * G[i, i])
        err[i, 3] = err[3, i] = e
    return float(np.max(err))


def pose_increment(R0, t0, R1, t1):
    """Relative motion between two poses as (translation, rotation-matrix difference)."""
    return np.asarray(t1, np.float64) - np.asarray(t0, np.float64), np.asarray(R1, np.float64) @ np.asarray(R0, np.float64).T
