"""Synthetic ICL-NUIM-shaped RGB-D sequences (ctypes front end of slam_b200/synth/synth.c).

Data for tests and bench.py only (``"data": "synthetic"``): an analytic living-room scene ray-cast
along a smooth orbit, in exactly the tracker's input formats.  Camera intrinsics default to the
reference's ICL-NUIM configuration (src/configs/ef_iclnuim.cfg:27-30, note the negative fy).
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

_PKG = Path(__file__).resolve().parent

ICL_NUIM = dict(width=640, height=480, fx=481.20, fy=-480.0, cx=319.5, cy=239.5)
REALSENSE_720P = dict(width=1280, height=720, fx=910.0, fy=910.0, cx=639.5, cy=359.5)


class _Cfg(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
                ("seed", C.c_uint64), ("depth_max", C.c_float), ("model_max", C.c_float), ("noise_mm", C.c_float)]


_lib = None


def _load():
    global _lib
    if _lib is None:
        path = _PKG / "libslam_synth.so"
        if not path.exists():
            from . import build
            build.build_synth()
        lib = C.CDLL(str(path))
        lib.synth_create.restype = C.c_void_p
        lib.synth_create.argtypes = [C.POINTER(_Cfg)]
        lib.synth_destroy.argtypes = [C.c_void_p]
        lib.synth_render_frame.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
        lib.synth_render_model.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.synth_trajectory.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_void_p]
        _lib = lib
    return _lib


class Scene:
    def __init__(self, width=640, height=480, fx=481.20, fy=-480.0, cx=319.5, cy=239.5, seed=0x51A7, depth_max=3.3, model_max=3.5, noise_mm=0.0):
        self.lib = _load()
        self.width, self.height = width, height
        self.fx, self.fy, self.cx, self.cy = fx, fy, cx, cy
        cfg = _Cfg(width, height, fx, fy, cx, cy, seed, depth_max, model_max, noise_mm)
        self._h = C.c_void_p(self.lib.synth_create(C.byref(cfg)))

    def __del__(self):
        if getattr(self, "_h", None):
            self.lib.synth_destroy(self._h)
            self._h = None

    def trajectory(self, n_frames: int, seed: int = 0x51A7) -> np.ndarray:
        poses = np.zeros((n_frames, 4, 4), dtype=np.float32)
        self.lib.synth_trajectory(self._h, n_frames, seed, poses.ctypes.data)
        return poses

    def render_frame(self, pose, noise_seed: int = 0, out=None):
        """-> depth uint16 (H,W) in mm, rgba uint8 (H,W,4)."""
        pose = np.ascontiguousarray(pose, dtype=np.float32)
        if out is None:
            depth = np.empty((self.height, self.width), dtype=np.uint16)
            rgba = np.empty((self.height, self.width, 4), dtype=np.uint8)
        else:
            depth, rgba = out
        self.lib.synth_render_frame(self._h, pose.ctypes.data, depth.ctypes.data, rgba.ctypes.data, noise_seed)
        return depth, rgba

    def render_model(self, pose, out=None):
        """-> vertices float32 (H,W,4), normals float32 (H,W,4) in the camera frame of `pose`, rgba uint8 (H,W,4)."""
        pose = np.ascontiguousarray(pose, dtype=np.float32)
        if out is None:
            v = np.empty((self.height, self.width, 4), dtype=np.float32)
            n = np.empty((self.height, self.width, 4), dtype=np.float32)
            rgba = np.empty((self.height, self.width, 4), dtype=np.uint8)
        else:
            v, n, rgba = out
        self.lib.synth_render_model(self._h, pose.ctypes.data, v.ctypes.data, n.ctypes.data, rgba.ctypes.data)
        return v, n, rgba


def surfels_from_frame(scene: "Scene", pose, time=1, stride=1, conf=None, seed=0) -> np.ndarray:
    """A surfel model (count x 12 float32: position+confidence, colour+times, normal+radius -- the reference's vertex buffer layout,
    src/gl/Vertex.cpp) made from one rendered view, with the formulas the reference initialises surfels with
    (src/model/shaders/data.vert:79-107: encodeColor, confidence; surfels.glsl:19-46: getRadius).  Test / bench data only."""
    v, n, rgba = scene.render_model(pose)
    ys, xs = np.mgrid[0:scene.height:stride, 0:scene.width:stride]
    v, n, rgba = v[ys, xs].reshape(-1, 4), n[ys, xs].reshape(-1, 4), rgba[ys, xs].reshape(-1, 4)
    ok = (v[:, 2] > 0) & (np.abs(n[:, 2]) > 1e-3)
    v, n, rgba, xs, ys = v[ok], n[ok], rgba[ok], xs.reshape(-1)[ok], ys.reshape(-1)[ok]
    pose = np.asarray(pose, dtype=np.float32)
    R, t = pose[:3, :3], pose[:3, 3]
    s = np.zeros((len(v), 12), np.float32)
    s[:, 0:3] = v[:, :3] @ R.T + t
    rad = np.hypot(xs - scene.cx, ys - scene.cy) / 400.0
    s[:, 3] = np.exp(-(rad * rad) / 0.72) * 25.0 if conf is None else conf          # confidence(x, y, weighting)
    rgb = rgba[:, :3].astype(np.int64)
    s[:, 4] = ((rgb[:, 0] << 16) + (rgb[:, 1] << 8) + rgb[:, 2]).astype(np.float32)  # encodeColor
    s[:, 6] = time                                                                  # initialisation time
    s[:, 7] = time                                                                  # last update
    s[:, 8:11] = n[:, :3] @ R.T
    mean_focal = (abs(scene.fx) + abs(scene.fy)) / 2.0
    radius = v[:, 2] / mean_focal * 1.41421356237 * stride
    s[:, 11] = np.minimum(2.0 * radius, radius / np.abs(n[:, 2]))                   # getRadius
    rng = np.random.default_rng(seed)
    return s[rng.permutation(len(s))]          # draw order must not matter for anything but exact depth ties


# ---- trajectory error metrics (definitions of the reference's benchmark scripts) --------------
def ate_rmse(gt_xyz: np.ndarray, est_xyz: np.ndarray) -> float:
    """Absolute trajectory error: RMSE of translations after Horn alignment (benchmark/evaluate_ate.py:47-79,162)."""
    gt = np.asarray(gt_xyz, dtype=np.float64).T
    est = np.asarray(est_xyz, dtype=np.float64).T
    gz = gt - gt.mean(1, keepdims=True)
    ez = est - est.mean(1, keepdims=True)
    W = np.zeros((3, 3))
    for c in range(gt.shape[1]):
        W += np.outer(ez[:, c], gz[:, c])
    U, d, Vh = np.linalg.svd(W.T)
    S = np.eye(3)
    if np.linalg.det(U) * np.linalg.det(Vh) < 0:
        S[2, 2] = -1
    rot = U @ S @ Vh
    trans = gt.mean(1, keepdims=True) - rot @ est.mean(1, keepdims=True)
    err = rot @ est + trans - gt
    return float(np.sqrt((err * err).sum(0).mean()))


def rpe_trans_mean(gt_T: np.ndarray, est_T: np.ndarray, delta: int = 1) -> float:
    """Relative pose error, mean translational part over all pairs (i, i+delta) (benchmark/evaluate_rpe.py:204-296,367)."""
    errs = []
    for i in range(len(gt_T) - delta):
        dg = np.linalg.inv(gt_T[i].astype(np.float64)) @ gt_T[i + delta].astype(np.float64)
        de = np.linalg.inv(est_T[i].astype(np.float64)) @ est_T[i + delta].astype(np.float64)
        e = np.linalg.inv(dg) @ de
        errs.append(np.linalg.norm(e[:3, 3]))
    return float(np.mean(errs))
