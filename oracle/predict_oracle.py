"""TEST INFRASTRUCTURE ONLY: ctypes front end of oracle/predict_oracle.c (part of oracle/liboracle.so).

CPU restatement of IndexMap::combinedPredict (src/model/IndexMap.cpp:243-341, splat.vert, combo_splat.frag) and the FillIn
passes (src/gl/FillIn.cpp:68-198, fill_*.frag).  PARITY UNPINNED against the reference itself (GLSL on an OpenGL context,
no golden images); see the header of predict_oracle.c for the rasterisation rules it adopts.  Only tests/, smoke() and
bench.py's cpu_baseline leg may import this module.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import cpu_oracle


class Cam(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("cx", C.c_float), ("cy", C.c_float), ("fx", C.c_float), ("fy", C.c_float), ("max_point", C.c_float)]


_bound = False


def _lib():
    global _bound
    lib = cpu_oracle.load()
    if not hasattr(lib, "predict_oracle_combined"):
        # a liboracle.so from before this file existed: rebuild in place and reload
        from slam_b200 import build
        build.build_oracle(force=True)
        cpu_oracle._lib = None
        lib = cpu_oracle.load()
    if not _bound:
        vp, i, f = C.c_void_p, C.c_int, C.c_float
        lib.predict_oracle_inverse4.argtypes = [vp, vp]
        lib.predict_oracle_combined.restype = C.c_longlong
        lib.predict_oracle_combined.argtypes = [vp, i, vp, C.POINTER(Cam), f, f, i, i, i, vp, vp, vp, vp, vp, vp]
        lib.predict_oracle_fill.argtypes = [C.POINTER(Cam), vp, vp, vp, vp, vp, i, i, vp, vp, vp]
        _bound = True
    return lib


def inverse4(pose) -> np.ndarray:
    m = np.ascontiguousarray(pose, np.float32).reshape(16)
    out = np.zeros(16, np.float32)
    _lib().predict_oracle_inverse4(m.ctypes.data, out.ctypes.data)
    return out.reshape(4, 4)


def combined_predict(surfels, pose, intr, depth_cutoff, conf_threshold, time, max_time, time_delta, max_point=2047.0, tinv=None):
    """intr = dict(width, height, cx, cy, fx, fy).  `tinv` overrides pose.inverse() (bit-exact comparisons feed both sides the same matrix)."""
    lib = _lib()
    s = np.ascontiguousarray(surfels, np.float32).reshape(-1, 12)
    W, H = intr["width"], intr["height"]
    cam = Cam(W, H, intr["cx"], intr["cy"], intr["fx"], intr["fy"], max_point)
    t = np.ascontiguousarray(inverse4(pose) if tinv is None else tinv, np.float32).reshape(16)
    out = dict(image=np.empty((H, W, 4), np.uint8), vertex=np.empty((H, W, 4), np.float32), normal=np.empty((H, W, 4), np.float32),
               time=np.empty((H, W), np.uint16), depth24=np.empty((H, W), np.uint32), winner=np.empty((H, W), np.int32))
    out["fragments"] = int(lib.predict_oracle_combined(s.ctypes.data, s.shape[0], t.ctypes.data, C.byref(cam), depth_cutoff, conf_threshold, time, max_time,
                                                        time_delta, out["image"].ctypes.data, out["vertex"].ctypes.data, out["normal"].ctypes.data,
                                                        out["time"].ctypes.data, out["depth24"].ctypes.data, out["winner"].ctypes.data))
    return out


def fill_in(intr, raw_depth, raw_rgba, vertex=None, normal=None, image=None, passthrough=False):
    """FillIn::vertex / normal / image for whichever `existing` textures are given -> dict of the filled ones."""
    lib = _lib()
    W, H = intr["width"], intr["height"]
    cam = Cam(W, H, intr["cx"], intr["cy"], intr["fx"], intr["fy"], 0.0)
    which = (1 if vertex is not None else 0) | (2 if normal is not None else 0) | (4 if image is not None else 0)
    d = np.ascontiguousarray(raw_depth, np.uint16)
    c = None if raw_rgba is None else np.ascontiguousarray(raw_rgba, np.uint8)
    ev = None if vertex is None else np.ascontiguousarray(vertex, np.float32)
    en = None if normal is None else np.ascontiguousarray(normal, np.float32)
    ei = None if image is None else np.ascontiguousarray(image, np.uint8)
    ov, on, oi = np.zeros((H, W, 4), np.float32), np.zeros((H, W, 4), np.float32), np.zeros((H, W, 4), np.uint8)
    ptr = lambda a: None if a is None else a.ctypes.data
    lib.predict_oracle_fill(C.byref(cam), ptr(ev), ptr(en), ptr(ei), d.ctypes.data, ptr(c), which, int(bool(passthrough)), ov.ctypes.data, on.ctypes.data,
                            oi.ctypes.data)
    res = {}
    if vertex is not None:
        res["vertex"] = ov
    if normal is not None:
        res["normal"] = on
    if image is not None:
        res["image"] = oi
    return res
