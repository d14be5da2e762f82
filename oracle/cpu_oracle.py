"""TEST INFRASTRUCTURE ONLY: ctypes front end of oracle/liboracle.so (oracle/odom_oracle.c).

CPU restatement of the reference's tracking path.  Only tests/, __graft_entry__.smoke() and
bench.py (cpu_baseline / --impl reference) may import this module; the product never does.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "liboracle.so"
_lib = None


class Intr(C.Structure):
    _fields_ = [("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float)]


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        import sys
        sys.path.insert(0, str(_HERE.parent))
        from slam_b200 import build
        build.build_oracle()
    lib = C.CDLL(str(LIB_PATH))
    vp, i, f = C.c_void_p, C.c_int, C.c_float
    lib.oracle_num_threads.restype = i
    lib.oracle_depth_bilateral.argtypes = [vp, i, i, f, vp]
    lib.oracle_create.restype = vp
    lib.oracle_create.argtypes = [i, i, f, f, f, f, f, f]
    lib.oracle_destroy.argtypes = [vp]
    lib.oracle_init_icp_depth.argtypes = [vp, vp, f]
    lib.oracle_init_icp_maps.argtypes = [vp, vp, vp, f]
    lib.oracle_init_icp_model.argtypes = [vp, vp, vp, f, vp]
    for n in ("oracle_init_rgb", "oracle_init_rgb_model", "oracle_init_first_rgb"):
        getattr(lib, n).argtypes = [vp, vp]
    lib.oracle_get_incremental_transformation.argtypes = [vp, vp, vp, i, f, i, i, i]
    lib.oracle_get_stats.argtypes = [vp, vp, vp, vp, vp]
    lib.oracle_buffer.restype = vp
    lib.oracle_buffer.argtypes = [vp, i, i]
    lib.oracle_pyr_down.argtypes = [vp, i, i, vp]
    lib.oracle_create_vmap.argtypes = [Intr, vp, i, i, vp, f]
    lib.oracle_create_nmap.argtypes = [vp, i, i, vp]
    lib.oracle_copy_maps.argtypes = [vp, vp, i, i, vp, vp]
    lib.oracle_resize_map.argtypes = [vp, i, i, vp, i]
    lib.oracle_transform_maps.argtypes = [vp, vp, i, i, vp, vp]
    lib.oracle_vertices_to_depth.argtypes = [vp, i, vp, f]
    lib.oracle_pyr_down_gauss_f.argtypes = [vp, i, i, vp]
    lib.oracle_pyr_down_gauss_u8.argtypes = [vp, i, i, vp]
    lib.oracle_bgr_to_intensity.argtypes = [vp, i, vp]
    lib.oracle_derivatives.argtypes = [vp, i, i, vp, vp]
    lib.oracle_project_points.argtypes = [vp, i, i, vp, Intr]
    lib.oracle_icp_step.argtypes = [vp, vp, vp, vp, vp, vp, Intr, vp, vp, f, f, i, i, vp, vp]
    lib.oracle_rgb_residual.argtypes = [f, vp, vp, vp, vp, vp, vp, vp, f, vp, vp, i, i, vp]
    lib.oracle_rgb_step.argtypes = [vp, f, vp, f, f, vp, vp, f, i, i, vp]
    lib.oracle_so3_step.argtypes = [vp, vp, vp, vp, vp, i, i, vp]
    _lib = lib
    return lib


def _p(a: np.ndarray):
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data


_TAP_DTYPE = {0: np.uint16, 1: np.float32, 2: np.float32, 3: np.float32, 4: np.float32, 5: np.float32, 6: np.float32, 7: np.uint8, 8: np.uint8,
              9: np.uint8, 10: np.int16, 11: np.int16, 12: np.float32, 13: np.uint8}
_TAP_ELEMS = {0: 1, 1: 3, 2: 3, 3: 3, 4: 3, 5: 1, 6: 1, 7: 1, 8: 1, 9: 1, 10: 1, 11: 1, 12: 3, 13: 16}


class CpuOdometry:
    """The CPU port behind the same Python surface as slam_b200.RGBDOdometry (host numpy arrays instead of device pointers)."""

    def __init__(self, width, height, cx, cy, fx, fy, distThresh=0.0, angleThresh=0.0):
        self.lib = load()
        self.width, self.height = width, height
        self._h = C.c_void_p(self.lib.oracle_create(width, height, cx, cy, fx, fy, distThresh, angleThresh))

    def close(self):
        if getattr(self, "_h", None):
            self.lib.oracle_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def initICP(self, a, depthCutoff, normals=None):
        if normals is None:
            self.lib.oracle_init_icp_depth(self._h, _p(np.ascontiguousarray(a)), depthCutoff)
        else:
            self.lib.oracle_init_icp_maps(self._h, _p(np.ascontiguousarray(a)), _p(np.ascontiguousarray(normals)), depthCutoff)

    def initICPModel(self, v, n, depthCutoff, modelPose):
        pose = np.ascontiguousarray(modelPose, dtype=np.float32).reshape(-1)
        self.lib.oracle_init_icp_model(self._h, _p(np.ascontiguousarray(v)), _p(np.ascontiguousarray(n)), depthCutoff, _p(pose))

    def initRGB(self, rgb): self.lib.oracle_init_rgb(self._h, _p(np.ascontiguousarray(rgb)))
    def initRGBModel(self, rgb): self.lib.oracle_init_rgb_model(self._h, _p(np.ascontiguousarray(rgb)))
    def initFirstRGB(self, rgb): self.lib.oracle_init_first_rgb(self._h, _p(np.ascontiguousarray(rgb)))

    def getIncrementalTransformation(self, trans, rot, rgbOnly, icpWeight, pyramid, fastOdom, so3):
        t = np.ascontiguousarray(trans, dtype=np.float32).reshape(-1).copy()
        r = np.ascontiguousarray(rot, dtype=np.float32).reshape(-1).copy()
        self.lib.oracle_get_incremental_transformation(self._h, _p(t), _p(r), int(bool(rgbOnly)), float(icpWeight), int(bool(pyramid)), int(bool(fastOdom)),
                                                       int(bool(so3)))
        return t.reshape(np.shape(trans)), r.reshape(np.shape(rot))

    def stats(self):
        six = np.zeros(6, np.float32)
        A = np.zeros(36, np.float64)
        b = np.zeros(6, np.float64)
        it = np.zeros(2, np.int32)
        self.lib.oracle_get_stats(self._h, _p(six), _p(A), _p(b), _p(it))
        return dict(lastICPError=six[0], lastICPCount=six[1], lastRGBError=six[2], lastRGBCount=six[3], lastSO3Error=six[4], lastSO3Count=six[5],
                    lastA=A.reshape(6, 6), lastb=b, so3_iterations=int(it[0]), gn_iterations=int(it[1]))

    def tap(self, tap, level):
        from slam_b200.odometry import shape_tap
        h, w = self.height >> level, self.width >> level
        ptr = self.lib.oracle_buffer(self._h, tap, level)
        n = h * w * _TAP_ELEMS[tap]
        dt = np.dtype(_TAP_DTYPE[tap])
        buf = np.frombuffer((C.c_uint8 * (n * dt.itemsize)).from_address(ptr), dtype=np.uint8).copy()
        return shape_tap(buf, tap, h, w)


def depth_bilateral(depth_u16: np.ndarray, max_depth_m: float) -> np.ndarray:
    """oracle/depth_filter_oracle.c: the reference's 13x13 bilateral depth pre-filter (gl/shaders/depth_bilateral.frag)."""
    lib = load()
    src = np.ascontiguousarray(depth_u16, dtype=np.uint16)
    dst = np.zeros_like(src)
    lib.oracle_depth_bilateral(src.ctypes.data, src.shape[0], src.shape[1], float(max_depth_m), dst.ctypes.data)
    return dst
