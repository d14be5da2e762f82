"""TEST INFRASTRUCTURE ONLY: ctypes front end of oracle/_ref/libslam_ref.so.

That library is the REFERENCE's own CUDA code (src/odom/reduce.cu, src/odom/utils.cu,
src/cuda/containers/device_memory.cpp compiled unmodified for sm_100a, see slam_b200/build.py)
plus oracle/ref_harness.cu, which replays src/odom/RGBDOdometryef.cpp call-for-call.  It is the
parity oracle of the `-m gpu` tests and the "reference's own CUDA path on B200" baseline of
bench.py.  Only tests/, __graft_entry__.smoke() and bench.py may import this module.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

from slam_b200.odometry import Stats, StepRecord, shape_tap

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "_ref" / "libslam_ref.so"
_lib = None


def available() -> bool:
    return LIB_PATH.exists()


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(f"{LIB_PATH} missing: build it where /root/reference exists (python -m slam_b200.build)")
    lib = C.CDLL(str(LIB_PATH))
    vp, fp, i, f = C.c_void_p, C.POINTER(C.c_float), C.c_int, C.c_float
    lib.ref_odom_create.restype = vp
    lib.ref_odom_create.argtypes = [i, i, f, f, f, f, f, f]
    lib.ref_odom_destroy.argtypes = [vp]
    lib.ref_odom_create_levels.restype = vp
    lib.ref_odom_create_levels.argtypes = [i, i, f, f, f, f, f, f, i]
    lib.ref_odom_set_iterations4.argtypes = [vp, i, i, i, i]
    lib.ref_odom_set_iterations.argtypes = [vp, i, i, i]
    lib.ref_odom_init_icp_depth.argtypes = [vp, vp, f]
    lib.ref_odom_init_icp_maps.argtypes = [vp, vp, vp, f]
    lib.ref_odom_init_icp_model.argtypes = [vp, vp, vp, f, fp]
    for n in ("ref_odom_init_rgb", "ref_odom_init_rgb_model", "ref_odom_init_first_rgb"):
        getattr(lib, n).argtypes = [vp, vp]
    lib.ref_odom_get_incremental_transformation.argtypes = [vp, fp, fp, i, f, i, i, i]
    lib.ref_odom_get_stats.argtypes = [vp, C.POINTER(Stats)]
    lib.ref_odom_set_trace.argtypes = [vp, i]
    lib.ref_odom_get_trace.argtypes = [vp, C.POINTER(StepRecord), i]
    lib.ref_odom_tap.argtypes = [vp, i, i, vp]
    lib.ref_op_pyr_down.argtypes = [vp, i, i, vp]
    lib.ref_op_create_vmap.argtypes = [f, f, f, f, vp, i, i, vp, f, i]
    lib.ref_op_create_nmap.argtypes = [vp, i, i, vp, i]
    lib.ref_op_transform_maps.argtypes = [vp, vp, i, i, fp, fp, vp, vp]
    lib.ref_op_copy_maps.argtypes = [vp, vp, i, i, vp, vp]
    lib.ref_op_resize_map.argtypes = [vp, i, i, vp, i]
    lib.ref_op_image_bgr_to_intensity.argtypes = [vp, i, i, vp]
    lib.ref_op_vertices_to_depth.argtypes = [vp, i, i, vp, f]
    lib.ref_op_project_to_point_cloud.argtypes = [vp, i, i, vp, f, f, f, f, i]
    lib.ref_op_pyr_down_gauss_f.argtypes = [vp, i, i, vp]
    lib.ref_op_pyr_down_uchar_gauss.argtypes = [vp, i, i, vp]
    lib.ref_op_compute_derivative_images.argtypes = [vp, i, i, vp, vp]
    lib.ref_op_icp_step.argtypes = [fp, fp, vp, vp, fp, fp, f, f, f, f, vp, vp, f, f, i, i, fp]
    lib.ref_op_compute_rgb_residual.argtypes = [f, vp, vp, vp, vp, vp, vp, vp, f, fp, fp, i, i, C.POINTER(C.c_int)]
    lib.ref_op_rgb_step.argtypes = [vp, f, vp, f, f, vp, vp, f, i, i, fp]
    lib.ref_op_so3_step.argtypes = [vp, vp, fp, fp, fp, i, i, fp]
    _lib = lib
    return lib


def _fptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _addr(x) -> int:
    return int(x.data_ptr()) if hasattr(x, "data_ptr") else int(x)


class RefOdometry:
    """The reference tracker (its own kernels) behind the same Python surface as slam_b200.RGBDOdometry."""

    def __init__(self, width, height, cx, cy, fx, fy, distThresh=0.0, angleThresh=0.0, iterations=None, num_levels=3):
        """num_levels = 4: the replay loops over four levels (the class itself hard-codes three, RGBDOdometryef.h:104; every
        wrapper it calls is level-agnostic, odom/utils.cuh:62-175)."""
        self.lib = load()
        self.width, self.height = width, height
        if num_levels == 3:
            self._h = C.c_void_p(self.lib.ref_odom_create(width, height, cx, cy, fx, fy, distThresh, angleThresh))
        else:
            self._h = C.c_void_p(self.lib.ref_odom_create_levels(width, height, cx, cy, fx, fy, distThresh, angleThresh, int(num_levels)))
            assert self._h, "reference replay: unsupported level count"
        if iterations:
            it = [int(v) for v in iterations] + [0] * 4
            if num_levels == 3:
                self.lib.ref_odom_set_iterations(self._h, *it[:3])
            else:
                self.lib.ref_odom_set_iterations4(self._h, *it[:4])

    def close(self):
        if getattr(self, "_h", None):
            self.lib.ref_odom_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def initICP(self, a, depthCutoff, normals=None):
        if normals is None:
            self.lib.ref_odom_init_icp_depth(self._h, _addr(a), depthCutoff)
        else:
            self.lib.ref_odom_init_icp_maps(self._h, _addr(a), _addr(normals), depthCutoff)

    def initICPModel(self, v, n, depthCutoff, modelPose):
        pose = np.ascontiguousarray(modelPose, dtype=np.float32).reshape(-1)
        self.lib.ref_odom_init_icp_model(self._h, _addr(v), _addr(n), depthCutoff, _fptr(pose))

    def initRGB(self, rgb): self.lib.ref_odom_init_rgb(self._h, _addr(rgb))
    def initRGBModel(self, rgb): self.lib.ref_odom_init_rgb_model(self._h, _addr(rgb))
    def initFirstRGB(self, rgb): self.lib.ref_odom_init_first_rgb(self._h, _addr(rgb))

    def getIncrementalTransformation(self, trans, rot, rgbOnly, icpWeight, pyramid, fastOdom, so3):
        t = np.ascontiguousarray(trans, dtype=np.float32).reshape(-1).copy()
        r = np.ascontiguousarray(rot, dtype=np.float32).reshape(-1).copy()
        self.lib.ref_odom_get_incremental_transformation(self._h, _fptr(t), _fptr(r), int(bool(rgbOnly)), float(icpWeight), int(bool(pyramid)),
                                                         int(bool(fastOdom)), int(bool(so3)))
        return t.reshape(np.shape(trans)), r.reshape(np.shape(rot))

    def stats(self) -> Stats:
        st = Stats()
        self.lib.ref_odom_get_stats(self._h, C.byref(st))
        return st

    def set_trace(self, on=True): self.lib.ref_odom_set_trace(self._h, int(on))

    def get_trace(self):
        arr = (StepRecord * 64)()
        n = self.lib.ref_odom_get_trace(self._h, arr, 64)
        return [arr[k].as_dict() for k in range(min(n, 64))]

    def tap(self, tap, level):
        from slam_b200.odometry import Tap
        h, w = self.height >> level, self.width >> level
        per = {0: 2, 1: 12, 2: 12, 3: 12, 4: 12, 5: 4, 6: 4, 7: 1, 8: 1, 9: 1, 10: 2, 11: 2, 12: 12, 13: 16}[tap]
        buf = np.empty(h * w * per, dtype=np.uint8)
        rc = self.lib.ref_odom_tap(self._h, tap, level, buf.ctypes.data)
        assert rc == 0
        return shape_tap(buf, tap, h, w)
