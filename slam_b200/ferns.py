"""ctypes mirror of the reference's ``Ferns`` class (src/lc/Ferns.h:28-181) over the C ABI (include/slam_ferns.h).

Method names and argument meaning follow the reference: ``addFrame(image, vertex, normal, pose, srcTime, threshold)``
(Ferns.cpp:83), ``findFrame(constraints, currPose, vertex, normal, image, time, lost)`` (Ferns.cpp:170).  ``GPUTexture*``
arguments become device pointers (``tensor.data_ptr()`` / torch tensors).  The library does all the work; no CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .odometry import OdometryError, _check, _fptr, load_library


class Fern(C.Structure):
    _fields_ = [("x", C.c_int32), ("y", C.c_int32), ("r", C.c_int32), ("g", C.c_int32), ("b", C.c_int32), ("d", C.c_int32)]


class FernsParams(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("cx", C.c_float), ("cy", C.c_float), ("fx", C.c_float), ("fy", C.c_float),
                ("num_ferns", C.c_int), ("max_depth_mm", C.c_int), ("photo_thresh", C.c_float), ("capacity", C.c_int), ("seed", C.c_uint32),
                ("device", C.c_int)]


class SurfaceConstraint(C.Structure):
    _fields_ = [("source", C.c_float * 4), ("target", C.c_float * 4)]


class Match(C.Structure):
    _fields_ = [("min_id", C.c_int), ("dissimilarity", C.c_float), ("block_hd_aware", C.c_float), ("icp_ran", C.c_int), ("icp_error", C.c_float),
                ("icp_count", C.c_float), ("photo_error", C.c_float), ("last_closest", C.c_int)]


_bound = False


def _bind(lib):
    global _bound
    if _bound:
        return
    vp, fp, i, f = C.c_void_p, C.POINTER(C.c_float), C.c_int, C.c_float
    ip, u8p = C.POINTER(C.c_int), C.POINTER(C.c_uint8)
    lib.slam_ferns_create.argtypes = [C.POINTER(FernsParams), C.POINTER(Fern), C.POINTER(vp)]
    lib.slam_ferns_destroy.argtypes = [vp]
    lib.slam_ferns_get_table.argtypes = [vp, C.POINTER(Fern)]
    lib.slam_ferns_num_frames.argtypes = [vp]
    lib.slam_ferns_add_frame.argtypes = [vp, vp, vp, vp, fp, i, f, ip]
    lib.slam_ferns_find_frame.argtypes = [vp, fp, vp, vp, vp, i, i, fp, C.POINTER(Match), C.POINTER(SurfaceConstraint), i, ip]
    lib.slam_ferns_encode.argtypes = [vp, vp, vp, vp, u8p, ip, u8p, fp, fp]
    lib.slam_ferns_search.argtypes = [vp, i, i, fp, ip, fp, fp]
    lib.slam_ferns_photometric_check.argtypes = [vp, i, fp, fp, fp, ip]
    lib.slam_ferns_get_frame.argtypes = [vp, i, u8p, fp, ip, ip]
    lib.slam_ferns_last_search_ms.argtypes = [vp, fp]
    _bound = True


def _addr(x) -> int:
    return int(x.data_ptr()) if hasattr(x, "data_ptr") else int(x)


def _pose(p) -> np.ndarray:
    return np.ascontiguousarray(p, dtype=np.float32).reshape(16).copy()


class Ferns:
    """Ferns(n, maxDepth, photoThresh, intr, w, h) -- Ferns.cpp:21-57; ``table`` / ``seed`` replace the time(0) seed."""

    def __init__(self, n, maxDepth, photoThresh, cx, cy, fx, fy, w, h, seed=0, table=None, capacity=1024, device=0):
        self.lib = load_library()
        _bind(self.lib)
        p = FernsParams(w, h, cx, cy, fx, fy, n, int(maxDepth), photoThresh, capacity, seed, device)
        tab = None
        if table is not None:
            tab = (Fern * n)(*[Fern(*[int(v) for v in row]) for row in table])
        self._h = C.c_void_p()
        _check(self.lib, self.lib.slam_ferns_create(C.byref(p), tab, C.byref(self._h)))
        self.num, self.width, self.height = n, w // 8, h // 8
        self.lastClosest = -1
        self.lastMatch = None

    def close(self):
        if self._h:
            self.lib.slam_ferns_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def conservatory(self) -> np.ndarray:
        """[num][6] int32: x, y, r, g, b, d"""
        t = (Fern * self.num)()
        _check(self.lib, self.lib.slam_ferns_get_table(self._h, t))
        return np.array([[e.x, e.y, e.r, e.g, e.b, e.d] for e in t], dtype=np.int32)

    def numFrames(self) -> int:
        return int(self.lib.slam_ferns_num_frames(self._h))

    def addFrame(self, imageTexture, vertexTexture, normalTexture, pose, srcTime, threshold) -> bool:
        added = C.c_int(0)
        P = _pose(pose)
        _check(self.lib, self.lib.slam_ferns_add_frame(self._h, _addr(imageTexture), _addr(vertexTexture), _addr(normalTexture), _fptr(P), int(srcTime),
                                                        float(threshold), C.byref(added)))
        return bool(added.value)

    def findFrame(self, constraints: list, currPose, vertexTexture, normalTexture, imageTexture, time, lost) -> np.ndarray:
        """Returns estPose (4x4); appends (source, target) pairs to ``constraints``; sets lastClosest / lastMatch."""
        P = _pose(currPose)
        est = np.zeros(16, np.float32)
        m = Match()
        cons = (SurfaceConstraint * self.num)()
        n = C.c_int(0)
        _check(self.lib, self.lib.slam_ferns_find_frame(self._h, _fptr(P), _addr(vertexTexture), _addr(normalTexture), _addr(imageTexture), int(time),
                                                         int(bool(lost)), _fptr(est), C.byref(m), cons, self.num, C.byref(n)))
        for k in range(n.value):
            constraints.append((np.array(cons[k].source[:], np.float32), np.array(cons[k].target[:], np.float32)))
        self.lastClosest = m.last_closest
        self.lastMatch = m
        return est.reshape(4, 4)

    # ---- operator-level taps
    def encode(self, imageTexture, vertexTexture, normalTexture):
        n = self.width * self.height
        codes = np.zeros(self.num, np.uint8)
        good = C.c_int(0)
        rgb, vert, norm = np.zeros((self.height, self.width, 3), np.uint8), np.zeros((self.height, self.width, 4), np.float32), np.zeros(
            (self.height, self.width, 4), np.float32)
        _check(self.lib, self.lib.slam_ferns_encode(self._h, _addr(imageTexture), _addr(vertexTexture), _addr(normalTexture),
                                                     codes.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(good), rgb.ctypes.data_as(C.POINTER(C.c_uint8)),
                                                     _fptr(vert), _fptr(norm)))
        return dict(codes=codes, goodCodes=good.value, rgb=rgb, vert=vert, norm=norm)

    def search(self, time=0, use_time=False):
        nf = self.numFrames()
        d = np.zeros(max(nf, 1), np.float32)
        mid, mn, hd = C.c_int(-1), C.c_float(0), C.c_float(0)
        _check(self.lib, self.lib.slam_ferns_search(self._h, int(time), int(bool(use_time)), _fptr(d), C.byref(mid), C.byref(mn), C.byref(hd)))
        return dict(dissim=d[:nf], minId=mid.value, minimum=mn.value, blockHDAware=hd.value)

    def photometricCheck(self, frame_id, estPose, fernPose):
        e, f = _pose(estPose), _pose(fernPose)
        err, cnt = C.c_float(0), C.c_int(0)
        _check(self.lib, self.lib.slam_ferns_photometric_check(self._h, int(frame_id), _fptr(e), _fptr(f), C.byref(err), C.byref(cnt)))
        return err.value, cnt.value

    def frame(self, frame_id):
        codes = np.zeros(self.num, np.uint8)
        pose = np.zeros(16, np.float32)
        t, g = C.c_int(0), C.c_int(0)
        _check(self.lib, self.lib.slam_ferns_get_frame(self._h, int(frame_id), codes.ctypes.data_as(C.POINTER(C.c_uint8)), _fptr(pose), C.byref(t), C.byref(g)))
        return dict(codes=codes, pose=pose.reshape(4, 4), srcTime=t.value, goodCodes=g.value)

    def lastSearchMs(self) -> float:
        ms = C.c_float(0)
        _check(self.lib, self.lib.slam_ferns_last_search_ms(self._h, C.byref(ms)))
        return ms.value


__all__ = ["Ferns", "OdometryError"]
