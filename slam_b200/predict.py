"""ctypes mirror of the reference's model-prediction producer over the C ABI (include/slam_predict.h).

Two GL classes of the reference feed the tracker its model maps:
``IndexMap::combinedPredict(pose, model, depthCutoff, confThreshold, time, maxTime, timeDelta, type)``
(src/model/IndexMap.cpp:243) and ``FillIn::vertex / normal / image(existing, raw, passthrough)`` (src/gl/FillIn.cpp:68-198),
driven by ``predict()`` in src/apps/elastic_fusion_file.cpp:17-44.  Here one handle owns both sets of textures;
``IndexMap`` and ``FillIn`` are thin views with the reference's method names and ``ModelPredictor.predict`` is the fused
two-launch call.  ``GPUTexture*`` arguments and the model VBO become device pointers (torch tensors or raw addresses).
The library does all the work; there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .odometry import OdometryError, _addr, _check, _fptr, load_library

SURFEL_FLOATS = 12


class PredictParams(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("cx", C.c_float), ("cy", C.c_float), ("fx", C.c_float), ("fy", C.c_float),
                ("max_point_size", C.c_float), ("device", C.c_int), ("stream", C.c_void_p)]


class Textures(C.Structure):
    _fields_ = [("image", C.c_void_p), ("vertex", C.c_void_p), ("normal", C.c_void_p), ("time", C.c_void_p), ("fill_image", C.c_void_p),
                ("fill_vertex", C.c_void_p), ("fill_normal", C.c_void_p), ("old_image", C.c_void_p), ("old_vertex", C.c_void_p), ("old_normal", C.c_void_p),
                ("old_time", C.c_void_p)]


_TEX = {"image": (0, np.uint8, 4), "vertex": (1, np.float32, 4), "normal": (2, np.float32, 4), "time": (3, np.uint16, 1),
        "fill_image": (4, np.uint8, 4), "fill_vertex": (5, np.float32, 4), "fill_normal": (6, np.float32, 4), "old_image": (7, np.uint8, 4),
        "old_vertex": (8, np.float32, 4), "old_normal": (9, np.float32, 4), "old_time": (10, np.uint16, 1)}

_bound = False


def _bind(lib):
    global _bound
    if _bound:
        return
    vp, fp, i, f = C.c_void_p, C.POINTER(C.c_float), C.c_int, C.c_float
    lib.slam_predict_create.argtypes = [C.POINTER(PredictParams), C.POINTER(vp)]
    lib.slam_predict_destroy.argtypes = [vp]
    lib.slam_predict_get_textures.argtypes = [vp, C.POINTER(Textures)]
    lib.slam_predict_combined.argtypes = [vp, vp, i, fp, f, f, i, i, i]
    lib.slam_predict_combined_type.argtypes = [vp, vp, i, fp, f, f, i, i, i, i]
    lib.slam_predict_fill_vertex.argtypes = [vp, vp, vp, i]
    lib.slam_predict_fill_normal.argtypes = [vp, vp, vp, i]
    lib.slam_predict_fill_image.argtypes = [vp, vp, vp, i]
    lib.slam_predict_frame.argtypes = [vp, vp, i, fp, f, f, i, i, i, vp, vp, i]
    lib.slam_predict_download.argtypes = [vp, i, vp]
    lib.slam_predict_get_tinv.argtypes = [vp, fp]
    lib.slam_predict_get_winners.argtypes = [vp, vp, vp]
    lib.slam_predict_last_ms.argtypes = [vp, fp, fp]
    lib.slam_predict_last_fragments.argtypes = [vp, C.POINTER(C.c_ulonglong)]
    _bound = True


def _pose(p) -> np.ndarray:
    return np.ascontiguousarray(p, dtype=np.float32).reshape(16).copy()


def _opt(x):
    return None if x is None else _addr(x)


class ModelPredictor:
    """IndexMap(width, height, intr) + FillIn(width, height, intr) on one handle."""

    def __init__(self, width, height, cx, cy, fx, fy, max_point_size=0.0, device=0, stream=None):
        self.lib = load_library()
        _bind(self.lib)
        p = PredictParams(width, height, cx, cy, fx, fy, max_point_size, device, stream)
        self._h = C.c_void_p()
        _check(self.lib, self.lib.slam_predict_create(C.byref(p), C.byref(self._h)))
        self.width, self.height = width, height
        self.textures = Textures()
        _check(self.lib, self.lib.slam_predict_get_textures(self._h, C.byref(self.textures)))
        self.indexMap = IndexMap(self)
        self.fillIn = FillIn(self)

    def close(self):
        if self._h:
            self.lib.slam_predict_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def predict(self, currPose, model, count, maxDepthProcessed, confidenceThreshold, tick, timeDelta, depthFiltered, rgb, maxTime=None,
                write_index_textures=False):
        """predict() of apps/elastic_fusion_file.cpp:17-44: combinedPredict(currPose, model, depth, conf, tick, tick, timeDelta, ACTIVE) followed
        by fillIn.vertex / normal / image(..., passthrough=false), as two launches."""
        P = _pose(currPose)
        _check(self.lib, self.lib.slam_predict_frame(self._h, _addr(model), int(count), _fptr(P), float(maxDepthProcessed), float(confidenceThreshold), int(tick),
                                                      int(tick if maxTime is None else maxTime), int(timeDelta), _addr(depthFiltered), _addr(rgb),
                                                      int(bool(write_index_textures))))

    def download(self, name: str) -> np.ndarray:
        idx, dtype, ch = _TEX[name]
        out = np.empty((self.height, self.width, ch) if ch > 1 else (self.height, self.width), dtype)
        _check(self.lib, self.lib.slam_predict_download(self._h, idx, out.ctypes.data))
        return out

    # ---- parity taps / timing aids
    def tInv(self) -> np.ndarray:
        t = np.zeros(16, np.float32)
        _check(self.lib, self.lib.slam_predict_get_tinv(self._h, _fptr(t)))
        return t.reshape(4, 4)

    def winners(self):
        d = np.empty((self.height, self.width), np.uint32)
        s = np.empty((self.height, self.width), np.int32)
        _check(self.lib, self.lib.slam_predict_get_winners(self._h, d.ctypes.data, s.ctypes.data))
        return d, s

    def lastMs(self):
        a, b = C.c_float(0), C.c_float(0)
        _check(self.lib, self.lib.slam_predict_last_ms(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def lastFragments(self) -> int:
        n = C.c_ulonglong(0)
        _check(self.lib, self.lib.slam_predict_last_fragments(self._h, C.byref(n)))
        return int(n.value)


class IndexMap:
    """The prediction half of src/model/IndexMap.h (combinedPredict + its four textures)."""

    ACTIVE, INACTIVE = 0, 1

    def __init__(self, owner: ModelPredictor):
        self._o = owner

    def combinedPredict(self, pose, model, count, depthCutoff, confThreshold, time, maxTime, timeDelta, predictionType=0):
        """model = device pointer of the surfel buffer (count x 12 floats), the reference's (vbo, count) pair."""
        o = self._o
        P = _pose(pose)
        _check(o.lib, o.lib.slam_predict_combined_type(o._h, _addr(model), int(count), _fptr(P), float(depthCutoff), float(confThreshold), int(time),
                                                        int(maxTime), int(timeDelta), int(predictionType)))

    def imageTex(self):
        return self._o.textures.image

    def vertexTex(self):
        return self._o.textures.vertex

    def normalTex(self):
        return self._o.textures.normal

    def timeTex(self):
        return self._o.textures.time

    def oldImageTex(self):
        return self._o.textures.old_image

    def oldVertexTex(self):
        return self._o.textures.old_vertex

    def oldNormalTex(self):
        return self._o.textures.old_normal

    def oldTimeTex(self):
        return self._o.textures.old_time


class FillIn:
    """src/gl/FillIn.h: vertex / normal / image(existing, raw, passthrough) and the three result textures."""

    def __init__(self, owner: ModelPredictor):
        self._o = owner

    def vertex(self, existingVertex, rawDepth, passthrough=False):
        o = self._o
        _check(o.lib, o.lib.slam_predict_fill_vertex(o._h, _opt(existingVertex), _addr(rawDepth), int(bool(passthrough))))

    def normal(self, existingNormal, rawDepth, passthrough=False):
        o = self._o
        _check(o.lib, o.lib.slam_predict_fill_normal(o._h, _opt(existingNormal), _addr(rawDepth), int(bool(passthrough))))

    def image(self, existingRgb, rawRgb, passthrough=False):
        o = self._o
        _check(o.lib, o.lib.slam_predict_fill_image(o._h, _opt(existingRgb), _addr(rawRgb), int(bool(passthrough))))

    @property
    def imageTexture(self):
        return self._o.textures.fill_image

    @property
    def vertexTexture(self):
        return self._o.textures.fill_vertex

    @property
    def normalTexture(self):
        return self._o.textures.fill_normal


__all__ = ["ModelPredictor", "IndexMap", "FillIn", "OdometryError", "SURFEL_FLOATS"]
