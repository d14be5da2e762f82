"""ctypes mirror of the reference's ``RGBDOdometryef`` interface over the C ABI (include/slam_odom.h).

Method names, argument meaning and call-order contract follow src/odom/RGBDOdometryef.h:28-70;
``GPUTexture*`` arguments become device pointers (``int`` addresses, e.g. ``tensor.data_ptr()``).
The library does all the work; this file only marshals arguments.  No CPU fallback: if
``libslam_odom.so`` is missing or no CUDA device is present the constructor raises.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from pathlib import Path

import numpy as np

_PKG = Path(__file__).resolve().parent
SLAM_MAX_LEVELS = 4


class OdometryError(RuntimeError):
    pass


class Tap:
    DEPTH_U16, VMAP_CURR, NMAP_CURR, VMAP_PREV, NMAP_PREV = 0, 1, 2, 3, 4
    LAST_DEPTH, NEXT_DEPTH, LAST_IMAGE, NEXT_IMAGE, LASTNEXT_IMAGE = 5, 6, 7, 8, 9
    DIDX, DIDY, CLOUD, CORRES = 10, 11, 12, 13

    DTYPE = {
        0: np.uint16, 1: np.float32, 2: np.float32, 3: np.float32, 4: np.float32, 5: np.float32, 6: np.float32,
        7: np.uint8, 8: np.uint8, 9: np.uint8, 10: np.int16, 11: np.int16, 12: np.float32, 13: np.uint8,
    }


class Params(C.Structure):
    _fields_ = [
        ("width", C.c_int), ("height", C.c_int),
        ("cx", C.c_float), ("cy", C.c_float), ("fx", C.c_float), ("fy", C.c_float),
        ("dist_thresh", C.c_float), ("angle_thresh", C.c_float),
        ("num_levels", C.c_int), ("iterations", C.c_int * SLAM_MAX_LEVELS),
        ("device", C.c_int), ("stream", C.c_void_p), ("batch", C.c_int), ("host_loop", C.c_int),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("lastICPError", C.c_float), ("lastICPCount", C.c_float),
        ("lastRGBError", C.c_float), ("lastRGBCount", C.c_float),
        ("lastSO3Error", C.c_float), ("lastSO3Count", C.c_float),
        ("lastA", C.c_double * 36), ("lastb", C.c_double * 6),
        ("so3_iterations", C.c_int), ("gn_iterations", C.c_int),
    ]


class StepRecord(C.Structure):
    _fields_ = [
        ("kind", C.c_int), ("level", C.c_int), ("iteration", C.c_int),
        ("so3", C.c_float * 11), ("icp", C.c_float * 29), ("rgb", C.c_float * 29),
        ("rgb_count", C.c_int), ("rgb_sigma", C.c_int),
        ("x", C.c_double * 6), ("Rcurr", C.c_float * 9), ("tcurr", C.c_float * 3),
        ("Rcurr_in", C.c_float * 9), ("tcurr_in", C.c_float * 3), ("krkinv_in", C.c_float * 9), ("kt_in", C.c_float * 3),
        ("sigma_in", C.c_float), ("so3_in", C.c_float * 27), ("t_cycles", C.c_uint * 8), ("t_solve", C.c_uint * 8),
    ]

    def as_dict(self):
        return dict(kind=self.kind, level=self.level, iteration=self.iteration, so3=np.array(self.so3[:], np.float32),
                    icp=np.array(self.icp[:], np.float32), rgb=np.array(self.rgb[:], np.float32), rgb_count=self.rgb_count,
                    rgb_sigma=self.rgb_sigma, x=np.array(self.x[:]), Rcurr=np.array(self.Rcurr[:], np.float32).reshape(3, 3),
                    tcurr=np.array(self.tcurr[:], np.float32), Rcurr_in=np.array(self.Rcurr_in[:], np.float32), tcurr_in=np.array(self.tcurr_in[:], np.float32),
                    krkinv_in=np.array(self.krkinv_in[:], np.float32), kt_in=np.array(self.kt_in[:], np.float32), sigma_in=float(self.sigma_in),
                    so3_in=np.array(self.so3_in[:], np.float32), t_cycles=np.array(self.t_cycles[:], np.int64), t_solve=np.array(self.t_solve[:], np.int64))


class FrameHost(C.Structure):
    _fields_ = [
        ("depth", C.c_void_p), ("rgba", C.c_void_p), ("model_vertices4", C.c_void_p), ("model_normals4", C.c_void_p),
        ("model_rgba", C.c_void_p), ("model_pose16", C.c_void_p), ("depth_cutoff", C.c_float), ("model_depth_cutoff", C.c_float),
    ]


_lib = None


def library_path() -> Path:
    return Path(os.environ.get("SLAM_ODOM_LIB", _PKG / "libslam_odom.so"))


def load_library():
    """Load libslam_odom.so (built in-tree by slam_b200/build.py). Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not path.exists():
        raise OdometryError(f"{path} not found: build it with `python -m slam_b200.build` (there is no CPU fallback)")
    lib = C.CDLL(str(path))
    vp, fp, i, f = C.c_void_p, C.POINTER(C.c_float), C.c_int, C.c_float
    lib.slam_odom_version.restype = C.c_char_p
    lib.slam_odom_last_error.restype = C.c_char_p
    lib.slam_odom_create.argtypes = [C.POINTER(Params), C.POINTER(vp)]
    lib.slam_odom_destroy.argtypes = [vp]
    lib.slam_odom_init_icp_depth.argtypes = [vp, vp, C.c_size_t, f]
    lib.slam_odom_init_icp_maps.argtypes = [vp, vp, vp, f]
    lib.slam_odom_init_icp_model.argtypes = [vp, vp, vp, f, fp]
    for n in ("slam_odom_init_rgb", "slam_odom_init_rgb_model", "slam_odom_init_first_rgb"):
        getattr(lib, n).argtypes = [vp, vp]
    lib.slam_odom_get_incremental_transformation.argtypes = [vp, fp, fp, i, f, i, i, i]
    lib.slam_odom_get_incremental_transformation_async.argtypes = [vp, fp, fp, i, f, i, i, i]
    lib.slam_odom_wait.argtypes = [vp, fp, fp]
    lib.slam_odom_get_covariance.argtypes = [vp, C.POINTER(C.c_double)]
    lib.slam_odom_get_stats.argtypes = [vp, C.POINTER(Stats)]
    lib.slam_odom_prefetch_host.argtypes = [vp, C.POINTER(FrameHost)]
    lib.slam_odom_track_host.argtypes = [vp, C.POINTER(FrameHost), fp, fp, i, f, i, i, i]
    lib.slam_odom_track_device.argtypes = [vp, C.POINTER(FrameHost), fp, fp, i, f, i, i, i]
    lib.slam_odom_track_host_next.argtypes = [vp, C.POINTER(FrameHost), C.POINTER(FrameHost), fp, fp, i, f, i, i, i]
    lib.slam_odom_track_sensor.argtypes = [vp, C.POINTER(FrameHost), C.POINTER(FrameHost), fp, fp, i, f, i, i, i]
    lib.slam_odom_tap_bytes.argtypes = [vp, i, i]
    lib.slam_odom_tap_bytes.restype = C.c_size_t
    lib.slam_odom_tap.argtypes = [vp, i, i, i, vp, C.c_size_t]
    lib.slam_odom_set_trace.argtypes = [vp, i]
    lib.slam_odom_get_trace.argtypes = [vp, i, C.POINTER(StepRecord), i, C.POINTER(i)]
    lib.slam_odom_launch_count.argtypes = [vp]
    lib.slam_odom_launch_count.restype = C.c_longlong
    lib.slam_odom_init_icp_depth_raw.argtypes = [vp, vp, f, f]
    lib.slam_op_depth_bilateral.argtypes = [vp, i, i, f, vp, i, vp]
    lib.slam_odom_score_poses.argtypes = [vp, i, i, i, fp, fp, fp, fp, fp, fp]
    lib.slam_odom_score_poses_best.argtypes = [vp, i, i, i, i, f, fp, fp, fp, fp, vp]
    lib.slam_odom_peer_export.argtypes = [vp, vp]
    lib.slam_odom_peer_connect.argtypes = [vp, i, i, vp]
    lib.slam_odom_score_poses_best_peers.argtypes = [vp, i, i, i, i, f, fp, fp, fp, fp, C.POINTER(C.c_ulonglong)]
    lib.slam_odom_set_profiling.argtypes = [vp, i]
    lib.slam_odom_set_split_launch.argtypes = [vp, i, C.POINTER(C.c_int)]
    lib.slam_odom_get_profile.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_longlong), i]
    lib.slam_odom_get_phase_cycles.argtypes = [vp, C.POINTER(C.c_ulonglong), i]
    lib.slam_odom_stream.argtypes = [vp]
    lib.slam_odom_stream.restype = vp
    lib.slam_op_workspace_bytes.restype = C.c_size_t
    # operator-level API
    lib.slam_op_pyr_down.argtypes = [vp, i, i, vp, vp]
    lib.slam_op_create_vmap.argtypes = [f, f, f, f, vp, i, i, vp, f, vp]
    lib.slam_op_create_nmap.argtypes = [vp, i, i, vp, vp]
    lib.slam_op_transform_maps.argtypes = [vp, vp, i, i, fp, fp, vp, vp, vp]
    lib.slam_op_copy_maps.argtypes = [vp, vp, i, i, vp, vp, vp]
    lib.slam_op_resize_vmap.argtypes = [vp, i, i, vp, vp]
    lib.slam_op_resize_nmap.argtypes = [vp, i, i, vp, vp]
    lib.slam_op_image_bgr_to_intensity.argtypes = [vp, i, i, vp, vp]
    lib.slam_op_vertices_to_depth.argtypes = [vp, i, i, vp, f, vp]
    lib.slam_op_project_to_point_cloud.argtypes = [vp, i, i, vp, f, f, f, f, i, vp]
    lib.slam_op_pyr_down_gauss_f.argtypes = [vp, i, i, vp, vp]
    lib.slam_op_pyr_down_uchar_gauss.argtypes = [vp, i, i, vp, vp]
    lib.slam_op_compute_derivative_images.argtypes = [vp, i, i, vp, vp, vp]
    lib.slam_op_icp_step.argtypes = [fp, fp, vp, vp, fp, fp, f, f, f, f, vp, vp, f, f, i, i, vp, vp, vp]
    lib.slam_op_compute_rgb_residual.argtypes = [f, vp, vp, vp, vp, vp, vp, vp, f, fp, fp, i, i, vp, vp, vp]
    lib.slam_op_rgb_step.argtypes = [vp, f, vp, f, f, vp, vp, f, i, i, vp, vp, vp]
    lib.slam_op_so3_step.argtypes = [vp, vp, fp, fp, fp, i, i, vp, vp, vp]
    _lib = lib
    return lib


_FP = C.POINTER(C.c_float)


def _fptr(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _check(lib, rc: int):
    if rc != 0:
        raise OdometryError(f"libslam_odom error {rc}: {lib.slam_odom_last_error().decode()}")


def _addr(x) -> int:
    """Device address of a torch tensor / raw int."""
    if hasattr(x, "data_ptr"):
        return int(x.data_ptr())
    return int(x)


class RGBDOdometry:
    """Same interface as the reference's ``RGBDOdometryef`` (src/odom/RGBDOdometryef.h:28-70)."""

    def __init__(self, width, height, cx, cy, fx, fy, distThresh=0.10, angleThresh=math.sin(20.0 * 3.14159254 / 180.0), *, num_levels=3,
                 iterations=None, device=0, stream=None, batch=1, host_loop=False):
        self.lib = load_library()
        p = Params()
        p.width, p.height = width, height
        p.cx, p.cy, p.fx, p.fy = cx, cy, fx, fy
        p.dist_thresh, p.angle_thresh = distThresh, angleThresh
        p.num_levels = num_levels
        for k, v in enumerate(iterations or []):
            p.iterations[k] = int(v)
        p.device = device
        p.stream = stream
        p.batch = batch
        p.host_loop = 1 if host_loop else 0
        self.width, self.height, self.num_levels, self.batch = width, height, num_levels, max(1, batch)
        self._h = C.c_void_p()
        _check(self.lib, self.lib.slam_odom_create(C.byref(p), C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None):
            self.lib.slam_odom_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- the reference's methods -------------------------------------------------
    def initICP(self, filteredDepth, depthCutoff, predictedNormals=None, *, pitch_bytes=0):
        """initICP(depth, cutoff) or initICP(vertices, normals, cutoff) (the two reference overloads)."""
        if predictedNormals is None:
            _check(self.lib, self.lib.slam_odom_init_icp_depth(self._h, _addr(filteredDepth), pitch_bytes, depthCutoff))
        else:
            _check(self.lib, self.lib.slam_odom_init_icp_maps(self._h, _addr(filteredDepth), _addr(predictedNormals), depthCutoff))

    def initICPModel(self, predictedVertices, predictedNormals, depthCutoff, modelPose):
        pose = np.ascontiguousarray(modelPose, dtype=np.float32).reshape(-1)
        assert pose.size == 16 * self.batch
        _check(self.lib, self.lib.slam_odom_init_icp_model(self._h, _addr(predictedVertices), _addr(predictedNormals), depthCutoff, _fptr(pose)))

    def initRGB(self, rgb):
        _check(self.lib, self.lib.slam_odom_init_rgb(self._h, _addr(rgb)))

    def initRGBModel(self, rgb):
        _check(self.lib, self.lib.slam_odom_init_rgb_model(self._h, _addr(rgb)))

    def initFirstRGB(self, rgb):
        _check(self.lib, self.lib.slam_odom_init_first_rgb(self._h, _addr(rgb)))

    def getIncrementalTransformation(self, trans, rot, rgbOnly, icpWeight, pyramid, fastOdom, so3):
        """trans (3,) / rot (3,3) float32, updated in place (batch: (B,3) / (B,3,3)); returns (trans, rot)."""
        t = np.ascontiguousarray(trans, dtype=np.float32).reshape(-1).copy()
        r = np.ascontiguousarray(rot, dtype=np.float32).reshape(-1).copy()
        assert t.size == 3 * self.batch and r.size == 9 * self.batch
        _check(self.lib, self.lib.slam_odom_get_incremental_transformation(self._h, _fptr(t), _fptr(r), int(bool(rgbOnly)), float(icpWeight),
                                                                            int(bool(pyramid)), int(bool(fastOdom)), int(bool(so3))))
        t = t.reshape(np.shape(trans))
        r = r.reshape(np.shape(rot))
        if isinstance(trans, np.ndarray) and trans.dtype == np.float32:
            trans[...] = t
        if isinstance(rot, np.ndarray) and rot.dtype == np.float32:
            rot[...] = r
        return t, r

    def getIncrementalTransformationAsync(self, trans, rot, rgbOnly, icpWeight, pyramid, fastOdom, so3):
        """Enqueue only (device-resident loop); collect with wait().  A pipelined caller may prepare the next frame in between."""
        t = np.ascontiguousarray(trans, dtype=np.float32).reshape(-1).copy()
        r = np.ascontiguousarray(rot, dtype=np.float32).reshape(-1).copy()
        _check(self.lib, self.lib.slam_odom_get_incremental_transformation_async(self._h, _fptr(t), _fptr(r), int(bool(rgbOnly)), float(icpWeight),
                                                                                  int(bool(pyramid)), int(bool(fastOdom)), int(bool(so3))))

    def wait(self):
        """-> (trans, rot) of the pending asynchronous track."""
        t = np.zeros(3 * self.batch, np.float32)
        r = np.zeros(9 * self.batch, np.float32)
        _check(self.lib, self.lib.slam_odom_wait(self._h, _fptr(t), _fptr(r)))
        return (t, r.reshape(3, 3)) if self.batch == 1 else (t.reshape(self.batch, 3), r.reshape(self.batch, 3, 3))

    def getCovariance(self):
        out = np.zeros(36 * self.batch, dtype=np.float64)
        _check(self.lib, self.lib.slam_odom_get_covariance(self._h, out.ctypes.data_as(C.POINTER(C.c_double))))
        return out.reshape((self.batch, 6, 6))[0] if self.batch == 1 else out.reshape((self.batch, 6, 6))

    # public fields of the reference class
    def stats(self, seq=0) -> Stats:
        arr = (Stats * self.batch)()
        _check(self.lib, self.lib.slam_odom_get_stats(self._h, arr))
        return arr[seq]

    @property
    def lastICPError(self): return self.stats().lastICPError
    @property
    def lastICPCount(self): return self.stats().lastICPCount
    @property
    def lastRGBError(self): return self.stats().lastRGBError
    @property
    def lastRGBCount(self): return self.stats().lastRGBCount
    @property
    def lastSO3Error(self): return self.stats().lastSO3Error
    @property
    def lastSO3Count(self): return self.stats().lastSO3Count
    @property
    def lastA(self): return np.array(self.stats().lastA[:]).reshape(6, 6)
    @property
    def lastb(self): return np.array(self.stats().lastb[:])

    # --- extras -------------------------------------------------------------------
    def make_frame(self, depth, rgba, model_vertices4, model_normals4, model_rgba, model_pose16, depth_cutoff, model_depth_cutoff) -> FrameHost:
        fr = FrameHost()
        fr.depth, fr.rgba = _addr(depth), _addr(rgba)
        fr.model_vertices4, fr.model_normals4, fr.model_rgba = _addr(model_vertices4), _addr(model_normals4), _addr(model_rgba)
        pose = np.ascontiguousarray(model_pose16, dtype=np.float32).reshape(-1)
        fr._pose_keepalive = pose
        fr.model_pose16 = pose.ctypes.data
        fr.depth_cutoff, fr.model_depth_cutoff = depth_cutoff, model_depth_cutoff
        return fr

    def _io_buffers(self):
        # per-handle in-out buffers with their ctypes pointers made once: the per-frame host cost of the mirror stays a few microseconds
        buf = getattr(self, "_io", None)
        if buf is None:
            t = np.zeros(3 * self.batch, np.float32)
            r = np.zeros(9 * self.batch, np.float32)
            buf = self._io = (t, r, t.ctypes.data_as(_FP), r.ctypes.data_as(_FP))
            self._r33 = r.reshape(3, 3) if self.batch == 1 else None
        return buf

    def _pose_in(self, t, r, trans, rot):
        # the common single-sequence call: (3,) and (3, 3) float32 arrays go straight into the call buffers
        if self._r33 is not None and type(trans) is np.ndarray and type(rot) is np.ndarray and trans.shape == (3,) and rot.shape == (3, 3):
            t[:] = trans
            self._r33[:] = rot
            return True
        t[:] = np.asarray(trans, dtype=np.float32).reshape(-1)
        r[:] = np.asarray(rot, dtype=np.float32).reshape(-1)
        return False

    def _track(self, fn, frame, trans, rot, rgbOnly, icpWeight, pyramid, fastOdom, so3, next_frame=None):
        t, r, tp, rp = self._io_buffers()
        fast = self._pose_in(t, r, trans, rot)
        flags = (1 if rgbOnly else 0, float(icpWeight), 1 if pyramid else 0, 1 if fastOdom else 0, 1 if so3 else 0)
        if next_frame is None:
            rc = fn(self._h, C.byref(frame), tp, rp, *flags)
        else:
            rc = fn(self._h, C.byref(frame), C.byref(next_frame), tp, rp, *flags)
        if rc != 0:
            _check(self.lib, rc)
        if fast:
            return t.copy(), self._r33.copy()
        return t.reshape(np.shape(trans)).copy(), r.reshape(np.shape(rot)).copy()

    def track_device(self, frame, trans, rot, rgbOnly=False, icpWeight=10.0, pyramid=True, fastOdom=False, so3=True):
        return self._track(self.lib.slam_odom_track_device, frame, trans, rot, rgbOnly, icpWeight, pyramid, fastOdom, so3)

    def track_host(self, frame, trans, rot, rgbOnly=False, icpWeight=10.0, pyramid=True, fastOdom=False, so3=True, next_frame=None):
        if next_frame is None:
            return self._track(self.lib.slam_odom_track_host, frame, trans, rot, rgbOnly, icpWeight, pyramid, fastOdom, so3)
        return self._track(self.lib.slam_odom_track_host_next, frame, trans, rot, rgbOnly, icpWeight, pyramid, fastOdom, so3, next_frame=next_frame)

    def track_sensor(self, frame, trans, rot, rgbOnly=False, icpWeight=10.0, pyramid=True, fastOdom=False, so3=True, next_frame=None):
        """The reference's data flow: frame.depth / frame.rgba in (pinned) host memory, the model prediction in device memory."""
        t, r, tp, rp = self._io_buffers()
        fast = self._pose_in(t, r, trans, rot)
        rc = self.lib.slam_odom_track_sensor(self._h, C.byref(frame), C.byref(next_frame) if next_frame is not None else None, tp, rp, 1 if rgbOnly else 0,
                                             float(icpWeight), 1 if pyramid else 0, 1 if fastOdom else 0, 1 if so3 else 0)
        if rc != 0:
            _check(self.lib, rc)
        if fast:
            return t.copy(), self._r33.copy()
        return t.reshape(np.shape(trans)).copy(), r.reshape(np.shape(rot)).copy()

    def prefetch_host(self, frame):
        _check(self.lib, self.lib.slam_odom_prefetch_host(self._h, C.byref(frame)))

    def set_trace(self, on=True):
        """True / 2: step records + full correspondence images (tests); 1: step records only (timelines); 0: off."""
        _check(self.lib, self.lib.slam_odom_set_trace(self._h, 2 if on is True else int(on)))

    def get_trace(self, seq=0):
        arr = (StepRecord * 64)()
        n = C.c_int(0)
        _check(self.lib, self.lib.slam_odom_get_trace(self._h, seq, arr, 64, C.byref(n)))
        return [arr[k].as_dict() for k in range(min(n.value, 64))]

    def initICPRaw(self, rawDepth, filterMaxDepth, depthCutoff):
        """The reference app's depth pre-filter (13x13 bilateral, apps/elastic_fusion_file.cpp:342-346) + initICP, on the device."""
        _check(self.lib, self.lib.slam_odom_init_icp_depth_raw(self._h, _addr(rawDepth), float(filterMaxDepth), float(depthCutoff)))

    def score_poses(self, level, prev_pose, trans_n, rot_n, seq=0):
        """Score n candidate poses of the current frame against the prepared model prediction (slam_odom_score_poses):
        -> (residual[n] = sum of squared point-to-plane distances, count[n] = inliers)."""
        prev = np.ascontiguousarray(prev_pose, dtype=np.float32)
        pt = np.ascontiguousarray(prev[:3, 3]).copy()
        pr = np.ascontiguousarray(prev[:3, :3]).reshape(-1).copy()
        t = np.ascontiguousarray(trans_n, dtype=np.float32).reshape(-1, 3)
        r = np.ascontiguousarray(rot_n, dtype=np.float32).reshape(-1, 9)
        assert len(t) == len(r)
        res = np.zeros(len(t), np.float32)
        cnt = np.zeros(len(t), np.float32)
        _check(self.lib, self.lib.slam_odom_score_poses(self._h, int(seq), int(level), len(t), _fptr(pt), _fptr(pr), _fptr(t), _fptr(r), _fptr(res), _fptr(cnt)))
        return res, cnt

    def score_poses_best(self, level, prev_pose, trans_n, rot_n, key_tensor, index_base=0, min_inliers=1.0, seq=0):
        """Enqueue the scoring of n hypotheses (global indices index_base ..) and fold the packed key of the best one into
        key_tensor (one int64 / uint64 element in device memory, preset to its maximum) with a 64-bit atomicMin.  No read-back."""
        prev = np.ascontiguousarray(prev_pose, dtype=np.float32)
        pt = np.ascontiguousarray(prev[:3, 3]).copy()
        pr = np.ascontiguousarray(prev[:3, :3]).reshape(-1).copy()
        t = np.ascontiguousarray(trans_n, dtype=np.float32).reshape(-1, 3)
        r = np.ascontiguousarray(rot_n, dtype=np.float32).reshape(-1, 9)
        assert len(t) == len(r) and len(t) > 0
        _check(self.lib, self.lib.slam_odom_score_poses_best(self._h, seq, int(level), len(t), int(index_base), float(min_inliers), _fptr(pt), _fptr(pr),
                                                             _fptr(t), _fptr(r), _addr(key_tensor)))

    def peer_export(self) -> bytes:
        """The CUDA IPC handle (64 bytes) of this handle's slot array for the cross-GPU minimum (see slam_odom.h)."""
        buf = C.create_string_buffer(64)
        _check(self.lib, self.lib.slam_odom_peer_export(self._h, buf))
        return buf.raw

    def peer_connect(self, rank: int, world: int, handles: bytes):
        assert len(handles) == 64 * world
        buf = C.create_string_buffer(handles, len(handles))
        _check(self.lib, self.lib.slam_odom_peer_connect(self._h, int(rank), int(world), buf))

    def score_poses_best_peers(self, level, prev_pose, trans_n, rot_n, index_base=0, min_inliers=1.0, seq=0) -> int:
        """Score this rank's block of hypotheses and return the winning packed key of ALL connected ranks (collective call)."""
        prev = np.ascontiguousarray(prev_pose, dtype=np.float32)
        pt = np.ascontiguousarray(prev[:3, 3]).copy()
        pr = np.ascontiguousarray(prev[:3, :3]).reshape(-1).copy()
        t = np.ascontiguousarray(trans_n, dtype=np.float32).reshape(-1, 3)
        r = np.ascontiguousarray(rot_n, dtype=np.float32).reshape(-1, 9)
        assert len(t) == len(r)
        key = C.c_ulonglong(0)
        _check(self.lib, self.lib.slam_odom_score_poses_best_peers(self._h, seq, int(level), len(t), int(index_base), float(min_inliers), _fptr(pt), _fptr(pr),
                                                                   _fptr(t) if len(t) else None, _fptr(r) if len(r) else None, C.byref(key)))
        return int(key.value)

    def set_split_launch(self, on=True) -> bool:
        """Cluster (SO3 pre-alignment) + fine-level kernel pair instead of one cooperative launch per frame; returns the previous setting."""
        prev = C.c_int(0)
        _check(self.lib, self.lib.slam_odom_set_split_launch(self._h, int(on), C.byref(prev)))
        return bool(prev.value)

    def set_profiling(self, on=True):
        _check(self.lib, self.lib.slam_odom_set_profiling(self._h, int(on)))

    def get_profile(self, reset=False):
        """-> (device milliseconds spent in the persistent Gauss-Newton kernel, number of its launches)."""
        ms, n = C.c_double(0), C.c_longlong(0)
        _check(self.lib, self.lib.slam_odom_get_profile(self._h, C.byref(ms), C.byref(n), int(reset)))
        return ms.value, n.value

    PHASES = ("staging", "so3 rest", "step set-up", "rgb assoc", "icp map", "icp reduce + count wait", "rgb products+reduce", "sums wait", "solve", "end barrier", "tail",
              "so3 map", "so3 reduce+post", "so3 wait", "so3 update", "launches", "stage L0", "stage L1", "stage L2", "stage L3", "stage so3 images", "hand-off wait", "ns: fine kernel start", "ns: hand-off seen")

    def get_phase_cycles(self, reset=False):
        """-> ({phase: SM cycles of the persistent kernel's leading CTA}, launches) accumulated since the last reset."""
        out = (C.c_ulonglong * 24)()
        _check(self.lib, self.lib.slam_odom_get_phase_cycles(self._h, out, int(reset)))
        return {name: int(out[k]) for k, name in enumerate(self.PHASES) if k != 15}, int(out[15])

    @property
    def stream(self) -> int:
        return int(self.lib.slam_odom_stream(self._h) or 0)

    def launch_count(self) -> int:
        return int(self.lib.slam_odom_launch_count(self._h))

    def tap(self, tap: int, level: int, seq: int = 0) -> np.ndarray:
        nbytes = self.lib.slam_odom_tap_bytes(self._h, tap, level)
        if nbytes == 0:
            raise OdometryError("bad tap/level")
        buf = np.empty(nbytes, dtype=np.uint8)
        _check(self.lib, self.lib.slam_odom_tap(self._h, tap, level, seq, buf.ctypes.data, nbytes))
        return shape_tap(buf, tap, self.height >> level, self.width >> level)


def shape_tap(buf: np.ndarray, tap: int, h: int, w: int) -> np.ndarray:
    a = buf.view(Tap.DTYPE[tap])
    if tap in (Tap.VMAP_CURR, Tap.NMAP_CURR, Tap.VMAP_PREV, Tap.NMAP_PREV):
        return a.reshape(3, h, w)
    if tap == Tap.CLOUD:
        return a.reshape(h, w, 3)
    if tap == Tap.CORRES:
        return a.reshape(h, w, 16)
    return a.reshape(h, w)


def corres_fields(c: np.ndarray):
    """Split a (h, w, 16) uint8 DataTerm image into zero(x,y), one(x,y), diff, valid (cuda/types.cuh:71-77)."""
    c = np.ascontiguousarray(c)
    s = c.view(np.int16).reshape(c.shape[0], c.shape[1], 8)
    diff = c.view(np.float32).reshape(c.shape[0], c.shape[1], 4)[..., 2]
    valid = c[..., 12] != 0
    return s[..., 0], s[..., 1], s[..., 2], s[..., 3], diff, valid
