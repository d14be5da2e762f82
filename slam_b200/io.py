"""Wire formats on either side of the tracking path: the reference's input readers and its pose log.

  * KlgReader / KlgWriter   the `.klg` RGB-D log of `inputs/RawLogReader.cpp:21-122`: int32 frame count, then per frame
                            int64 timestamp, int32 depthSize, int32 imageSize, depth bytes (raw u16 or zlib), image
                            bytes (raw RGB8, JPEG, or none)
  * TumListReader           the `rgb.txt` / `depth.txt` list pair of `inputs/FileReader.cpp:3-67` (TUM / ICL-NUIM png
                            export): one "<stamp> <file>" record per line, 16-bit depth in 1/5000 m converted to millimetres
  * PoseLogWriter           `tick tx ty tz qx qy qz qw`, one line per tracked frame (`apps/elastic_fusion_file.cpp:383-387`),
                            the format `benchmark/evaluate_ate.py` / `evaluate_rpe.py` read
  * rgb_to_rgba             RGB8 -> the RGBA8 texel layout the tracker's image inputs use (GPUTexture::RGB)

Host-side Python (the reference's readers are host code as well); PARITY UNPINNED against the reference's binaries (they need
Pangolin / OpenCV 2 to build): the formats are restated from the cited sources and pinned by round trips and by independent
implementations (zlib, Pillow / OpenCV codecs, scipy's rotation conversions) in tests/test_io_formats.py.
"""
from __future__ import annotations

import io as _io
import struct
import zlib
from pathlib import Path

import numpy as np


def rgb_to_rgba(rgb: np.ndarray) -> np.ndarray:
    """[H, W, 3] u8 -> [H, W, 4] u8 with alpha 255 (a GL_RGB upload into an RGBA8 texture)."""
    rgb = np.asarray(rgb, np.uint8)
    out = np.empty(rgb.shape[:2] + (4,), np.uint8)
    out[..., :3] = rgb
    out[..., 3] = 255
    return out


# ------------------------------------------------------------------------------------------------ .klg
class KlgWriter:
    """Writes the format RawLogReader reads.  depth: 'raw' or 'zlib'; image: 'raw', 'jpeg' or 'none'."""

    def __init__(self, path, width: int, height: int, depth: str = "zlib", image: str = "raw", jpeg_quality: int = 90):
        assert depth in ("raw", "zlib") and image in ("raw", "jpeg", "none")
        self.path, self.width, self.height = Path(path), width, height
        self.depth_mode, self.image_mode, self.jpeg_quality = depth, image, jpeg_quality
        self.fp = open(self.path, "wb")
        self.fp.write(struct.pack("<i", 0))   # frame count, patched on close (RawLogReader.cpp:31)
        self.frames = 0

    def write(self, timestamp: int, depth_u16: np.ndarray, rgb_u8: np.ndarray | None):
        d = np.ascontiguousarray(depth_u16, dtype="<u2")
        assert d.shape == (self.height, self.width)
        dbytes = d.tobytes() if self.depth_mode == "raw" else zlib.compress(d.tobytes())
        if self.depth_mode == "zlib" and len(dbytes) == self.width * self.height * 2:
            dbytes = d.tobytes()   # the reader tells raw from compressed by the size alone (RawLogReader.cpp:88)
        if self.image_mode == "none" or rgb_u8 is None:
            ibytes = b""
        else:
            c = np.ascontiguousarray(rgb_u8, dtype=np.uint8)
            assert c.shape == (self.height, self.width, 3)
            if self.image_mode == "raw":
                ibytes = c.tobytes()
            else:
                from PIL import Image
                buf = _io.BytesIO()
                Image.fromarray(c, "RGB").save(buf, format="JPEG", quality=self.jpeg_quality)
                ibytes = buf.getvalue()
        self.fp.write(struct.pack("<qii", int(timestamp), len(dbytes), len(ibytes)))   # :74-79
        self.fp.write(dbytes)
        self.fp.write(ibytes)
        self.frames += 1

    def close(self):
        if self.fp:
            self.fp.seek(0)
            self.fp.write(struct.pack("<i", self.frames))
            self.fp.close()
            self.fp = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


class KlgReader:
    """inputs/RawLogReader.cpp: sequential reader; iterating yields (timestamp, depth u16 [H, W] mm, rgb u8 [H, W, 3])."""

    def __init__(self, path, width: int, height: int, flip_colors: bool = False):
        self.path, self.width, self.height, self.flip_colors = Path(path), width, height, flip_colors
        self.num_pixels = width * height
        self.fp = open(self.path, "rb")
        (self.num_frames,) = struct.unpack("<i", self._read(4))   # :31
        self.current_frame = 0

    def _read(self, n: int) -> bytes:
        b = self.fp.read(n)
        if len(b) != n:
            raise EOFError(f"{self.path}: truncated log (wanted {n} bytes, got {len(b)})")
        return b

    def has_more(self) -> bool:
        return self.current_frame + 1 < self.num_frames   # RawLogReader::hasMore (the last frame is never delivered)

    def get_next(self):
        timestamp, depth_size, image_size = struct.unpack("<qii", self._read(16))   # :74-79
        dbytes = self._read(depth_size)
        ibytes = self._read(image_size) if image_size > 0 else b""
        if depth_size != self.num_pixels * 2:   # :88-96
            dbytes = zlib.decompress(dbytes)
            if len(dbytes) != self.num_pixels * 2:
                raise ValueError(f"{self.path}: frame {self.current_frame}: depth inflates to {len(dbytes)} bytes")
        depth = np.frombuffer(dbytes, dtype="<u2").reshape(self.height, self.width).copy()
        if image_size == self.num_pixels * 3:   # :98-109
            rgb = np.frombuffer(ibytes, np.uint8).reshape(self.height, self.width, 3).copy()
        elif image_size > 0:
            from PIL import Image
            rgb = np.asarray(Image.open(_io.BytesIO(ibytes)).convert("RGB"), np.uint8).copy()
            if rgb.shape != (self.height, self.width, 3):
                raise ValueError(f"{self.path}: frame {self.current_frame}: JPEG is {rgb.shape}")
        else:
            rgb = np.zeros((self.height, self.width, 3), np.uint8)
        if self.flip_colors:   # :114-120
            rgb = rgb[..., ::-1].copy()
        self.current_frame += 1
        return timestamp, depth, rgb

    def __iter__(self):
        while self.has_more():
            yield self.get_next()

    def close(self):
        if self.fp:
            self.fp.close()
            self.fp = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


# ------------------------------------------------------------------------------------------------ TUM-style lists
class TumListReader:
    """inputs/FileReader.cpp: `rgb.txt` and `depth.txt` with one "<frame> <relative file>" record per line, read in lock step
    (no timestamp association, as in the reference); depth png in 1/5000 m -> millimetres with OpenCV's saturating
    round-half-to-even conversion (`convertTo(CV_16UC1, 1000 / 5000)`, :52)."""

    def __init__(self, rgb_info, depth_info, dataset_dir, width: int, height: int):
        self.dataset_dir = str(dataset_dir)
        self.width, self.height = width, height
        with open(rgb_info) as f:
            text = f.read()
        self.num_frames = text.count("\n")   # :11-15: counts newlines
        self.rgb_records = [ln.split()[:2] for ln in text.splitlines() if len(ln.split()) >= 2]
        with open(depth_info) as f:
            self.depth_records = [ln.split()[:2] for ln in f.read().splitlines() if len(ln.split()) >= 2]
        self.current_frame = 0

    def has_more(self) -> bool:
        return self.current_frame + 1 < self.num_frames   # :64-67

    @staticmethod
    def depth_to_mm(raw16: np.ndarray) -> np.ndarray:
        scaled = raw16.astype(np.float64) * (1000.0 * 1.0 / 5000.0)
        return np.clip(np.rint(scaled), 0, 65535).astype(np.uint16)   # cv::saturate_cast<ushort>(cvRound(x))

    def get_next(self):
        import cv2
        frame, file_rgb = self.rgb_records[self.current_frame]
        _, file_depth = self.depth_records[self.current_frame]
        bgr = cv2.imread(self.dataset_dir + file_rgb, cv2.IMREAD_UNCHANGED)     # :50 (channel order as stored by OpenCV: BGR)
        raw = cv2.imread(self.dataset_dir + file_depth, cv2.IMREAD_UNCHANGED)   # :51
        if bgr is None or raw is None:
            raise FileNotFoundError(f"{self.dataset_dir}{file_rgb} / {file_depth}")
        self.current_frame += 1
        return frame, self.depth_to_mm(raw), bgr

    def __iter__(self):
        while self.has_more():
            yield self.get_next()


# ------------------------------------------------------------------------------------------------ pose log
def quaternion_from_rotation(R) -> np.ndarray:
    """Eigen::Quaternionf(Matrix3f) (Eigen/src/Geometry/Quaternion.h, quaternionbase_assign_impl): returns (x, y, z, w)."""
    m = np.asarray(R, np.float32)
    t = np.float32(m[0, 0] + m[1, 1] + m[2, 2])
    q = np.zeros(4, np.float32)   # x, y, z, w
    if t > 0:
        t = np.sqrt(t + np.float32(1.0))
        q[3] = np.float32(0.5) * t
        t = np.float32(0.5) / t
        q[0] = (m[2, 1] - m[1, 2]) * t
        q[1] = (m[0, 2] - m[2, 0]) * t
        q[2] = (m[1, 0] - m[0, 1]) * t
    else:
        i = 0
        if m[1, 1] > m[0, 0]:
            i = 1
        if m[2, 2] > m[i, i]:
            i = 2
        j = (i + 1) % 3
        k = (j + 1) % 3
        t = np.sqrt(m[i, i] - m[j, j] - m[k, k] + np.float32(1.0))
        q[i] = np.float32(0.5) * t
        t = np.float32(0.5) / t
        q[3] = (m[k, j] - m[j, k]) * t
        q[j] = (m[j, i] + m[i, j]) * t
        q[k] = (m[k, i] + m[i, k]) * t
    return q


def _ostream_float(v) -> str:
    """std::ostream << float with the default format: %g with 6 significant digits."""
    return "%g" % float(np.float32(v))


class PoseLogWriter:
    """apps/elastic_fusion_file.cpp:383-387: `tick tx ty tz qx qy qz qw` per tracked frame."""

    def __init__(self, path):
        self.fp = open(path, "w")

    def write(self, tick: int, trans, rot):
        q = quaternion_from_rotation(rot)
        t = np.asarray(trans, np.float32).reshape(3)
        self.fp.write(" ".join([str(int(tick))] + [_ostream_float(x) for x in (t[0], t[1], t[2], q[0], q[1], q[2], q[3])]) + "\n")

    def close(self):
        if self.fp:
            self.fp.close()
            self.fp = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def read_pose_log(path):
    """-> (ticks [n], poses [n, 4, 4] float64); the inverse of PoseLogWriter, as benchmark/evaluate_ate.py parses it."""
    ticks, poses = [], []
    with open(path) as f:
        for ln in f:
            v = ln.split()
            if len(v) != 8 or ln.startswith("#"):
                continue
            x, y, z, w = (float(a) for a in v[4:8])
            n = np.sqrt(x * x + y * y + z * z + w * w)
            x, y, z, w = x / n, y / n, z / n, w / n
            T = np.eye(4)
            T[:3, :3] = [[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                         [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                         [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]]
            T[:3, 3] = [float(a) for a in v[1:4]]
            ticks.append(float(v[0]))
            poses.append(T)
    return np.array(ticks), np.array(poses)
