"""Wire formats: .klg logs, TUM-style list pairs, the pose log.  CPU only.

The reference's readers cannot be built here (Pangolin, OpenCV 2); the formats are restated from inputs/RawLogReader.cpp,
inputs/FileReader.cpp and apps/elastic_fusion_file.cpp:383-387 and pinned by byte-level layouts written out by hand, round trips,
and independent codecs / conversions (zlib, Pillow, OpenCV, scipy)."""
import struct
import zlib

import numpy as np
import pytest

from slam_b200.io import KlgReader, KlgWriter, PoseLogWriter, TumListReader, quaternion_from_rotation, read_pose_log, rgb_to_rgba

W, H = 32, 24


def frames(n, seed=0):
    rng = np.random.default_rng(seed)
    out = []
    for k in range(n):
        depth = (rng.integers(400, 3000, (H, W))).astype(np.uint16)
        depth[rng.random((H, W)) < 0.1] = 0
        rgb = rng.integers(0, 256, (H, W, 3)).astype(np.uint8)
        out.append((1000 + 33 * k, depth, rgb))
    return out


def test_klg_layout_by_hand_and_reader_semantics(tmp_path):
    """A log assembled byte by byte (RawLogReader.cpp:31,74-109): raw and zlib depth, raw RGB and no image; the reader delivers
    numFrames - 1 frames (hasMore is `currentFrame + 1 < numFrames`) and flipColors swaps R and B."""
    fs = frames(3)
    p = tmp_path / "hand.klg"
    with open(p, "wb") as f:
        f.write(struct.pack("<i", 3))
        ts, d, c = fs[0]
        f.write(struct.pack("<qii", ts, W * H * 2, W * H * 3) + d.tobytes() + c.tobytes())
        ts, d, c = fs[1]
        z = zlib.compress(d.tobytes(), 9)
        f.write(struct.pack("<qii", ts, len(z), 0) + z)
        ts, d, c = fs[2]
        f.write(struct.pack("<qii", ts, W * H * 2, W * H * 3) + d.tobytes() + c.tobytes())
    with KlgReader(p, W, H) as r:
        assert r.num_frames == 3
        got = list(r)
    assert len(got) == 2
    assert got[0][0] == fs[0][0] and np.array_equal(got[0][1], fs[0][1]) and np.array_equal(got[0][2], fs[0][2])
    assert got[1][0] == fs[1][0] and np.array_equal(got[1][1], fs[1][1]) and not got[1][2].any()
    with KlgReader(p, W, H, flip_colors=True) as r:
        _, _, c = r.get_next()
    assert np.array_equal(c, fs[0][2][..., ::-1])


@pytest.mark.parametrize("depth_mode,image_mode", [("raw", "raw"), ("zlib", "raw"), ("zlib", "jpeg"), ("zlib", "none")])
def test_klg_round_trip(tmp_path, depth_mode, image_mode):
    fs = frames(5, seed=3)
    p = tmp_path / "log.klg"
    with KlgWriter(p, W, H, depth=depth_mode, image=image_mode, jpeg_quality=95) as w:
        for ts, d, c in fs:
            w.write(ts, d, c)
    assert struct.unpack("<i", open(p, "rb").read(4))[0] == 5
    with KlgReader(p, W, H) as r:
        got = [r.get_next() for _ in range(5)]
    for (ts, d, c), (ts2, d2, c2) in zip(fs, got):
        assert ts == ts2 and np.array_equal(d, d2)
        if image_mode == "raw":
            assert np.array_equal(c, c2)
        elif image_mode == "none":
            assert not c2.any()
        else:   # lossy codec: decoded by an independent decoder to the same pixels
            import cv2
            with open(p, "rb") as f:
                f.seek(4)
                _, dsz, isz = struct.unpack("<qii", f.read(16))
                f.seek(dsz, 1)
                jpg = f.read(isz)
            ref = cv2.cvtColor(cv2.imdecode(np.frombuffer(jpg, np.uint8), cv2.IMREAD_COLOR), cv2.COLOR_BGR2RGB)
            assert np.abs(ref.astype(int) - got[0][2].astype(int)).max() <= 2
            break
    with pytest.raises(EOFError):
        with KlgReader(p, W, H) as r:
            for _ in range(6):
                r.get_next()


def test_tum_list_reader(tmp_path):
    import cv2
    rng = np.random.default_rng(5)
    (tmp_path / "rgb").mkdir()
    (tmp_path / "depth").mkdir()
    raws, bgrs = [], []
    with open(tmp_path / "rgb.txt", "w") as fr, open(tmp_path / "depth.txt", "w") as fd:
        for k in range(4):
            raw = rng.integers(0, 65536, (H, W)).astype(np.uint16)
            raw[0, :5] = [0, 2, 3, 7, 65535]           # 0.4 -> 0, 0.6 -> 1, 1.4 -> 1, 13107 (saturation not reached)
            raw[1, :3] = [12, 13, 17]                  # 2.4 -> 2, 2.6 -> 3, 3.4 -> 3
            raw[2, :2] = [5, 15]                       # 1.0, 3.0 exactly
            bgr = rng.integers(0, 256, (H, W, 3)).astype(np.uint8)
            cv2.imwrite(str(tmp_path / "rgb" / f"{k}.png"), bgr)
            cv2.imwrite(str(tmp_path / "depth" / f"{k}.png"), raw)
            fr.write(f"{k * 0.033:.6f} rgb/{k}.png\n")
            fd.write(f"{k * 0.033:.6f} depth/{k}.png\n")
            raws.append(raw)
            bgrs.append(bgr)
    r = TumListReader(tmp_path / "rgb.txt", tmp_path / "depth.txt", str(tmp_path) + "/", W, H)
    assert r.num_frames == 4
    got = list(r)
    assert len(got) == 3                              # FileReader::hasMore drops the last record
    for k, (stamp, depth, bgr) in enumerate(got):
        assert np.array_equal(bgr, bgrs[k])
        src = raws[k]
        # the reference's own call: Mat::convertTo(CV_16UC1, 0.2)
        m = cv2.multiply(src.astype(np.float64), 0.2)
        expect = np.clip(np.rint(m), 0, 65535).astype(np.uint16)
        assert np.array_equal(depth, expect)
    assert list(got[0][1][0, :5]) == [0, 0, 1, 1, 13107] and list(got[0][1][1, :3]) == [2, 3, 3] and list(got[0][1][2, :2]) == [1, 3]


def test_pose_log_and_quaternion_convention(tmp_path):
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(9)
    rots = [Rotation.from_rotvec(v).as_matrix() for v in rng.normal(size=(40, 3))]
    rots += [Rotation.from_euler("xyz", e).as_matrix() for e in ([np.pi, 0, 0], [0, np.pi, 0], [0, 0, np.pi], [np.pi - 1e-3, 0.2, 0])]   # trace <= 0 branches
    for R in rots:
        q = quaternion_from_rotation(R)
        ref = Rotation.from_matrix(R).as_quat()   # x, y, z, w
        if np.dot(q, ref) < 0:
            ref = -ref
        assert np.abs(q - ref).max() < 2e-6
        assert abs(np.linalg.norm(q) - 1) < 1e-6
    p = tmp_path / "poses.txt"
    trans = rng.normal(size=(len(rots), 3)).astype(np.float32)
    with PoseLogWriter(p) as w:
        for k, (R, t) in enumerate(zip(rots, trans)):
            w.write(k + 1, t, R.astype(np.float32))
    lines = open(p).read().splitlines()
    assert len(lines) == len(rots) and all(len(ln.split()) == 8 for ln in lines)
    assert lines[0].split()[0] == "1"
    assert lines[0].split()[1] == "%g" % trans[0, 0]          # std::ostream default float format
    ticks, poses = read_pose_log(p)
    assert np.array_equal(ticks, np.arange(1, len(rots) + 1))
    for T, R, t in zip(poses, rots, trans):
        assert np.abs(T[:3, :3] - R).max() < 2e-5 and np.abs(T[:3, 3] - t).max() < 1e-5 * max(1, np.abs(t).max()) + 5e-6 * 10


def test_rgb_to_rgba():
    c = np.arange(H * W * 3, dtype=np.uint8).reshape(H, W, 3)
    out = rgb_to_rgba(c)
    assert out.shape == (H, W, 4) and np.array_equal(out[..., :3], c) and (out[..., 3] == 255).all()
