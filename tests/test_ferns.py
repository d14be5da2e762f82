"""Fern relocaliser: CPU restatement self-checks (no GPU) and CUDA-vs-restatement parity (B200).

The integer work -- resize, codes, co-occurrence counts, dissimilarities, arg-min, blockHDAware, addFrame decisions --
is compared bit for bit.  photometricCheck goes through an fp32 4x4 inverse (Eigen's in the reference, unpinned), so its
integer count may differ by a boundary pixel or two and its value is compared at 2 %."""
import numpy as np
import pytest

from oracle.ferns_oracle import BAD, FernsOracle, resize
from tests.support import ICL

W, H = ICL["width"], ICL["height"]
MAXD = 3000


def random_table(rng, n=500):
    return np.stack([rng.integers(0, W // 8, n), rng.integers(0, H // 8, n), rng.integers(0, 256, n), rng.integers(0, 256, n), rng.integers(0, 256, n),
                     rng.integers(400, MAXD + 1, n)], axis=1).astype(np.int32)


def random_frame(rng, holes=0.2):
    rgba = rng.integers(0, 256, (H, W, 4), dtype=np.uint8)
    vert = rng.normal(0, 1, (H, W, 4)).astype(np.float32)
    vert[..., 2] = rng.uniform(0.3, 4.0, (H, W)).astype(np.float32)
    vert[rng.random((H, W)) < holes, 2] = 0.0
    norm = rng.normal(0, 1, (H, W, 4)).astype(np.float32)
    return rgba, vert, norm


def make_oracle(table):
    return FernsOracle(table, MAXD, 115.0, ICL["cx"], ICL["cy"], ICL["fx"], ICL["fy"], W, H)


# ------------------------------------------------------------------ CPU
def test_resize_takes_the_centre_texel_of_every_8x8_block():
    a = np.arange(H * W, dtype=np.int32).reshape(H, W, 1)
    s = resize(a)
    assert s.shape == (H // 8, W // 8, 1)
    assert s[0, 0, 0] == 4 * W + 4 and s[3, 5, 0] == (8 * 3 + 4) * W + 8 * 5 + 4


def test_codes_follow_the_reference_bit_layout():
    table = np.array([[1, 2, 100, 100, 100, 1000], [3, 4, 0, 255, 10, 400], [5, 6, 1, 1, 1, 500]], np.int32)
    o = make_oracle(table)
    rgb = np.zeros((H // 8, W // 8, 3), np.uint8)
    vert = np.zeros((H // 8, W // 8, 4), np.float32)
    rgb[2, 1] = (101, 100, 200)
    vert[2, 1, 2] = 1.0005          # int(1000.5) = 1000, not > 1000
    rgb[4, 3] = (1, 255, 11)
    vert[4, 3, 2] = 0.4011          # 401 > 400
    codes, good = o.encode(rgb, vert)   # fern 2 sits on a hole
    assert list(codes) == [0b1010, 0b1011, BAD] and good == 2


def test_inverted_lists_count_equal_valid_codes():
    rng = np.random.default_rng(5)
    o = make_oracle(random_table(rng, 120))
    for t in range(6):
        rgba, vert, norm = random_frame(rng)
        o.addFrame(rgba, vert, norm, np.eye(4), t, -1.0)   # threshold -1: every frame is kept
    assert len(o.frames) == 6
    rgba, vert, norm = random_frame(rng)
    codes, good = o.encode(resize(rgba)[..., :3], resize(vert))
    co = o.co_occurrences(codes)
    brute = np.array([int(((f["codes"] == codes) & (codes != BAD)).sum()) for f in o.frames])
    assert np.array_equal(co, brute)


def test_add_frame_threshold_rule():
    rng = np.random.default_rng(6)
    o = make_oracle(random_table(rng, 200))
    rgba, vert, norm = random_frame(rng)
    assert o.addFrame(rgba, vert, norm, np.eye(4), 0, 0.5)          # the first frame is always kept
    assert not o.addFrame(rgba, vert, norm, np.eye(4), 1, 0.5)      # identical frame: dissimilarity 0
    rgba2, vert2, norm2 = random_frame(rng)
    assert o.addFrame(rgba2, vert2, norm2, np.eye(4), 2, 0.5)       # independent frame: ~15/16 of the codes differ
    empty = np.zeros_like(vert)
    assert not o.addFrame(rgba, empty, norm, np.eye(4), 3, 0.5)     # no good codes


def test_python_mirror_has_the_reference_method_names():
    from slam_b200.ferns import Ferns
    for name in ("addFrame", "findFrame", "conservatory", "lastClosest" if False else "numFrames"):
        assert hasattr(Ferns, name), name


# ------------------------------------------------------------------ GPU
@pytest.fixture(scope="module")
def torch_cuda(built):
    import torch
    assert torch.cuda.is_available()
    return torch


def dev(t, *arrays):
    return [t.from_numpy(a).to("cuda:0") for a in arrays]


@pytest.mark.gpu
def test_generated_conservatory_is_in_range_and_seeded(torch_cuda):
    from slam_b200.ferns import Ferns
    a = Ferns(500, MAXD, 115.0, ICL["cx"], ICL["cy"], ICL["fx"], ICL["fy"], W, H, seed=7)
    b = Ferns(500, MAXD, 115.0, ICL["cx"], ICL["cy"], ICL["fx"], ICL["fy"], W, H, seed=7)
    c = Ferns(500, MAXD, 115.0, ICL["cx"], ICL["cy"], ICL["fx"], ICL["fy"], W, H, seed=8)
    ta, tb, tc = a.conservatory, b.conservatory, c.conservatory
    assert np.array_equal(ta, tb) and not np.array_equal(ta, tc)
    assert ta[:, 0].min() >= 0 and ta[:, 0].max() < W // 8 and ta[:, 1].max() < H // 8
    assert ta[:, 2:5].min() >= 0 and ta[:, 2:5].max() <= 255 and ta[:, 5].min() >= 400 and ta[:, 5].max() <= MAXD
    for f in (a, b, c):
        f.close()


@pytest.mark.gpu
def test_encode_search_and_database_match_the_restatement(torch_cuda):
    from slam_b200.ferns import Ferns
    t = torch_cuda
    rng = np.random.default_rng(11)
    table = random_table(rng)
    o = make_oracle(table)
    g = Ferns(500, MAXD, 115.0, ICL["cx"], ICL["cy"], ICL["fx"], ICL["fy"], W, H, table=table, capacity=64)
    base = random_frame(rng)
    kept = 0
    for k in range(24):
        # a mixture of fresh frames and perturbed copies of the first, so that some are rejected by the threshold
        if k % 3 == 0:
            fr = random_frame(rng, holes=0.1 + 0.03 * k)
        else:
            fr = [a.copy() for a in base]
            m = rng.random((H, W)) < 0.02 * k
            fr[0][m] = rng.integers(0, 256, (int(m.sum()), 4), dtype=np.uint8)
        pose = np.eye(4, dtype=np.float32)
        pose[:3, 3] = rng.normal(0, 1, 3)
        d = dev(t, *fr)
        enc = g.encode(*d)
        rgb_s, vert_s = resize(fr[0])[..., :3], resize(fr[1])
        codes, good = o.encode(rgb_s, vert_s)
        assert np.array_equal(enc["codes"], codes) and enc["goodCodes"] == good
        assert np.array_equal(enc["rgb"], rgb_s) and np.array_equal(enc["vert"], vert_s) and np.array_equal(enc["norm"], resize(fr[2]))
        if o.frames:
            dis, _ = o.dissimilarities(codes, good)
            s = g.search()
            assert np.array_equal(s["dissim"].view(np.uint32), dis.view(np.uint32)), k
            assert s["minId"] == int(np.argmin(dis)) and s["minimum"] == dis.min()
        a_ref = o.addFrame(fr[0], fr[1], fr[2], pose, 10 * k, 0.25)
        a_gpu = g.addFrame(d[0], d[1], d[2], pose, 10 * k, 0.25)
        assert a_ref == a_gpu, k
        kept += int(a_ref)
    assert 2 < kept < 24 and g.numFrames() == kept == len(o.frames)
    for i, fr in enumerate(o.frames):
        got = g.frame(i)
        assert np.array_equal(got["codes"], fr["codes"]) and got["srcTime"] == fr["srcTime"] and got["goodCodes"] == fr["goodCodes"]
        assert np.array_equal(got["pose"], fr["pose"])
    # findFrame's search rule: only key frames older than 300 ticks are eligible
    q = random_frame(rng)
    dq = dev(t, *q)
    g.encode(*dq)
    for time in (100, 320, 1000):
        ref = o.search(q[0], q[1], q[2], time)
        s = g.search(time, use_time=True)
        assert s["minId"] == ref["minId"], time
        if ref["minId"] >= 0:
            assert s["minimum"] == ref["minimum"] and np.float32(s["blockHDAware"]) == ref["blockHDAware"]
    # an all-holes query has no good codes: nothing matches, nothing is added
    empty = np.zeros_like(q[1])
    de = dev(t, q[0], empty, q[2])
    assert g.encode(*de)["goodCodes"] == 0
    assert g.search()["minId"] == -1
    assert not g.addFrame(de[0], de[1], de[2], np.eye(4), 5000, 0.25)
    g.close()


@pytest.mark.gpu
def test_photometric_check_matches_the_restatement(torch_cuda, icl_sequence):
    from slam_b200.ferns import Ferns
    t = torch_cuda
    scene, intr, poses = icl_sequence
    rng = np.random.default_rng(3)
    table = random_table(rng)
    o = make_oracle(table)
    g = Ferns(500, MAXD, 115.0, ICL["cx"], ICL["cy"], ICL["fx"], ICL["fy"], W, H, table=table, capacity=8)
    mv, mn, mrgba = scene.render_model(poses[100])
    d = dev(t, mrgba, mv, mn)
    assert g.addFrame(d[0], d[1], d[2], poses[100], 0, 0.2) and o.addFrame(mrgba, mv, mn, poses[100], 0, 0.2)
    qv, qn, qrgba = scene.render_model(poses[104])
    dq = dev(t, qrgba, qv, qn)
    enc = g.encode(*dq)
    for est in (poses[104], poses[100], poses[110]):
        ref_err, ref_cnt = o.photometricCheck(enc["vert"], enc["rgb"], est, poses[100], o.frames[0]["rgb"])
        err, cnt = g.photometricCheck(0, est, poses[100])
        assert ref_cnt > 100 and abs(cnt - ref_cnt) <= 3, (cnt, ref_cnt)
        assert abs(err - ref_err) <= 0.02 * ref_err, (err, ref_err)
    # the right pose explains the colours better than a wrong one
    good_err, _ = g.photometricCheck(0, poses[104], poses[100])
    bad_err, _ = g.photometricCheck(0, poses[130], poses[100])
    assert good_err < bad_err
    g.close()


@pytest.mark.gpu
def test_find_frame_relocalises_on_the_synthetic_sequence(torch_cuda, icl_sequence):
    """Key frames every 25 poses of the orbit; a later frame near key frame 4 must be matched to it, refined by the
    1/8-resolution ICP to the true pose, accepted, and yield the reference's every-tenth-fern constraints."""
    from slam_b200.ferns import Ferns
    t = torch_cuda
    scene, intr, poses = icl_sequence
    g = Ferns(500, MAXD, 115.0, ICL["cx"], ICL["cy"], ICL["fx"], ICL["fy"], W, H, seed=1234, capacity=64)
    o = make_oracle(g.conservatory)
    for k in range(0, 400, 25):
        mv, mn, mrgba = scene.render_model(poses[k])
        d = dev(t, mrgba, mv, mn)
        assert g.addFrame(d[0], d[1], d[2], poses[k], k, 1e-3) == o.addFrame(mrgba, mv, mn, poses[k], k, 1e-3)
    assert g.numFrames() == len(o.frames) >= 10
    q = 102
    qv, qn, qrgba = scene.render_model(poses[q])
    dq = dev(t, qrgba, qv, qn)
    cons = []
    wrong = poses[q].copy()
    wrong[:3, 3] += 0.4       # the tracker's (lost) pose estimate: only used for the source points of the constraints
    est = g.findFrame(cons, wrong, dq[1], dq[2], dq[0], 1000, lost=True)
    m = g.lastMatch
    ref = o.search(qrgba, qv, qn, 1000)
    assert m.min_id == ref["minId"] == 4 and np.float32(m.dissimilarity) == ref["minimum"]
    assert np.float32(m.block_hd_aware) == ref["blockHDAware"] and m.block_hd_aware > 0.3
    assert m.icp_ran == 1 and m.icp_error < 3e-4 and m.icp_count > 1400
    assert np.abs(est[:3, 3] - poses[q][:3, 3]).max() < 5e-3 and np.abs(est[:3, :3] - poses[q][:3, :3]).max() < 5e-3
    assert g.lastClosest == 4 and m.photo_error < 115.0
    # constraints: every (num / 50)-th fern with a usable vertex; source = currPose * v, target = estPose * v
    tab = g.conservatory
    vs = resize(qv)
    expect = [vs[y, x] for (x, y, *_r) in tab[::10] if vs[y, x, 2] > 0 and int(np.float32(vs[y, x, 2]) * np.float32(1000.0)) < MAXD]
    assert len(cons) == len(expect) > 20
    for (src, dst), v in zip(cons, expect):
        p = np.array([v[0], v[1], v[2], 1.0], np.float32)
        assert np.allclose(src, wrong @ p, atol=1e-5) and np.allclose(dst, est @ p, atol=1e-5)
    # a query that is too recent for every key frame matches nothing and returns the identity
    cons2 = []
    est2 = g.findFrame(cons2, wrong, dq[1], dq[2], dq[0], 200, lost=True)
    assert g.lastClosest == -1 and not cons2 and np.array_equal(est2, np.eye(4, dtype=np.float32))
    g.close()
