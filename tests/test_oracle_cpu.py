"""Pin the CPU oracle (oracle/odom_oracle.c) against golden vectors produced by the REFERENCE's own CUDA kernels
(tests/golden/ref_track_128x96.npz, made on a B200 by tests/golden/make_golden.py from oracle/_ref).

The reference is compiled with approximate division / rsqrt and FMA contraction (src/CMakeLists.txt:115-116), which
a CPU cannot reproduce bit for bit: integer images may differ where a quotient falls within an ulp of a truncation
boundary, float maps agree to a few ulp, masks agree except on boundary pixels.  The tolerances below encode that.
"""
from pathlib import Path

import numpy as np
import pytest

from slam_b200.odometry import corres_fields

GOLD = Path(__file__).resolve().parent / "golden" / "ref_track_128x96.npz"


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


@pytest.fixture(scope="module")
def tracked(gold):
    from oracle.cpu_oracle import CpuOdometry
    W, H = int(gold["width"]), int(gold["height"])
    o = CpuOdometry(W, H, float(gold["cx"]), float(gold["cy"]), float(gold["fx"]), float(gold["fy"]))
    o.initFirstRGB(gold["first_rgba"])
    o.initICPModel(gold["mv"], gold["mn"], 20.0, gold["model_pose"])
    o.initRGBModel(gold["mrgba"])
    o.initICP(gold["depth"], 3.0)
    o.initRGB(gold["rgba"])
    pre = {(tap, l): o.tap(tap, l) for tap in range(10) for l in range(3)}
    pose = gold["model_pose"]
    t, r = o.getIncrementalTransformation(pose[:3, 3].copy(), pose[:3, :3].copy(), False, 10.0, True, False, True)
    post = {(tap, l): o.tap(tap, l) for tap in (10, 11, 12, 13) for l in range(3)}
    return o, pre, post, t, r


NAMES = {0: "depth_u16", 1: "vmap_curr", 2: "nmap_curr", 3: "vmap_prev", 4: "nmap_prev", 5: "last_depth", 6: "next_depth", 7: "last_image", 8: "next_image",
         9: "lastnext_image", 10: "didx", 11: "didy", 12: "cloud", 13: "corres"}


def test_integer_images_match_reference(gold, tracked):
    _, pre, post, _, _ = tracked
    for tap in (0, 7, 8, 9):
        for l in range(3):
            a, b = pre[(tap, l)], gold[f"{NAMES[tap]}_{l}"]
            frac = (a != b).mean()
            assert frac <= 2e-3, f"{NAMES[tap]} level {l}: {frac:.2e} of the pixels differ"
            assert np.abs(a.astype(np.int64) - b.astype(np.int64)).max() <= 1, f"{NAMES[tap]} level {l}: off by more than one count"
    for tap in (10, 11):
        for l in range(3):
            a, b = post[(tap, l)], gold[f"{NAMES[tap]}_{l}"]
            assert (a != b).mean() <= 2e-3 and np.abs(a.astype(np.int64) - b.astype(np.int64)).max() <= 1, f"{NAMES[tap]} level {l}"


def test_float_maps_match_reference(gold, tracked):
    _, pre, post, _, _ = tracked
    for tap in (1, 2, 3, 4):
        for l in range(3):
            a, b = pre[(tap, l)], gold[f"{NAMES[tap]}_{l}"]
            assert np.array_equal(np.isnan(a[0]), np.isnan(b[0])), f"{NAMES[tap]} level {l}: validity mask differs"
            ok = ~np.isnan(b[0])
            tol = 2e-6 if tap in (1, 3) else 2e-5   # normals go through rsqrt.approx on the GPU
            for c in range(3):
                assert np.allclose(a[c][ok], b[c][ok], rtol=tol, atol=tol), f"{NAMES[tap]} level {l} plane {c}: max {np.abs(a[c][ok] - b[c][ok]).max()}"
    for tap in (5, 6):
        for l in range(3):
            a, b = pre[(tap, l)], gold[f"{NAMES[tap]}_{l}"]
            assert np.array_equal(np.isnan(a), np.isnan(b))
            ok = ~np.isnan(b)
            assert np.allclose(a[ok], b[ok], rtol=2e-6, atol=1e-6)
    for l in range(3):
        a, b = post[(12, l)], gold[f"cloud_{l}"]
        ok = ~np.isnan(b)
        assert np.array_equal(np.isnan(a), np.isnan(b)) and np.allclose(a[ok], b[ok], rtol=2e-6, atol=1e-6)


def test_gauss_newton_steps_match_reference(gold, tracked):
    """First SO3 step and first ICP+RGB step see (nearly) identical inputs: counts within 0.5 %, sums within 1e-3;
    the final pose of the 22-step solve within 1e-4 of the reference's."""
    from oracle.cpu_oracle import Intr, load
    import ctypes as C
    lib = load()
    o, pre, post, t, r = tracked
    W, H = int(gold["width"]), int(gold["height"])
    kinds, levels = gold["step_kind"], gold["step_level"]
    # --- replay every reference step through the CPU operators with the reference's own step inputs
    for s in range(int(gold["n_steps"])):
        l = int(levels[s])
        h, w = H >> l, W >> l
        if kinds[s] == 0:
            out = np.zeros(16, np.float32)
            si = gold["step_so3_in"][s]
            keep = [np.ascontiguousarray(a) for a in (gold[f"lastnext_image_{l}"], gold[f"next_image_{l}"], si[0:9], si[9:18], si[18:27])]   # keep alive
            lib.oracle_so3_step(*(a.ctypes.data for a in keep), h, w, out.ctypes.data)
            ref = gold["step_so3"][s]
            assert abs(out[10] - ref[10]) <= max(3, 5e-3 * ref[10]), f"so3 step {s}: count {out[10]} vs {ref[10]}"
            assert abs(out[9] - ref[9]) <= 2e-2 * ref[9], f"so3 step {s}: residual {out[9]} vs {ref[9]}"
            continue
        div = np.float32(1 << l)
        k = Intr(*(float(np.float32(gold[n]) / div) for n in ("fx", "fy", "cx", "cy")))
        out = np.zeros(32, np.float32)
        pose = gold["model_pose"]
        args = [np.ascontiguousarray(a) for a in (gold["step_Rcurr_in"][s], gold["step_tcurr_in"][s], gold[f"vmap_curr_{l}"], gold[f"nmap_curr_{l}"],
                                                  gold["step_so3_in"][s][:9], pose[:3, 3].astype(np.float32), gold[f"vmap_prev_{l}"], gold[f"nmap_prev_{l}"])]
        lib.oracle_icp_step(args[0].ctypes.data, args[1].ctypes.data, args[2].ctypes.data, args[3].ctypes.data, args[4].ctypes.data, args[5].ctypes.data, k,
                            args[6].ctypes.data, args[7].ctypes.data, 0.10, float(np.float32(np.sin(20.0 * 3.14159254 / 180.0))), h, w, out.ctypes.data, None)
        ref = gold["step_icp"][s]
        assert abs(out[28] - ref[28]) <= max(3, 5e-3 * ref[28]), f"icp step {s}: inliers {out[28]} vs {ref[28]}"
        for idx in (0, 7, 13, 18, 22, 25):   # diagonal of JtJ
            assert abs(out[idx] - ref[idx]) <= 1e-2 * abs(ref[idx]) + 1e-6, f"icp step {s}: JtJ[{idx}] {out[idx]} vs {ref[idx]}"
    # --- whole solve
    assert np.abs(t - gold["trans"]).max() < 1e-4 and np.abs(r - gold["rot"]).max() < 1e-4, f"pose differs: {np.abs(t - gold['trans']).max()}"
    st = o.stats()
    assert st["gn_iterations"] == 19
    assert abs(st["lastICPCount"] - gold["stats"][1]) <= 5e-3 * gold["stats"][1]
    assert abs(st["lastRGBCount"] - gold["stats"][3]) <= 2e-2 * gold["stats"][3] + 5
    # the solve moved towards the ground truth
    gt = gold["gt_pose"][:3, 3]
    assert np.linalg.norm(t - gt) < 0.5 * np.linalg.norm(gold["model_pose"][:3, 3] - gt)


def test_rgb_correspondence_mask_close_to_reference(gold, tracked):
    _, pre, post, _, _ = tracked
    for l in range(3):
        _, _, _, _, _, vm = corres_fields(post[(13, l)])
        _, _, _, _, _, vr = corres_fields(gold[f"corres_{l}"])
        assert (vm != vr).sum() <= max(4, 2e-2 * vr.sum()), f"level {l}: {(vm != vr).sum()} of {vr.sum()} correspondences differ"
