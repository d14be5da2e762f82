"""Depth pre-filter: the 13x13 bilateral filter that produces the input of initICP.

PARITY UNPINNED against the reference itself: there it is a GLSL fragment shader (gl/shaders/depth_bilateral.frag) and no GL
context, golden image or test exists.  What is pinned here:
  * the C restatement (oracle/depth_filter_oracle.c, fp32, statement by statement) against an independent fp64 evaluation of
    the shader's formula: equal, except +-1 mm where sum1/sum2 is within fp32 noise of a .5 boundary  (CPU)
  * the CUDA kernel against the C restatement on full-size frames with noise, holes, out-of-range values and at ragged /
    tiny sizes: bit-exact on >= 99.9 % of the pixels, +-1 mm on the rest (fp32 tolerance: CUDA expf is 2 ulp, libm < 1 ulp;
    the validity mask -- zeros -- must be identical)                                                              (GPU)
"""
import numpy as np
import pytest


def synthetic_depth(rows, cols, seed=0):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:rows, 0:cols]
    d = 1400.0 + 600.0 * np.sin(xx / 37.0) * np.cos(yy / 29.0) + 0.8 * xx
    d[rows // 3: rows // 3 + rows // 5, cols // 4: cols // 4 + cols // 6] -= 500.0          # a step edge (the filter must not blur across it)
    d += rng.normal(0.0, 4.0, d.shape)
    d = np.clip(d, 0, 65535).astype(np.uint16)
    holes = rng.random(d.shape) < 0.02
    d[holes] = 0
    d[rng.random(d.shape) < 0.005] = 250          # below the 300 mm floor
    d[rng.random(d.shape) < 0.005] = 4500         # beyond maxD
    return d


def bilateral_fp64(d, max_d):
    """The shader's formula in float64 (pure numpy, small images only)."""
    rows, cols = d.shape
    out = np.zeros_like(d)
    cut = int(np.float32(max_d) * np.float32(1000.0))
    src = d.astype(np.float64)
    for y in range(rows):
        for x in range(cols):
            v = int(d[y, x])
            if v > cut or v < 300:
                continue
            y0, y1 = max(y - 6, 0), min(y + 7, rows)
            x0, x1 = max(x - 6, 0), min(x + 7, cols)
            win = src[y0:y1, x0:x1]
            gy, gx = np.mgrid[y0:y1, x0:x1]
            w = np.exp(-(((x - gx) ** 2 + (y - gy) ** 2) * 0.024691358 + (v - win) ** 2 * 0.000555556))
            out[y, x] = int(np.floor((win * w).sum() / w.sum() + 0.5))
    return out


def test_oracle_matches_fp64_formula(built):
    from oracle.cpu_oracle import depth_bilateral
    for shape, seed in (((40, 52), 1), ((13, 7), 2), ((5, 31), 3)):
        d = synthetic_depth(*shape, seed=seed)
        got = depth_bilateral(d, 3.0)
        want = bilateral_fp64(d, 3.0)
        assert np.array_equal(got == 0, want == 0)
        diff = np.abs(got.astype(np.int32) - want.astype(np.int32))
        assert diff.max() <= 1 and (diff != 0).mean() < 0.01, f"{shape}: max {diff.max()}, differing {(diff != 0).mean():.4f}"
    # values outside [300 mm, maxD] are dropped, the rest stays within the noise of the input
    d = synthetic_depth(40, 52, seed=1)
    got = depth_bilateral(d, 3.0)
    assert np.all(got[(d < 300) | (d > 3000)] == 0) and np.all(got[(d >= 300) & (d <= 3000)] > 0)
    # the step edge survives: pixels next to the 500 mm step keep their side
    inside = got[40 // 3 + 2, 52 // 4 + 2]
    outside = got[40 // 3 - 2, 52 // 4 + 2]
    assert abs(int(outside) - int(inside)) > 350


@pytest.mark.gpu
@pytest.mark.parametrize("shape,n", [((480, 640), 1), ((720, 1280), 1), ((94, 126), 3), ((9, 5), 2), ((1, 1), 1)])
def test_cuda_depth_filter_matches_oracle(built, shape, n):
    import torch
    from oracle.cpu_oracle import depth_bilateral
    from slam_b200.odometry import load_library
    lib = load_library()
    imgs = np.stack([synthetic_depth(*shape, seed=10 + k) for k in range(n)])
    src = torch.from_numpy(imgs.view(np.int16)).to("cuda:0")
    dst = torch.zeros_like(src)
    assert lib.slam_op_depth_bilateral(src.data_ptr(), shape[0], shape[1], 3.0, dst.data_ptr(), n, None) == 0
    torch.cuda.synchronize()
    got = dst.cpu().numpy().view(np.uint16)
    for k in range(n):
        want = depth_bilateral(imgs[k], 3.0)
        assert np.array_equal(got[k] == 0, want == 0), f"image {k}: validity mask differs"
        diff = np.abs(got[k].astype(np.int32) - want.astype(np.int32))
        assert diff.max() <= 1, f"image {k}: max |diff| {diff.max()} mm"
        assert (diff != 0).mean() <= 1e-3, f"image {k}: {(diff != 0).sum()} of {diff.size} pixels differ"


@pytest.mark.gpu
def test_init_icp_depth_raw_equals_filter_then_init(built, icl_sequence):
    import torch
    from slam_b200 import RGBDOdometry, Tap
    from slam_b200.odometry import load_library
    from tests.support import frame_pair, to_device
    scene, intr, poses = icl_sequence
    fr = frame_pair(scene, poses, 300)
    d = to_device(fr)
    lib = load_library()
    filt = torch.zeros_like(d["depth"])
    assert lib.slam_op_depth_bilateral(d["depth"].data_ptr(), 480, 640, 3.0, filt.data_ptr(), 1, None) == 0
    torch.cuda.synchronize()
    args = (intr["width"], intr["height"], intr["cx"], intr["cy"], intr["fx"], intr["fy"])
    a, b = RGBDOdometry(*args), RGBDOdometry(*args)
    a.set_trace(1)
    b.set_trace(1)
    a.initICPRaw(d["depth"], 3.0, 3.0)
    b.initICP(filt, 3.0)
    for level in range(3):
        for tap in (Tap.VMAP_CURR, Tap.NMAP_CURR):
            x, y = a.tap(tap, level), b.tap(tap, level)
            assert np.array_equal(x.view(np.uint32), y.view(np.uint32))
    a.close()
    b.close()
