"""Operator-level parity on the GPU: each slam_op_* entry point of the C ABI against the
reference's own wrapper of the same name (src/odom/utils.cuh:62-175) compiled for sm_100a
(oracle/_ref/libslam_ref.so), on identical synthetic inputs.

Bar (BASELINE.json north_star): integer / index / mask outputs bit-exact, float maps bit-exact
(same per-pixel expression trees, same nvcc numeric flags), reduced sums within 1e-4 relative.
"""
import ctypes as C

import numpy as np
import pytest

from tests.support import (ANGLE_THRESH, DEPTH_CUTOFF, frame_pair, planar_map_mismatch, se3_sums_rel_err, so3_sums_rel_err)

pytestmark = pytest.mark.gpu


def fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


@pytest.fixture(scope="module")
def env(icl_sequence, ref_lib):
    import torch
    from slam_b200.odometry import load_library
    scene, intr, poses = icl_sequence
    lib = load_library()
    fr = frame_pair(scene, poses, 100)
    dev = "cuda:0"
    t = lambda a: torch.from_numpy(a).to(dev)
    d = dict(depth=t(fr["depth"].view(np.int16)), rgba=t(fr["rgba"]), mv=t(fr["mv"]), mn=t(fr["mn"]), mrgba=t(fr["mrgba"]))
    ws = torch.zeros(lib.slam_op_workspace_bytes(), dtype=torch.uint8, device=dev)
    return dict(torch=torch, lib=lib, ref=ref_lib, intr=intr, fr=fr, d=d, ws=ws, dev=dev, poses=poses)


def nan_f32(torch, shape, dev):
    return torch.full(shape, float("nan"), dtype=torch.float32, device=dev)


def depth_pyramid(env):
    torch, lib, dev = env["torch"], env["lib"], env["dev"]
    H, W = 480, 640
    levels = [env["d"]["depth"]]
    for l in range(2):
        dst = torch.zeros((H >> (l + 1), W >> (l + 1)), dtype=torch.int16, device=dev)
        assert lib.slam_op_pyr_down(levels[-1].data_ptr(), H >> l, W >> l, dst.data_ptr(), None) == 0
        levels.append(dst)
    torch.cuda.synchronize()
    return levels


def test_pyr_down_bit_exact(env):
    torch, ref, dev = env["torch"], env["ref"], env["dev"]
    mine = depth_pyramid(env)
    src = env["d"]["depth"]
    for l in range(2):
        r = torch.zeros_like(mine[l + 1])
        ref.ref_op_pyr_down(src.data_ptr(), 480 >> l, 640 >> l, r.data_ptr())
        torch.cuda.synchronize()
        assert torch.equal(mine[l + 1], r), f"level {l + 1}: {(mine[l + 1] != r).sum().item()} pixels differ"
        src = r


def vmaps_and_nmaps(env, which):
    torch, lib, ref, dev, intr = env["torch"], env["lib"], env["ref"], env["dev"], env["intr"]
    pyr = depth_pyramid(env)
    out = []
    for l in range(3):
        h, w = 480 >> l, 640 >> l
        div = float(1 << l)
        fx, fy, cx, cy = (float(np.float32(intr[k]) / np.float32(div)) for k in ("fx", "fy", "cx", "cy"))
        v = nan_f32(torch, (3, h, w), dev)
        n = nan_f32(torch, (3, h, w), dev)
        if which == "mine":
            assert lib.slam_op_create_vmap(fx, fy, cx, cy, pyr[l].data_ptr(), h, w, v.data_ptr(), DEPTH_CUTOFF, None) == 0
            assert lib.slam_op_create_nmap(v.data_ptr(), h, w, n.data_ptr(), None) == 0
        else:
            ref.ref_op_create_vmap(fx, fy, cx, cy, pyr[l].data_ptr(), h, w, v.data_ptr(), DEPTH_CUTOFF, 1)
            ref.ref_op_create_nmap(v.data_ptr(), h, w, n.data_ptr(), 1)
        torch.cuda.synchronize()
        out.append((v, n))
    return out


def test_vmap_nmap_bit_exact(env):
    mine = vmaps_and_nmaps(env, "mine")
    ref = vmaps_and_nmaps(env, "ref")
    for l in range(3):
        for name, a, b in (("vmap", mine[l][0], ref[l][0]), ("nmap", mine[l][1], ref[l][1])):
            nan_diff, val_diff = planar_map_mismatch(a.cpu().numpy(), b.cpu().numpy())
            assert nan_diff == 0 and val_diff == 0, f"{name} level {l}: nan pattern differs at {nan_diff}, values at {val_diff} pixels"
    valid = ~np.isnan(mine[0][0].cpu().numpy()[0])
    assert 0.3 < valid.mean() < 1.0   # the frame has both valid and invalid (beyond depthCutoff / zero depth) pixels


def model_maps(env, which):
    torch, lib, ref, dev = env["torch"], env["lib"], env["ref"], env["dev"]
    d = env["d"]
    pose = env["fr"]["model_pose"]
    R = np.ascontiguousarray(pose[:3, :3], dtype=np.float32).reshape(-1)
    t = np.ascontiguousarray(pose[:3, 3], dtype=np.float32)
    cam, glob = [], []
    v = nan_f32(torch, (3, 480, 640), dev)
    n = nan_f32(torch, (3, 480, 640), dev)
    if which == "mine":
        assert lib.slam_op_copy_maps(d["mv"].data_ptr(), d["mn"].data_ptr(), 480, 640, v.data_ptr(), n.data_ptr(), None) == 0
    else:
        ref.ref_op_copy_maps(d["mv"].data_ptr(), d["mn"].data_ptr(), 480, 640, v.data_ptr(), n.data_ptr())
    cam.append((v, n))
    for l in range(1, 3):
        h, w = 480 >> l, 640 >> l
        v2 = nan_f32(torch, (3, h, w), dev)
        n2 = nan_f32(torch, (3, h, w), dev)
        if which == "mine":
            assert lib.slam_op_resize_vmap(cam[-1][0].data_ptr(), h * 2, w * 2, v2.data_ptr(), None) == 0
            assert lib.slam_op_resize_nmap(cam[-1][1].data_ptr(), h * 2, w * 2, n2.data_ptr(), None) == 0
        else:
            ref.ref_op_resize_map(cam[-1][0].data_ptr(), h * 2, w * 2, v2.data_ptr(), 0)
            ref.ref_op_resize_map(cam[-1][1].data_ptr(), h * 2, w * 2, n2.data_ptr(), 1)
        cam.append((v2, n2))
    for l in range(3):
        h, w = 480 >> l, 640 >> l
        vg = cam[l][0].clone()
        ng = cam[l][1].clone()
        if which == "mine":
            assert lib.slam_op_transform_maps(cam[l][0].data_ptr(), cam[l][1].data_ptr(), h, w, fp(R), fp(t), vg.data_ptr(), ng.data_ptr(), None) == 0
        else:
            ref.ref_op_transform_maps(cam[l][0].data_ptr(), cam[l][1].data_ptr(), h, w, fp(R), fp(t), vg.data_ptr(), ng.data_ptr())
        glob.append((vg, ng))
    torch.cuda.synchronize()
    return cam, glob


def test_model_map_operators_bit_exact(env):
    cam_m, glob_m = model_maps(env, "mine")
    cam_r, glob_r = model_maps(env, "ref")
    for l in range(3):
        for name, a, b in (("copy/resize v", cam_m[l][0], cam_r[l][0]), ("copy/resize n", cam_m[l][1], cam_r[l][1]),
                           ("transform v", glob_m[l][0], glob_r[l][0]), ("transform n", glob_m[l][1], glob_r[l][1])):
            nan_diff, val_diff = planar_map_mismatch(a.cpu().numpy(), b.cpu().numpy())
            assert nan_diff == 0 and val_diff == 0, f"{name} level {l}: nan {nan_diff}, values {val_diff}"


def rgbd_pyramids(env, which, rgba_key="rgba"):
    torch, lib, ref, dev = env["torch"], env["lib"], env["ref"], env["dev"]
    d = env["d"]
    depth = [torch.zeros((480, 640), dtype=torch.float32, device=dev)]
    image = [torch.zeros((480, 640), dtype=torch.uint8, device=dev)]
    if which == "mine":
        assert lib.slam_op_vertices_to_depth(d["mv"].data_ptr(), 480, 640, depth[0].data_ptr(), 6.0, None) == 0
        assert lib.slam_op_image_bgr_to_intensity(d[rgba_key].data_ptr(), 480, 640, image[0].data_ptr(), None) == 0
    else:
        ref.ref_op_vertices_to_depth(d["mv"].data_ptr(), 480, 640, depth[0].data_ptr(), 6.0)
        ref.ref_op_image_bgr_to_intensity(d[rgba_key].data_ptr(), 480, 640, image[0].data_ptr())
    for l in range(2):
        h, w = 480 >> l, 640 >> l
        dd = torch.zeros((h // 2, w // 2), dtype=torch.float32, device=dev)
        ii = torch.zeros((h // 2, w // 2), dtype=torch.uint8, device=dev)
        if which == "mine":
            assert lib.slam_op_pyr_down_gauss_f(depth[-1].data_ptr(), h, w, dd.data_ptr(), None) == 0
            assert lib.slam_op_pyr_down_uchar_gauss(image[-1].data_ptr(), h, w, ii.data_ptr(), None) == 0
        else:
            ref.ref_op_pyr_down_gauss_f(depth[-1].data_ptr(), h, w, dd.data_ptr())
            ref.ref_op_pyr_down_uchar_gauss(image[-1].data_ptr(), h, w, ii.data_ptr())
        depth.append(dd)
        image.append(ii)
    grads = []
    for l in range(3):
        h, w = 480 >> l, 640 >> l
        dx = torch.zeros((h, w), dtype=torch.int16, device=dev)
        dy = torch.zeros((h, w), dtype=torch.int16, device=dev)
        if which == "mine":
            assert lib.slam_op_compute_derivative_images(image[l].data_ptr(), h, w, dx.data_ptr(), dy.data_ptr(), None) == 0
        else:
            ref.ref_op_compute_derivative_images(image[l].data_ptr(), h, w, dx.data_ptr(), dy.data_ptr())
        grads.append((dx, dy))
    torch.cuda.synchronize()
    return depth, image, grads


def test_rgbd_pyramids_and_derivatives_bit_exact(env):
    dm, im, gm = rgbd_pyramids(env, "mine")
    dr, ir, gr = rgbd_pyramids(env, "ref")
    for l in range(3):
        a, b = dm[l].cpu().numpy(), dr[l].cpu().numpy()
        assert np.array_equal(np.isnan(a), np.isnan(b)), f"float depth level {l}: NaN pattern differs"
        ok = ~np.isnan(a)
        assert np.array_equal(a.view(np.uint32)[ok], b.view(np.uint32)[ok]), f"float depth level {l}: {(a.view(np.uint32)[ok] != b.view(np.uint32)[ok]).sum()} differ"
        assert env["torch"].equal(im[l], ir[l]), f"intensity level {l}: {(im[l] != ir[l]).sum().item()} differ"
        assert env["torch"].equal(gm[l][0], gr[l][0]) and env["torch"].equal(gm[l][1], gr[l][1]), f"derivatives level {l} differ"


def test_point_cloud_bit_exact(env):
    torch, lib, ref, dev, intr = env["torch"], env["lib"], env["ref"], env["dev"], env["intr"]
    depth, _, _ = rgbd_pyramids(env, "ref")
    for l in range(3):
        h, w = 480 >> l, 640 >> l
        a = torch.zeros((h, w, 3), dtype=torch.float32, device=dev)
        b = torch.zeros((h, w, 3), dtype=torch.float32, device=dev)
        assert lib.slam_op_project_to_point_cloud(depth[l].data_ptr(), h, w, a.data_ptr(), intr["fx"], intr["fy"], intr["cx"], intr["cy"], l, None) == 0
        ref.ref_op_project_to_point_cloud(depth[l].data_ptr(), h, w, b.data_ptr(), intr["fx"], intr["fy"], intr["cx"], intr["cy"], l)
        torch.cuda.synchronize()
        an, bn = a.cpu().numpy(), b.cpu().numpy()
        assert np.array_equal(np.isnan(an), np.isnan(bn))
        ok = ~np.isnan(an)
        assert np.array_equal(an.view(np.uint32)[ok], bn.view(np.uint32)[ok]), f"level {l}"


def icp_inputs(env, perturb=True):
    """Current maps from the depth frame, model maps in the global frame, and a slightly wrong pose."""
    curr = vmaps_and_nmaps(env, "ref")
    _, glob = model_maps(env, "ref")
    pose = env["fr"]["model_pose"].astype(np.float32)
    Rprev = pose[:3, :3].copy()
    tprev = pose[:3, 3].copy()
    Rcurr, tcurr = Rprev.copy(), tprev.copy()
    if perturb:
        tcurr = tcurr + np.array([0.004, -0.003, 0.005], np.float32)
    Rprev_inv = np.linalg.inv(Rprev.astype(np.float64)).astype(np.float32)
    return curr, glob, Rcurr, tcurr, Rprev_inv, tprev


def run_icp(env, which, level, curr, glob, Rcurr, tcurr, Rprev_inv, tprev, vc_override=None):
    torch, lib, ref, dev, intr = env["torch"], env["lib"], env["ref"], env["dev"], env["intr"]
    h, w = 480 >> level, 640 >> level
    div = np.float32(1 << level)
    fx, fy, cx, cy = (float(np.float32(intr[k]) / div) for k in ("fx", "fy", "cx", "cy"))
    vc = curr[level][0] if vc_override is None else vc_override
    args = (fp(np.ascontiguousarray(Rcurr.reshape(-1))), fp(np.ascontiguousarray(tcurr)), vc.data_ptr(), curr[level][1].data_ptr(),
            fp(np.ascontiguousarray(Rprev_inv.reshape(-1))), fp(np.ascontiguousarray(tprev)), fx, fy, cx, cy, glob[level][0].data_ptr(),
            glob[level][1].data_ptr(), 0.10, float(np.float32(ANGLE_THRESH)), h, w)
    if which == "mine":
        out = torch.zeros(32, dtype=torch.float32, device=dev)
        assert lib.slam_op_icp_step(*args, env["ws"].data_ptr(), out.data_ptr(), None) == 0
        torch.cuda.synchronize()
        return out.cpu().numpy()[:29]
    host = np.zeros(32, dtype=np.float32)
    ref.ref_op_icp_step(*args, fp(host))
    return host[:29]


def test_icp_step_sums_and_inlier_count(env):
    inp = icp_inputs(env)
    for level in range(3):
        m = run_icp(env, "mine", level, *inp)
        r = run_icp(env, "ref", level, *inp)
        assert r[28] > 0.2 * (480 >> level) * (640 >> level), "degenerate test frame"
        assert m[28] == r[28], f"level {level}: inlier count {m[28]} vs {r[28]}"
        err = se3_sums_rel_err(m, r)
        assert err < 1e-4, f"level {level}: JtJ/Jtr relative error {err}"


def test_icp_mask_per_row_bit_exact(env):
    """The reference kernel has no per-pixel output; isolate rows by NaN-ing every other row of the
    current vertex map and compare the inlier count of each row band: any single-pixel disagreement
    of the correspondence mask changes a band's count."""
    torch = env["torch"]
    curr, glob, Rcurr, tcurr, Rprev_inv, tprev = icp_inputs(env)
    level = 0
    h, w = 480, 640
    base = curr[level][0]
    mismatched = 0
    for band in range(0, h, 8):
        vc = base.clone()
        vc[0, :band, :] = float("nan")
        vc[0, band + 8:, :] = float("nan")
        m = run_icp(env, "mine", level, curr, glob, Rcurr, tcurr, Rprev_inv, tprev, vc_override=vc)
        r = run_icp(env, "ref", level, curr, glob, Rcurr, tcurr, Rprev_inv, tprev, vc_override=vc)
        mismatched += int(m[28] != r[28])
    assert mismatched == 0, f"{mismatched} of {h // 8} row bands have different inlier counts"


def rgb_inputs(env):
    """last* from the model prediction, next* from a model rendered at the current pose (model-to-model style, so that
    nextDepth != lastDepth and the warp matters)."""
    torch, dev = env["torch"], env["dev"]
    scene_fr = env["fr"]
    dl, il, _ = rgbd_pyramids(env, "ref", "mrgba")
    # "next" side: vertices rendered at the gt pose of the current frame
    from tests.support import make_scene
    scene, _ = make_scene(640, 480)
    mv2, mn2, mrgba2 = scene.render_model(scene_fr["gt_pose"])
    saved = dict(env["d"])
    env["d"]["mv"] = torch.from_numpy(mv2).to(dev)
    env["d"]["rgba2"] = torch.from_numpy(mrgba2).to(dev)
    dn, inn, gn = rgbd_pyramids(env, "ref", "rgba2")
    env["d"].update(saved)
    # relative motion: Rt = inverse of (T_model^-1 * T_curr), as RGBDOdometryef.cpp:422-432 with resultRt = I => Rt = I
    return dl, il, dn, inn, gn


def test_rgb_residual_mask_bit_exact_and_rgb_step(env):
    torch, lib, ref, dev, intr = env["torch"], env["lib"], env["ref"], env["dev"], env["intr"]
    dl, il, dn, inn, gn = rgb_inputs(env)
    from slam_b200.odometry import corres_fields
    # a small non-identity warp
    ang = 0.004
    Rrel = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]])
    trel = np.array([0.006, -0.002, 0.004])
    for level in range(3):
        h, w = 480 >> level, 640 >> level
        div = np.float32(1 << level)
        fx, fy, cx, cy = (float(np.float32(intr[k]) / div) for k in ("fx", "fy", "cx", "cy"))
        K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], dtype=np.float64)
        krk = (K @ Rrel @ np.linalg.inv(K)).astype(np.float32)
        kt = (K @ trel).astype(np.float32)
        min_scale = [1600.0, 576.0, 64.0][level]
        cm = torch.zeros((h, w, 16), dtype=torch.uint8, device=dev)
        cr = torch.zeros((h, w, 16), dtype=torch.uint8, device=dev)
        out2 = torch.zeros(2, dtype=torch.int32, device=dev)
        args = (min_scale, gn[level][0].data_ptr(), gn[level][1].data_ptr(), dl[level].data_ptr(), dn[level].data_ptr(), il[level].data_ptr(),
                inn[level].data_ptr())
        tail = (0.07, fp(np.ascontiguousarray(kt)), fp(np.ascontiguousarray(krk.reshape(-1))), h, w)
        assert lib.slam_op_compute_rgb_residual(*args, cm.data_ptr(), *tail, env["ws"].data_ptr(), out2.data_ptr(), None) == 0
        host2 = (C.c_int * 2)()
        ref.ref_op_compute_rgb_residual(*args, cr.data_ptr(), *tail, host2)
        torch.cuda.synchronize()
        zxm, zym, oxm, oym, dfm, vm = corres_fields(cm.cpu().numpy())
        zxr, zyr, oxr, oyr, dfr, vr = corres_fields(cr.cpu().numpy())
        assert vr.sum() > 200, f"level {level}: degenerate test input ({vr.sum()} correspondences)"
        assert np.array_equal(vm, vr), f"level {level}: correspondence mask differs at {(vm != vr).sum()} pixels"
        for a, b in ((zxm, zxr), (zym, zyr), (oxm, oxr), (oym, oyr)):
            assert np.array_equal(a[vr], b[vr]), f"level {level}: correspondence indices differ"
        assert np.array_equal(dfm[vr], dfr[vr])
        mine2 = out2.cpu().numpy()
        assert mine2[0] == host2[0] and mine2[1] == host2[1], f"level {level}: count/sigma {mine2} vs {host2[0]},{host2[1]}"

        # rgbStep on the reference's correspondence image and point cloud
        cloud = torch.zeros((h, w, 3), dtype=torch.float32, device=dev)
        ref.ref_op_project_to_point_cloud(dl[level].data_ptr(), h, w, cloud.data_ptr(), intr["fx"], intr["fy"], intr["cx"], intr["cy"], level)
        sigma = float(np.float32(np.sqrt(float(host2[0]))))
        out29 = torch.zeros(32, dtype=torch.float32, device=dev)
        sargs = (cr.data_ptr(), sigma, cloud.data_ptr(), fx, fy, gn[level][0].data_ptr(), gn[level][1].data_ptr(), 0.125, h, w)
        assert lib.slam_op_rgb_step(*sargs, env["ws"].data_ptr(), out29.data_ptr(), None) == 0
        host29 = np.zeros(32, dtype=np.float32)
        ref.ref_op_rgb_step(*sargs, fp(host29))
        torch.cuda.synchronize()
        m29 = out29.cpu().numpy()[:29]
        assert m29[28] == host29[28]
        err = se3_sums_rel_err(m29, host29[:29])
        assert err < 1e-4, f"level {level}: rgbStep JtJ/Jtr relative error {err}"


def test_so3_step(env):
    torch, lib, ref, dev, intr = env["torch"], env["lib"], env["ref"], env["dev"], env["intr"]
    _, il, _, inn, _ = rgb_inputs(env)
    level = 2
    h, w = 480 >> level, 640 >> level
    div = np.float32(1 << level)
    fx, fy, cx, cy = (float(np.float32(intr[k]) / div) for k in ("fx", "fy", "cx", "cy"))
    K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], dtype=np.float64)
    ang = 0.01
    R = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]])
    H = (K @ R @ np.linalg.inv(K)).astype(np.float32).reshape(-1)
    Kinv = np.linalg.inv(K).astype(np.float32).reshape(-1)
    KR = (K @ R).astype(np.float32).reshape(-1)
    out = torch.zeros(16, dtype=torch.float32, device=dev)
    args = (il[level].data_ptr(), inn[level].data_ptr(), fp(H), fp(Kinv), fp(KR), h, w)
    assert lib.slam_op_so3_step(*args, env["ws"].data_ptr(), out.data_ptr(), None) == 0
    host = np.zeros(16, dtype=np.float32)
    ref.ref_op_so3_step(*args, fp(host))
    torch.cuda.synchronize()
    m = out.cpu().numpy()[:11]
    assert m[10] == host[10] and host[10] > 1000
    err = so3_sums_rel_err(m, host[:11])
    assert err < 1e-4, f"so3Step relative error {err}"


def random_small_motion(rng, trans=0.02, rot_deg=1.0):
    ax = rng.normal(size=3)
    ax /= np.linalg.norm(ax)
    ang = np.radians(rng.uniform(-rot_deg, rot_deg))
    K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
    R = np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K
    t = rng.uniform(-trans, trans, size=3)
    return R, t


def test_icp_association_fuzz_bit_exact(env):
    """60 random small motions x 307200 pixels: every 1-ulp difference in the association arithmetic (R v + t, projection,
    round-to-nearest pixel, distance / angle gates) would flip ~1e-5 of the pixels, i.e. show up in the inlier count of some
    of the 60 x 3 runs.  Counts must be identical everywhere; sums within 1e-4."""
    rng = np.random.default_rng(1234)
    curr, glob, Rcurr0, tcurr0, Rprev_inv, tprev = icp_inputs(env, perturb=False)
    bad = []
    for trial in range(60):
        R, t = random_small_motion(rng)
        Rcurr = (R @ Rcurr0.astype(np.float64)).astype(np.float32)
        tcurr = (tcurr0.astype(np.float64) + t).astype(np.float32)
        for level in range(3):
            m = run_icp(env, "mine", level, curr, glob, Rcurr, tcurr, Rprev_inv, tprev)
            r = run_icp(env, "ref", level, curr, glob, Rcurr, tcurr, Rprev_inv, tprev)
            if m[28] != r[28] or se3_sums_rel_err(m, r) > 1e-4:
                bad.append((trial, level, m[28], r[28], se3_sums_rel_err(m, r)))
    assert not bad, f"{len(bad)} of 180 runs differ: {bad[:5]}"


def test_rgb_association_fuzz_bit_exact(env):
    """40 random warps: the correspondence image (valid mask, both pixel indices, diff) must be identical pixel by pixel."""
    torch, lib, ref, dev, intr = env["torch"], env["lib"], env["ref"], env["dev"], env["intr"]
    from slam_b200.odometry import corres_fields
    dl, il, dn, inn, gn = rgb_inputs(env)
    rng = np.random.default_rng(99)
    bad = []
    total_valid = 0
    for trial in range(40):
        Rrel, trel = random_small_motion(rng, trans=0.015, rot_deg=0.6)
        level = trial % 3
        h, w = 480 >> level, 640 >> level
        div = np.float32(1 << level)
        fx, fy, cx, cy = (float(np.float32(intr[k]) / div) for k in ("fx", "fy", "cx", "cy"))
        K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], dtype=np.float64)
        krk = np.ascontiguousarray((K @ Rrel @ np.linalg.inv(K)).astype(np.float32).reshape(-1))
        kt = np.ascontiguousarray((K @ trel).astype(np.float32))
        cm = torch.zeros((h, w, 16), dtype=torch.uint8, device=dev)
        cr = torch.zeros((h, w, 16), dtype=torch.uint8, device=dev)
        out2 = torch.zeros(2, dtype=torch.int32, device=dev)
        args = ([1600.0, 576.0, 64.0][level], gn[level][0].data_ptr(), gn[level][1].data_ptr(), dl[level].data_ptr(), dn[level].data_ptr(),
                il[level].data_ptr(), inn[level].data_ptr())
        tail = (0.07, fp(kt), fp(krk), h, w)
        assert lib.slam_op_compute_rgb_residual(*args, cm.data_ptr(), *tail, env["ws"].data_ptr(), out2.data_ptr(), None) == 0
        host2 = (C.c_int * 2)()
        ref.ref_op_compute_rgb_residual(*args, cr.data_ptr(), *tail, host2)
        torch.cuda.synchronize()
        zxm, zym, oxm, oym, dfm, vm = corres_fields(cm.cpu().numpy())
        zxr, zyr, oxr, oyr, dfr, vr = corres_fields(cr.cpu().numpy())
        total_valid += int(vr.sum())
        mine2 = out2.cpu().numpy()
        ok = (np.array_equal(vm, vr) and np.array_equal(zxm[vr], zxr[vr]) and np.array_equal(zym[vr], zyr[vr]) and np.array_equal(dfm[vr], dfr[vr])
              and mine2[0] == host2[0] and mine2[1] == host2[1])
        if not ok:
            bad.append((trial, level, int((vm != vr).sum()), int(mine2[0]), int(host2[0])))
    assert total_valid > 100000
    assert not bad, f"{len(bad)} of 40 warps differ: {bad[:5]}"


def test_so3_fuzz(env):
    torch, lib, ref, dev, intr = env["torch"], env["lib"], env["ref"], env["dev"], env["intr"]
    _, il, _, inn, _ = rgb_inputs(env)
    level = 2
    h, w = 480 >> level, 640 >> level
    div = np.float32(1 << level)
    fx, fy, cx, cy = (float(np.float32(intr[k]) / div) for k in ("fx", "fy", "cx", "cy"))
    K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], dtype=np.float64)
    rng = np.random.default_rng(7)
    bad = []
    for trial in range(60):
        R, _ = random_small_motion(rng, rot_deg=3.0)
        H = np.ascontiguousarray((K @ R @ np.linalg.inv(K)).astype(np.float32).reshape(-1))
        Kinv = np.ascontiguousarray(np.linalg.inv(K).astype(np.float32).reshape(-1))
        KR = np.ascontiguousarray((K @ R).astype(np.float32).reshape(-1))
        out = torch.zeros(16, dtype=torch.float32, device=dev)
        args = (il[level].data_ptr(), inn[level].data_ptr(), fp(H), fp(Kinv), fp(KR), h, w)
        assert lib.slam_op_so3_step(*args, env["ws"].data_ptr(), out.data_ptr(), None) == 0
        host = np.zeros(16, dtype=np.float32)
        ref.ref_op_so3_step(*args, fp(host))
        torch.cuda.synchronize()
        m = out.cpu().numpy()[:11]
        if m[10] != host[10] or so3_sums_rel_err(m, host[:11]) > 1e-4:
            bad.append((trial, m[10], host[10], so3_sums_rel_err(m, host[:11])))
    assert not bad, f"{len(bad)} of 60 differ: {bad[:5]}"
