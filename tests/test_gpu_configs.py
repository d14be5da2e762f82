"""BASELINE.json configs[2]: 1280x720 RealSense-shaped depth, 4-level pyramid, ICP-only tracking with 10/5/4/4 iterations.

The reference class is compiled for three levels (NUM_PYRS, RGBDOdometryef.h:104) but every wrapper it calls is level-agnostic
(odom/utils.cuh:62-175), so the replay (oracle/ref_harness.cu) takes the level count at run time:
  * at 1280x720 with three levels the whole tracker is compared with the reference replay (prepared maps bit-exact, pose);
  * with FOUR levels and 10/5/4/4 iterations the Gauss-Newton chain is compared with the 4-level reference replay step by step
    (same bars as the 3-level tests: first step exact, sums, increments and poses within 1e-5) and every level's prepared maps
    bit for bit;
  * the fourth level's buffers are also compared with the reference's own per-level operators (pyrDown, createVMap,
    createNMap, resizeVMap / resizeNMap, tranformMaps) applied once more, and the 4-level device-resident loop with the
    host-stepped loop and the ground truth.
"""
import ctypes as C

import numpy as np
import pytest

from tests.support import DEPTH_CUTOFF, MODEL_CUTOFF, frame_pair, make_scene, planar_map_mismatch, run_frame, to_device

pytestmark = pytest.mark.gpu

W, H = 1280, 720
ITER4 = (10, 5, 4, 4)


def fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


@pytest.fixture(scope="module")
def hd(built, ref_lib):
    import torch
    from oracle.ref_cuda import RefOdometry
    from slam_b200 import RGBDOdometry
    scene, intr = make_scene(W, H)
    poses = scene.trajectory(1000)
    fr = frame_pair(scene, poses, 200)
    d = to_device(fr)
    torch.cuda.synchronize()
    return dict(torch=torch, scene=scene, intr=intr, poses=poses, fr=fr, d=d, Ref=RefOdometry, Odo=RGBDOdometry, ref=ref_lib)


def test_720p_three_levels_match_reference(hd):
    from slam_b200 import Tap
    i = hd["intr"]
    mine = hd["Odo"](i["width"], i["height"], i["cx"], i["cy"], i["fx"], i["fy"])
    mine.set_trace(1)
    ref = hd["Ref"](i["width"], i["height"], i["cx"], i["cy"], i["fx"], i["fy"])
    kw = dict(so3=False, rgbOnly=False, icpWeight=100.0, pyramid=True, fastOdom=False)
    tm, rm = run_frame(mine, hd["d"], **kw)
    tr, rr = run_frame(ref, hd["d"], **kw)
    for level in range(3):
        for tap in (Tap.VMAP_CURR, Tap.NMAP_CURR, Tap.VMAP_PREV, Tap.NMAP_PREV):
            nan_diff, val_diff = planar_map_mismatch(mine.tap(tap, level), ref.tap(tap, level))
            assert nan_diff == 0 and val_diff == 0, f"tap {tap} level {level}: nan {nan_diff} values {val_diff}"
    assert np.abs(tm - tr).max() < 1e-5 and np.abs(rm - rr).max() < 1e-5
    sm, sr = mine.stats(), ref.stats()
    assert abs(sm.lastICPCount - sr.lastICPCount) <= max(8, 2e-3 * sr.lastICPCount)
    gt = hd["fr"]["gt_pose"]
    assert np.linalg.norm(tm - gt[:3, 3]) < 0.002
    mine.close()
    ref.close()


def reference_level3_maps(hd, taps2):
    """Level-3 buffers from the level-2 taps with the reference's own operators."""
    torch, ref = hd["torch"], hd["ref"]
    dev = "cuda:0"
    i = hd["intr"]
    h2, w2 = H >> 2, W >> 2
    h3, w3 = h2 // 2, w2 // 2
    out = {}
    d2 = torch.from_numpy(taps2["depth"].view(np.int16)).to(dev)
    d3 = torch.zeros((h3, w3), dtype=torch.int16, device=dev)
    ref.ref_op_pyr_down(d2.data_ptr(), h2, w2, d3.data_ptr())
    fx, fy, cx, cy = (float(np.float32(i[k]) / np.float32(8.0)) for k in ("fx", "fy", "cx", "cy"))
    v3 = torch.full((3, h3, w3), float("nan"), dtype=torch.float32, device=dev)
    n3 = torch.full((3, h3, w3), float("nan"), dtype=torch.float32, device=dev)
    ref.ref_op_create_vmap(fx, fy, cx, cy, d3.data_ptr(), h3, w3, v3.data_ptr(), DEPTH_CUTOFF, 1)
    ref.ref_op_create_nmap(v3.data_ptr(), h3, w3, n3.data_ptr(), 1)
    torch.cuda.synchronize()
    out["depth"], out["vcurr"], out["ncurr"] = d3.cpu().numpy().view(np.uint16), v3.cpu().numpy(), n3.cpu().numpy()
    # model maps: camera-frame pyramid from the RGBA32F inputs, then one more resize, then the transform
    d = hd["d"]
    v = torch.full((3, H, W), float("nan"), dtype=torch.float32, device=dev)
    n = torch.full((3, H, W), float("nan"), dtype=torch.float32, device=dev)
    ref.ref_op_copy_maps(d["mv"].data_ptr(), d["mn"].data_ptr(), H, W, v.data_ptr(), n.data_ptr())
    for l in range(1, 4):
        h, w = H >> l, W >> l
        v2 = torch.full((3, h, w), float("nan"), dtype=torch.float32, device=dev)
        n2 = torch.full((3, h, w), float("nan"), dtype=torch.float32, device=dev)
        ref.ref_op_resize_map(v.data_ptr(), h * 2, w * 2, v2.data_ptr(), 0)
        ref.ref_op_resize_map(n.data_ptr(), h * 2, w * 2, n2.data_ptr(), 1)
        v, n = v2, n2
    pose = hd["fr"]["model_pose"]
    R = np.ascontiguousarray(pose[:3, :3], dtype=np.float32).reshape(-1)
    t = np.ascontiguousarray(pose[:3, 3], dtype=np.float32)
    vg, ng = v.clone(), n.clone()
    ref.ref_op_transform_maps(v.data_ptr(), n.data_ptr(), h3, w3, fp(R), fp(t), vg.data_ptr(), ng.data_ptr())
    torch.cuda.synchronize()
    out["vprev"], out["nprev"] = vg.cpu().numpy(), ng.cpu().numpy()
    return out


def test_720p_four_levels(hd):
    from slam_b200 import Tap
    i = hd["intr"]
    kw = dict(so3=False, rgbOnly=False, icpWeight=100.0, pyramid=True, fastOdom=False)
    host = hd["Odo"](i["width"], i["height"], i["cx"], i["cy"], i["fx"], i["fy"], num_levels=4, iterations=ITER4, host_loop=True)
    host.set_trace(True)
    th, rh = run_frame(host, hd["d"], **kw)
    # ---- the fourth level's prepared buffers against the reference's operators
    taps2 = dict(depth=host.tap(Tap.DEPTH_U16, 2))
    want = reference_level3_maps(hd, taps2)
    assert np.array_equal(host.tap(Tap.DEPTH_U16, 3), want["depth"])
    for tap, key in ((Tap.VMAP_CURR, "vcurr"), (Tap.NMAP_CURR, "ncurr"), (Tap.VMAP_PREV, "vprev"), (Tap.NMAP_PREV, "nprev")):
        nan_diff, val_diff = planar_map_mismatch(host.tap(tap, 3), want[key])
        assert nan_diff == 0 and val_diff == 0, f"{key} level 3: nan {nan_diff} values {val_diff}"
    trh = host.get_trace()
    assert [r["level"] for r in trh] == [3] * 4 + [2] * 4 + [1] * 5 + [0] * 10
    # ---- device-resident 4-level loop == host-stepped loop
    devo = hd["Odo"](i["width"], i["height"], i["cx"], i["cy"], i["fx"], i["fy"], num_levels=4, iterations=ITER4)
    devo.set_trace(1)
    td, rd = run_frame(devo, hd["d"], **kw)
    trd = devo.get_trace()
    assert len(trd) == len(trh)
    assert trd[0]["icp"][28] == trh[0]["icp"][28]          # first step: same inlier mask
    assert np.abs(td - th).max() < 1e-5 and np.abs(rd - rh).max() < 1e-5
    gt = hd["fr"]["gt_pose"]
    assert np.linalg.norm(td - gt[:3, 3]) < 0.002
    # ---- and the one-call-per-frame entry point (fast path, no trace)
    fast = hd["Odo"](i["width"], i["height"], i["cx"], i["cy"], i["fx"], i["fy"], num_levels=4, iterations=ITER4)
    d = hd["d"]
    frame = fast.make_frame(d["depth"], d["rgba"], d["mv"], d["mn"], d["mrgba"], d["model_pose"], DEPTH_CUTOFF, MODEL_CUTOFF)
    pose = d["model_pose"]
    tf, rf = fast.track_device(frame, pose[:3, 3].copy(), pose[:3, :3].copy(), False, 100.0, True, False, False)
    assert np.array_equal(tf, td) and np.array_equal(rf, rd)
    for o in (host, devo, fast):
        o.close()


def test_720p_four_levels_match_the_four_level_reference_replay(hd):
    """configs[2] against the reference itself: the replay harness run with four levels and 10/5/4/4 iterations."""
    from slam_b200 import Tap
    from tests.test_gpu_tracking import compare_traces
    i = hd["intr"]
    kw = dict(so3=False, rgbOnly=False, icpWeight=100.0, pyramid=True, fastOdom=False)
    ref = hd["Ref"](i["width"], i["height"], i["cx"], i["cy"], i["fx"], i["fy"], iterations=ITER4, num_levels=4)
    ref.set_trace(True)
    tr, rr = run_frame(ref, hd["d"], **kw)
    ref_trace = ref.get_trace()
    assert [r["level"] for r in ref_trace] == [3] * 4 + [2] * 4 + [1] * 5 + [0] * 10
    for which in ("host loop", "device loop"):
        mine = hd["Odo"](i["width"], i["height"], i["cx"], i["cy"], i["fx"], i["fy"], num_levels=4, iterations=ITER4, host_loop=(which == "host loop"))
        mine.set_trace(True)
        tm, rm = run_frame(mine, hd["d"], **kw)
        for level in range(4):
            for tap in (Tap.VMAP_CURR, Tap.NMAP_CURR, Tap.VMAP_PREV, Tap.NMAP_PREV):
                nan_diff, val_diff = planar_map_mismatch(mine.tap(tap, level), ref.tap(tap, level))
                assert nan_diff == 0 and val_diff == 0, f"{which}: tap {tap} level {level}: nan {nan_diff} values {val_diff}"
            assert np.array_equal(mine.tap(Tap.DEPTH_U16, level), ref.tap(Tap.DEPTH_U16, level)), f"{which}: depth pyramid level {level}"
        compare_traces(mine.get_trace(), ref_trace, f"720p 4-level {which}")
        assert np.abs(tm - tr).max() < 1e-5 and np.abs(rm - rr).max() < 1e-5, f"{which}: pose differs from the 4-level reference replay"
        sm, sr = mine.stats(), ref.stats()
        assert abs(sm.lastICPCount - sr.lastICPCount) <= max(8, 2e-3 * sr.lastICPCount)
        mine.close()
    ref.close()


# ------------------------------------------------------------------------------------------------------------------
# BASELINE.json configs[4]: relocalisation scoring, 256 pose hypotheses per frame, each scored by the ICP residual
# reduction; here on one GPU (the sharding + min-allreduce logic is covered on CPU by tests/test_synth_and_sharding.py).
def perturbed_poses(pose, n, seed=7):
    """n candidate poses around `pose`: hypothesis 0 is the pose itself, the others are displaced by 5 .. 60 mm / up to 3 deg."""
    def rodrigues_np(w):
        th = np.linalg.norm(w)
        if th < 1e-12:
            return np.eye(3)
        k = w / th
        K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
        return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K

    rng = np.random.default_rng(seed)
    T = np.repeat(pose[:3, 3][None].astype(np.float32), n, 0)
    R = np.repeat(pose[:3, :3][None].astype(np.float32), n, 0)
    for k in range(1, n):
        d = rng.normal(size=3)
        T[k] += (d / np.linalg.norm(d) * rng.uniform(0.005, 0.06)).astype(np.float32)
        w = rng.normal(size=3)
        w = w / np.linalg.norm(w) * np.deg2rad(rng.uniform(0.0, 3.0))
        R[k] = (rodrigues_np(w) @ pose[:3, :3]).astype(np.float32)
    return T, R


def test_pose_hypothesis_scores_match_reference_icp_step(built, ref_lib, icl_sequence):
    import torch
    from slam_b200 import RGBDOdometry, Tap
    from slam_b200.relocalise import icp_error, score_sharded
    from tests.support import ANGLE_THRESH
    scene, intr, poses = icl_sequence
    fr = frame_pair(scene, poses, 500)
    d = to_device(fr)
    torch.cuda.synchronize()
    odo = RGBDOdometry(intr["width"], intr["height"], intr["cx"], intr["cy"], intr["fx"], intr["fy"])
    odo.set_trace(1)   # keeps the prepared maps readable through the taps
    odo.initICPModel(d["mv"], d["mn"], MODEL_CUTOFF, d["model_pose"])
    odo.initICP(d["depth"], DEPTH_CUTOFF)
    model = fr["model_pose"].astype(np.float32)
    N = 256
    T, R = perturbed_poses(fr["gt_pose"].astype(np.float32), N)
    for level in (2, 0):
        res, cnt = odo.score_poses(level, model, T, R)
        assert res.shape == (N,) and np.all(cnt > 0)
        # ---- the reference's icpStep on the same maps, same poses (a sample of the hypotheses)
        h, w = 480 >> level, 640 >> level
        div = np.float32(1 << level)
        fx, fy, cx, cy = (float(np.float32(intr[k]) / div) for k in ("fx", "fy", "cx", "cy"))
        maps = [torch.from_numpy(odo.tap(t, level)).to("cuda:0") for t in (Tap.VMAP_CURR, Tap.NMAP_CURR, Tap.VMAP_PREV, Tap.NMAP_PREV)]
        Rprev_inv = np.linalg.inv(model[:3, :3].astype(np.float64)).astype(np.float32)
        tprev = model[:3, 3].copy()
        for k in (0, 1, 17, 100, 255):
            host = np.zeros(32, dtype=np.float32)
            ref_lib.ref_op_icp_step(fp(np.ascontiguousarray(R[k].reshape(-1))), fp(np.ascontiguousarray(T[k])), maps[0].data_ptr(), maps[1].data_ptr(),
                                    fp(np.ascontiguousarray(Rprev_inv.reshape(-1))), fp(tprev), fx, fy, cx, cy, maps[2].data_ptr(), maps[3].data_ptr(), 0.10,
                                    float(np.float32(ANGLE_THRESH)), h, w, fp(host))
            assert cnt[k] == host[28], f"level {level} hypothesis {k}: inliers {cnt[k]} vs {host[28]}"
            assert abs(res[k] - host[27]) <= 1e-4 * abs(host[27]), f"level {level} hypothesis {k}: residual {res[k]} vs {host[27]}"
        # ---- the unperturbed pose wins, and the error grows with the displacement
        err = icp_error(res, cnt, min_inliers=0.2 * h * w)
        best, e, _ = score_sharded(odo, level, model, T, R, min_inliers=0.2 * h * w)
        assert best == int(np.argmin(err)) and abs(e - err.min()) < 1e-12
        assert best == 0, f"level {level}: hypothesis {best} beats the true pose ({err[best]} vs {err[0]})"
        # deterministic run to run
        res2, cnt2 = odo.score_poses(level, model, T, R)
        assert np.array_equal(res, res2) and np.array_equal(cnt, cnt2)
    odo.close()


def test_device_side_best_key_equals_host_packing_for_any_sharding(built, icl_sequence):
    """slam_odom_score_poses_best (product path of configs[4]): the scoring launch folds the packed key of its block's best
    hypothesis into one device word.  Whatever the number of blocks the 256 hypotheses are cut into (1, 2, 4, 8 "ranks",
    emulated one after the other on this GPU with the min taken over their words), the winner is the one the host-side packing
    of all per-hypothesis sums picks: same index, same float32 error bits."""
    import torch
    from slam_b200 import RGBDOdometry
    from slam_b200.relocalise import INT64_MAX, icp_error, pack_keys, perturbed_hypotheses, shard_range, unpack_key
    scene, intr, poses = icl_sequence
    fr = frame_pair(scene, poses, 640)
    d = to_device(fr)
    torch.cuda.synchronize()
    odo = RGBDOdometry(intr["width"], intr["height"], intr["cx"], intr["cy"], intr["fx"], intr["fy"])
    odo.initICPModel(d["mv"], d["mn"], MODEL_CUTOFF, d["model_pose"])
    odo.initICP(d["depth"], DEPTH_CUTOFF)
    model = fr["model_pose"].astype(np.float32)
    N = 256
    T, R = perturbed_hypotheses(fr["gt_pose"], N)     # sigma_t 5 cm, sigma_r 3 deg, seed 0xBEEF, hypothesis 0 = gt
    assert np.allclose(T[0], fr["gt_pose"][:3, 3], atol=1e-7) and np.allclose(R[0], fr["gt_pose"][:3, :3], atol=1e-7)
    stream = torch.cuda.ExternalStream(odo.stream)
    for level in (0, 2):
        min_inl = 1400 >> (2 * level)
        res, cnt = odo.score_poses(level, model, T, R)
        keys = pack_keys(icp_error(res, cnt, min_inl), np.arange(N))
        want = unpack_key(int(keys.min()))
        for world in (1, 2, 4, 8):
            words = []
            for rank in range(world):
                lo, hi = shard_range(N, rank, world)
                key = torch.full((1,), INT64_MAX, dtype=torch.int64, device="cuda:0")
                with torch.cuda.stream(stream):
                    odo.score_poses_best(level, model, T[lo:hi], R[lo:hi], key, index_base=lo, min_inliers=min_inl)
                    words.append(int(key.item()))
            got = unpack_key(min(words))
            assert got[1] == want[1] and np.float32(got[0]).view(np.uint32) == np.float32(want[0]).view(np.uint32), f"level {level}, {world} blocks: {got} vs {want}"
        assert want[1] == 0, f"level {level}: hypothesis {want[1]} beats the true pose"
    # a guard nobody passes leaves the word at +inf | smallest index
    key = torch.full((1,), INT64_MAX, dtype=torch.int64, device="cuda:0")
    with torch.cuda.stream(stream):
        odo.score_poses_best(0, model, T[:4], R[:4], key, index_base=10, min_inliers=1e9)
        e, i = unpack_key(int(key.item()))
    assert np.isinf(e) and i == 10
    odo.close()


def _nccl_worker(rank, world, port, out):
    import os
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    from slam_b200 import RGBDOdometry
    from slam_b200.relocalise import INT64_MAX, broadcast_frame, connect_peers, perturbed_hypotheses, score_sharded, score_sharded_device, score_sharded_peers
    from tests.support import make_scene
    scene, intr = make_scene(640, 480)
    poses = scene.trajectory(1000)
    # every rank renders a DIFFERENT frame; rank 0's is the one that counts
    fr = frame_pair(scene, poses, 300 + 37 * rank)
    dev = f"cuda:{rank}"
    d = to_device(fr, dev)
    meta = torch.from_numpy(np.concatenate([fr["model_pose"].reshape(-1), fr["gt_pose"].reshape(-1)]).astype(np.float64)).to(dev)
    broadcast_frame([d["depth"], d["mv"], d["mn"], meta])
    meta = meta.cpu().numpy()
    model, gt = meta[:16].reshape(4, 4).astype(np.float32), meta[16:].reshape(4, 4)
    odo = RGBDOdometry(intr["width"], intr["height"], intr["cx"], intr["cy"], intr["fx"], intr["fy"], device=rank)
    odo.initICPModel(d["mv"], d["mn"], MODEL_CUTOFF, model)
    odo.initICP(d["depth"], DEPTH_CUTOFF)
    T, R = perturbed_hypotheses(gt, 256)
    key = torch.full((1,), INT64_MAX, dtype=torch.int64, device=dev)
    stream = torch.cuda.ExternalStream(odo.stream, device=dev)
    res = {}
    peers = connect_peers(odo, rank, world)
    for level in (0, 2):
        best, err = score_sharded_device(odo, level, model, T, R, key, rank, world, min_inliers=1400 >> (2 * level), stream=stream)
        alone = score_sharded(odo, level, model, T, R, 0, 1, min_inliers=1400 >> (2 * level))
        res[level] = (best, float(np.float32(err)), alone[0], float(np.float32(alone[1])))
        if peers:
            # the same decision over NVLink peer memory (three frames in a row: both slot parities and their reuse), and a frame
            # in which no hypothesis is acceptable anywhere (every rank publishes INT64_MAX)
            for _ in range(3):
                pb, pe = score_sharded_peers(odo, level, model, T, R, rank, world, min_inliers=1400 >> (2 * level))
                assert (pb, float(np.float32(pe))) == res[level][:2], f"rank {rank} level {level}: peers {pb} / {pe} != NCCL {res[level][:2]}"
            assert odo.score_poses_best_peers(level, model, T[:0], R[:0]) == INT64_MAX   # collective: both ranks call it
    res["peers"] = peers
    out[rank] = res
    odo.close()
    dist.destroy_process_group()


def test_hypotheses_sharded_over_two_gpus_with_nccl(built):
    """configs[4] on hardware: two ranks, two GPUs, rank 0's frame broadcast with NCCL, 128 hypotheses scored per rank, ONE
    min-all-reduce of a device word; every rank must name the winner a single GPU names when it scores all 256 alone."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_nccl_worker, args=(2, 29611, out), nprocs=2, join=True)
    assert len(out) == 2
    assert out[0]["peers"] == out[1]["peers"]
    for level in (0, 2):
        b0, e0, a0, ae0 = out[0][level]
        b1, e1, a1, ae1 = out[1][level]
        assert (b0, e0) == (b1, e1) == (a0, ae0) == (a1, ae1), f"level {level}: {out[0][level]} vs {out[1][level]}"
