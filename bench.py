#!/usr/bin/env python
"""bench.py -- RGB-D camera-tracking throughput on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W                 # our arm (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K --warmup W   # reference arm (host cores)
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...  # N > 1, one rank per GPU

Workload (BASELINE.json configs[1]): ICP+RGB+SO3 frame-to-model odometry over a synthetic ICL-NUIM-shaped
640x480 sequence (3-level pyramid, 10/5/4 iterations, icpWeight 10, so3 on).  A *step* is one tracked frame:
initICPModel -> initRGBModel -> initICP(depth) -> initRGB -> getIncrementalTransformation, i.e. everything
RGBDOdometryef does per frame (apps/elastic_fusion_file.cpp:366-374).  Inputs are open-loop (the model
prediction of frame k is ray-cast at the ground-truth pose of frame k-1, outside the timed region), so every
rank can replay its own sequence without a renderer in the loop.

  value  frames/s, whole job, inputs already resident in HBM (device pointers through the C ABI)
  e2e    frames/s through the host-buffer entry point of the C ABI (slam_odom_track_host): every step's depth,
         RGB and model maps are copied from pinned host memory (H2D, prefetched one frame ahead on a copy
         stream) and the pose is read back (D2H) inside the timed region
  roofline     persistent Gauss-Newton kernel: algorithmic bytes (48 + 30 + 32 B per
               pixel-iteration, 2 B per SO3 pixel-iteration) / its CUDA-event duration, vs the measured HBM peak
  batched      (extra) BASELINE.json configs[3]: 64 independent sequences, 64 / N per GPU in one batched handle; frames/s and
               the HBM roofline of its ICP/RGB reduction launches
  cpu_baseline the CPU port (oracle/odom_oracle.c, OpenMP, all host cores) on a bounded sample of the same frames
  ref_cuda     (extra) the reference's own kernels + launch/sync pattern (oracle/_ref) on the same GPU and frames:
               the denominator of the north-star's ">= 20x the reference's own CUDA path"

Multi-GPU: independent sequences, one per rank (weak scaling), no collective on the data path; the only
collectives are the barrier around the timed region and a max-reduce of the elapsed time.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

W, H = 640, 480
LEVELS = 3
ITERS = (10, 5, 4)
N_FRAMES_DISTINCT = 96          # distinct frames cycled through: 96 x 12.9 MB = 1.24 GB of inputs >> 126 MB of L2
DEPTH_CUTOFF, MODEL_CUTOFF = 3.0, 20.0
BYTES_PER_FRAME_IN = W * H * (2 + 4 + 16 + 16 + 4)      # depth u16 + rgba + vertices + normals + model rgba
BYTES_SENSOR_IN = W * H * (2 + 4)                          # what the sensor delivers per frame: depth u16 + rgba


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# stdout carries exactly ONE line (the JSON record): libraries that print there from native code (NCCL's version banner,
# the reference kernels' "performance database" notes) are diverted to stderr by re-pointing fd 1 for the run.
_REAL_STDOUT = None


def divert_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        os.write(1, data)
    else:
        os.write(_REAL_STDOUT, data)


def algorithmic_bytes_per_frame(so3_iters: float) -> float:
    """SURVEY.md 8(d) / BASELINE.md 3: ICP 48 + RGB residual 30 + RGB step 32 B per pixel-iteration; SO3 2 B/px-iteration."""
    px_iter = sum((W >> l) * (H >> l) * ITERS[l] for l in range(LEVELS))
    return px_iter * 110.0 + so3_iters * (W >> 2) * (H >> 2) * 2.0


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def make_frames(seed: int, n: int):
    """n open-loop frame pairs of one synthetic sequence (numpy, host)."""
    from slam_b200.synth import Scene
    scene = Scene(seed=0x51A7)
    poses = scene.trajectory(1000, seed=0x51A7 + seed)
    frames = []
    for i in range(n):
        k = 100 + i          # consecutive frames: the SO3 pre-alignment compares each image with the one tracked just before
        depth, rgba = scene.render_frame(poses[k])
        mv, mn, mrgba = scene.render_model(poses[k - 1])
        frames.append(dict(depth=depth, rgba=rgba, mv=mv, mn=mn, mrgba=mrgba, model_pose=poses[k - 1].copy(), gt_pose=poses[k].copy()))
    first_rgba = scene.render_frame(poses[99])[1]
    return frames, first_rgba


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    QUERY = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for t, line in self.samples:
            if t < t0 - 0.05 or t > t1 + 0.15:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:   # region shorter than the sampling period: take whatever we have
            for t, line in self.samples[-3:]:
                f = [x.strip() for x in line.split(",")]
                try:
                    sm.append(float(f[0]))
                    mx.append(float(f[1]))
                except Exception:
                    pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU arms
def cpu_port_fps(frames, first_rgba, n_frames: int):
    """The CPU port on all host cores over n_frames of the workload; returns (frames/s, cores)."""
    from oracle.cpu_oracle import CpuOdometry, load
    lib = load()
    cores = lib.oracle_num_threads()
    odo = CpuOdometry(W, H, 319.5, 239.5, 481.20, -480.0)
    odo.initFirstRGB(first_rgba)

    def one(fr):
        odo.initICPModel(fr["mv"], fr["mn"], MODEL_CUTOFF, fr["model_pose"])
        odo.initRGBModel(fr["mrgba"])
        odo.initICP(fr["depth"], DEPTH_CUTOFF)
        odo.initRGB(fr["rgba"])
        p = fr["model_pose"]
        return odo.getIncrementalTransformation(p[:3, 3].copy(), p[:3, :3].copy(), False, 10.0, True, False, True)

    one(frames[0])   # warm-up (page faults, OpenMP pool)
    t0 = time.perf_counter()
    for i in range(n_frames):
        one(frames[i % len(frames)])
    dt = time.perf_counter() - t0
    odo.close()
    return n_frames / dt, cores


def run_reference_arm(args, rank, world):
    """--impl reference: the reference has no CPU tracking path; its reduction math restated for the host
    (oracle/odom_oracle.c, kind "port") runs on all host cores.  Rank 0 only; other ranks exit."""
    if rank != 0:
        return
    frames, first_rgba = make_frames(0, 8)
    # a step = one frame (~0.15 s on 8 cores): K steps + W warm-up stay within minutes for the driver's K
    cpu_port_fps(frames, first_rgba, min(3, max(1, args.warmup)))   # warm-up
    steps = max(1, args.steps)
    fps, cores = cpu_port_fps(frames, first_rgba, steps)
    line = {
        "impl": "reference", "metric": "ICP+RGB tracking frames/sec @640x480", "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 / fps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(1),
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                         "sample": f"{steps} frames of the workload (full per-frame path: pyramids, maps, SO3 + 19 ICP+RGB iterations), oracle/odom_oracle.c, OpenMP"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "the reference publishes no CPU path for RGBDOdometryef; this is its reduction math restated in C on the host cores (context only)",
    }
    emit(line)


def workload_config(n_gpus):
    return {"workload": "configs[1]: ICP+RGB+SO3 frame-to-model odometry, synthetic ICL-NUIM-shaped 640x480 sequence, 3-level pyramid, 10/5/4 iterations, "
                        "icpWeight 10, open-loop inputs", "frames_distinct": N_FRAMES_DISTINCT, "sequences": n_gpus,
            "l2_policy": f"inputs larger than L2: {N_FRAMES_DISTINCT} distinct frames x {BYTES_PER_FRAME_IN / 1e6:.1f} MB cycled", "parallelism": f"1 sequence per GPU x {n_gpus} (every rank tracks its own copy of the same sequence)",
            "pre_roll": "the W warm-up steps, then 0.25 s of untimed tracking while the clock sampler starts (the device does not idle in front of the timed region)"}


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    from slam_b200 import RGBDOdometry

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = f"cuda:{local_rank}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))

    t_gen = time.perf_counter()
    # Every rank tracks its own copy of the SAME synthetic sequence: weak scaling wants identical work per GPU, and the time of a
    # frame depends on its content (SO3 iterations until convergence, valid pixels) -- with a different trajectory per rank the
    # max-over-ranks time measured the slowest trajectory (+17 us per frame on rank 1 of 2), not the system.
    frames, first_rgba = make_frames(0, N_FRAMES_DISTINCT)
    log(f"[rank {rank}] generated {len(frames)} synthetic frames in {time.perf_counter() - t_gen:.1f}s")

    # ---- device-resident copies of all inputs
    def up(a):
        a = a.view(np.int16) if a.dtype == np.uint16 else a
        return torch.from_numpy(a).to(dev)

    dframes = [{k: (up(v) if k not in ("model_pose", "gt_pose") else v) for k, v in fr.items()} for fr in frames]
    dfirst = up(first_rgba)
    # ---- pinned host copies for the end-to-end arm
    def pin(a):
        a = a.view(np.int16) if a.dtype == np.uint16 else a
        return torch.from_numpy(a).pin_memory()

    hframes = [{k: (pin(v) if k not in ("model_pose", "gt_pose") else v) for k, v in fr.items()} for fr in frames]
    torch.cuda.synchronize()

    odo = RGBDOdometry(W, H, 319.5, 239.5, 481.20, -480.0, device=local_rank)
    odo.initFirstRGB(dfirst)
    dev_frames = [odo.make_frame(d["depth"], d["rgba"], d["mv"], d["mn"], d["mrgba"], d["model_pose"], DEPTH_CUTOFF, MODEL_CUTOFF) for d in dframes]
    host_frames = [odo.make_frame(d["depth"], d["rgba"], d["mv"], d["mn"], d["mrgba"], d["model_pose"], DEPTH_CUTOFF, MODEL_CUTOFF) for d in hframes]
    # the reference's data flow (apps/elastic_fusion_file.cpp:301-374): the sensor frame arrives in host memory every frame, the
    # model prediction is rendered on the GPU and never leaves it
    sensor_frames = [odo.make_frame(hd["depth"], hd["rgba"], d["mv"], d["mn"], d["mrgba"], d["model_pose"], DEPTH_CUTOFF, MODEL_CUTOFF) for hd, d in zip(hframes, dframes)]
    priors = [(fr["model_pose"][:3, 3].copy(), fr["model_pose"][:3, :3].copy()) for fr in frames]
    nf = len(frames)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    stream = torch.cuda.ExternalStream(odo.stream, device=dev)

    # ================= value: inputs resident in HBM =================
    for i in range(args.warmup):
        odo.track_device(dev_frames[i % nf], *priors[i % nf])
    sampler = ClockSampler(local_rank)
    sampler.start()
    # nvidia-smi needs ~0.25 s to deliver its first sample: the device keeps tracking (untimed) meanwhile instead of idling, so that
    # the timed region starts from the loaded state (clocks, caches, the library's pool of timing events) that a running tracker is in
    odo.set_profiling(True)
    t_pre = time.perf_counter()
    i_pre = 0
    while time.perf_counter() - t_pre < 0.25:
        odo.track_device(dev_frames[i_pre % nf], *priors[i_pre % nf])
        i_pre += 1
    odo.set_profiling(True)
    odo.get_profile(reset=True)
    launches0 = odo.launch_count()
    import gc
    gc.collect()
    gc.disable()      # no collector pause inside the timed regions (a 20-step region lasts 3.5 ms)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter()
    e0.record(stream)
    so3_iters = 0
    last = None
    for i in range(args.steps):
        j = (args.warmup + i) % nf
        last = odo.track_device(dev_frames[j], *priors[j])
    e1.record(stream)
    barrier()
    w1 = time.perf_counter()
    dev_ms = e0.elapsed_time(e1)
    clocks = sampler.stop(w0, w1)
    log(f"[rank {rank}] value loop: {dev_ms / args.steps * 1e3:.1f} us per step on this rank")
    gn_ms, gn_launches = odo.get_profile(reset=True)
    odo.set_profiling(False)
    launches = odo.launch_count() - launches0
    so3_iters = odo.stats().so3_iterations
    elapsed_ms = max_over_ranks(dev_ms)
    value = world * args.steps / (elapsed_ms / 1e3)
    # sanity: the last tracked pose must be close to the ground truth of that frame (no work skipped)
    jlast = (args.warmup + args.steps - 1) % nf
    err_mm = float(np.linalg.norm(last[0] - frames[jlast]["gt_pose"][:3, 3]) * 1e3)
    prior_mm = float(np.linalg.norm(priors[jlast][0] - frames[jlast]["gt_pose"][:3, 3]) * 1e3)

    # ================= e2e: host buffers through the C ABI =================
    def time_e2e(track, frames_):
        # the warm-up steps run the same pipeline as the timed ones (the next frame's H2D copies are issued behind this frame's kernels),
        # so the first timed step finds its inputs prefetched like every later one
        for i in range(args.warmup):
            track(frames_[i % nf], *priors[i % nf], next_frame=frames_[(i + 1) % nf])
        barrier()
        t0 = time.perf_counter()
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2.record(stream)
        for i in range(args.steps):
            j = (args.warmup + i) % nf
            jn = (args.warmup + i + 1) % nf
            # this frame's kernels are enqueued, then the next frame's H2D copies are issued (they overlap this frame's solve),
            # then the pose of this frame is read back (D2H)
            track(frames_[j], *priors[j], next_frame=frames_[jn] if i + 1 < args.steps else None)
        e3.record(stream)
        barrier()
        wall_ms = (time.perf_counter() - t0) * 1e3        # host wall clock of this rank: what the caller of the API sees
        return max_over_ranks(max(e2.elapsed_time(e3), wall_ms))

    # headline: the sensor frame (u16 depth + RGBA8, 1.84 MB) comes from pinned host memory every step, the model prediction is
    # device-resident as in the reference; pose read back every step
    e2e_ms = time_e2e(odo.track_sensor, sensor_frames)
    e2e_value = world * args.steps / (e2e_ms / 1e3)
    # pessimistic line: the 11 MB model prediction travels over PCIe as well (slam_odom_track_host)
    e2e_all_ms = time_e2e(odo.track_host, host_frames)
    e2e_all_value = world * args.steps / (e2e_all_ms / 1e3)
    gc.enable()

    # ================= roofline of the dominant kernel =================
    peak, peak_src = measured_peak()
    alg_bytes = algorithmic_bytes_per_frame(so3_iters)
    gn_us = gn_ms / max(1, gn_launches) * 1e3
    achieved = alg_bytes / (gn_us * 1e-6) / 1e9 if gn_launches else None
    traffic = None
    tp = ROOT / "profiles" / "gn_kernel_traffic.json"
    if tp.exists():
        try:
            traffic = json.loads(tp.read_text()).get("dram_bytes_per_launch")
        except Exception:
            traffic = None

    # ================= batched extra: BASELINE.json configs[3] =================
    batched = None
    if not args.no_batched:
        try:
            batched = run_batched(args, rank, local_rank, world, dframes, hframes, dfirst, frames, barrier, max_over_ranks, peak, peak_src)
        except Exception as e:   # pragma: no cover
            batched = {"value": None, "note": f"failed: {e}"}

    extras = None
    if not args.no_batched:
        try:
            extras = run_other_configs(args, rank, local_rank, world, odo, dframes, frames, barrier, max_over_ranks)
        except Exception as e:   # pragma: no cover
            extras = {"note": f"failed: {e}"}

    line = None
    if rank == 0:
        # ---- context baselines (N = 1 only): CPU port and the reference's own CUDA path
        cpu = None
        ref_cuda = None
        if world == 1 and not args.no_baselines:
            try:
                n_cpu = 60
                fps_cpu, cores = cpu_port_fps(frames[:8], first_rgba, n_cpu)
                cpu = {"value": fps_cpu, "unit": "frames/s", "cores": cores, "kind": "port",
                       "sample": f"{n_cpu} frames of the same workload (full per-frame path), oracle/odom_oracle.c with OpenMP on all host cores"}
            except Exception as e:   # pragma: no cover
                cpu = {"value": None, "unit": "frames/s", "cores": None, "kind": "port", "sample": f"failed: {e}"}
            try:
                from oracle import ref_cuda as rc
                if rc.available():
                    ref = rc.RefOdometry(W, H, 319.5, 239.5, 481.20, -480.0)
                    ref.initFirstRGB(dfirst)

                    def ref_frame(d):
                        ref.initICPModel(d["mv"], d["mn"], MODEL_CUTOFF, d["model_pose"])
                        ref.initRGBModel(d["mrgba"])
                        ref.initICP(d["depth"], DEPTH_CUTOFF)
                        ref.initRGB(d["rgba"])
                        p = d["model_pose"]
                        return ref.getIncrementalTransformation(p[:3, 3].copy(), p[:3, :3].copy(), False, 10.0, True, False, True)

                    for i in range(10):
                        ref_frame(dframes[i % nf])
                    torch.cuda.synchronize()
                    nref = 100
                    t0 = time.perf_counter()
                    for i in range(nref):
                        ref_frame(dframes[(10 + i) % nf])
                    torch.cuda.synchronize()
                    ref_cuda = {"value": nref / (time.perf_counter() - t0), "unit": "frames/s",
                                "kind": "reference kernels (src/odom/*.cu compiled for sm_100a) with the reference's launch/sync/malloc pattern, same GPU, same frames",
                                "sample": f"{nref} frames, device-resident inputs"}
                    ref.close()
            except Exception as e:   # pragma: no cover
                ref_cuda = {"value": None, "note": f"failed: {e}"}

        line = {
            "metric": "ICP+RGB tracking frames/sec @640x480", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(world),
            "icp_iterations_per_s": value * sum(ITERS),
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": BYTES_SENSOR_IN + 64, "d2h_bytes_per_step": 48,
                    "ms_per_step": e2e_ms / args.steps,
                    "note": "slam_odom_track_sensor: depth + RGBA of every step from pinned host memory, model prediction device-resident (the "
                            "reference renders it into GL textures and never moves it), pose read back"},
            "e2e_all_host": {"value": e2e_all_value, "unit": "frames/s", "h2d_bytes_per_step": BYTES_PER_FRAME_IN + 64, "d2h_bytes_per_step": 48,
                             "ms_per_step": e2e_all_ms / args.steps, "note": "slam_odom_track_host: the 11 MB model prediction crosses PCIe too (PCIe-bound)"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "k_gn_persistent", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                         "traffic": traffic, "peak_source": peak_src, "us_per_launch": gn_us, "launches": int(gn_launches),
                         "algorithmic_bytes_per_launch": alg_bytes, "share_of_step": (gn_ms / dev_ms) if dev_ms else None,
                         "launch_shape": "split: SO3 on one 16-CTA cluster + fine-level kernel on the other SMs, timed as one bracket" if launches == 3 * args.steps else "one cooperative launch",
                         "note": "one bracket = all SO3 + 19 ICP/RGB iterations of a frame (split launch: the cluster kernel and the fine-level kernel that runs "
                                 "next to it); the 45 MB working set stays in the 126 MB L2, so DRAM traffic is far below the algorithmic bytes and the path "
                                 "is latency-bound (a chain of 23-29 dependent reductions + fp64 solves), not bandwidth-bound"},
            "batched": batched,
            "other_configs": extras,
            "cpu_baseline": cpu,
            "ref_cuda": ref_cuda,
            "clocks": clocks,
            "check": {"last_frame_error_mm": err_mm, "prior_error_mm": prior_mm},
        }
    odo.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        emit(line)


def run_batched(args, rank, local_rank, world, dframes, hframes, dfirst, frames, barrier, max_over_ranks, peak, peak_src):
    """BASELINE.json configs[3]: 64 independent 640x480 sequences, 64 / N per GPU, one handle with batch = 64 / N per rank
    (the batched streaming engine: lock-step map-reduce launches over all sequences).  A step = one frame of every
    sequence.  Reported: whole-job frames/s, and the HBM roofline of the ICP/RGB reduction launches (phase A + phase B),
    timed by CUDA events on the handle's stream inside the library."""
    import torch
    from slam_b200 import RGBDOdometry
    dev = f"cuda:{local_rank}"
    total = 64
    B = max(1, total // world)
    nf = len(dframes)
    keys = ("depth", "rgba", "mv", "mn", "mrgba")
    odo = RGBDOdometry(W, H, 319.5, 239.5, 481.20, -480.0, device=local_rank, batch=B)
    odo.initFirstRGB(torch.stack([dfirst] * B))
    sets = []
    for off in (0, 1):     # two consecutive frames of every sequence (2 x 826 MB at B = 64) + a 2.6 GB arena: far beyond L2
        d = {k: torch.stack([dframes[(off + 5 * b) % nf][k] for b in range(B)]) for k in keys}
        P = np.stack([frames[(off + 5 * b) % nf]["model_pose"] for b in range(B)])
        G = np.stack([frames[(off + 5 * b) % nf]["gt_pose"][:3, 3] for b in range(B)])
        sets.append((d, P, odo.make_frame(d["depth"], d["rgba"], d["mv"], d["mn"], d["mrgba"], P, DEPTH_CUTOFF, MODEL_CUTOFF), G))
    torch.cuda.synchronize()
    steps = max(16, min(40, args.steps // 16))   # at least 16 batched steps whatever --steps says
    warm = 3

    def step(i):
        d, P, fr, G = sets[i % 2]
        return odo.track_device(fr, P[:, :3, 3].copy(), P[:, :3, :3].copy())

    for i in range(warm):
        step(i)
    stream = torch.cuda.ExternalStream(odo.stream, device=dev)
    odo.set_profiling(True)
    odo.get_profile(reset=True)
    l0 = odo.launch_count()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    out = None
    for i in range(steps):
        out = step(warm + i)
    e1.record(stream)
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    red_ms, _ = odo.get_profile(reset=True)
    odo.set_profiling(False)
    launches = odo.launch_count() - l0
    G = sets[(warm + steps - 1) % 2][3]
    err_mm = float(np.linalg.norm(out[0].reshape(B, 3) - G, axis=1).max() * 1e3)
    value = world * B * steps / (ms / 1e3)
    # ---- end to end: the same batch from pinned host memory (H2D of every input inside the timed region)
    e2e = None
    if B * BYTES_PER_FRAME_IN * 2 < 4e9:
        hsets = []
        for off in (0, 1):
            d = {k: torch.stack([hframes[(off + 5 * b) % nf][k] for b in range(B)]).pin_memory() for k in keys}
            P = np.stack([frames[(off + 5 * b) % nf]["model_pose"] for b in range(B)])
            hsets.append((d, P, odo.make_frame(d["depth"], d["rgba"], d["mv"], d["mn"], d["mrgba"], P, DEPTH_CUTOFF, MODEL_CUTOFF)))
        for i in range(2):
            d, P, fr = hsets[i % 2]
            odo.track_host(fr, P[:, :3, 3].copy(), P[:, :3, :3].copy())
        barrier()
        t0 = time.perf_counter()
        n_e2e = max(4, steps // 2)
        odo.prefetch_host(hsets[0][2])
        for i in range(n_e2e):
            d, P, fr = hsets[i % 2]
            odo.track_host(fr, P[:, :3, 3].copy(), P[:, :3, :3].copy(), next_frame=hsets[(i + 1) % 2][2] if i + 1 < n_e2e else None)
        barrier()
        e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
        e2e = {"value": world * B * n_e2e / (e2e_ms / 1e3), "unit": "frames/s", "h2d_bytes_per_step": B * (BYTES_PER_FRAME_IN + 64), "d2h_bytes_per_step": B * 48,
               "ms_per_step": e2e_ms / n_e2e, "note": "PCIe-bound: 12.9 MB of host inputs per tracked frame"}
    odo.close()
    px_iter = sum((W >> l) * (H >> l) * ITERS[l] for l in range(LEVELS))
    alg = px_iter * 110.0 * B            # per step and GPU: 48 + 30 + 32 B per pixel-iteration
    achieved = alg * steps / (red_ms * 1e-3) / 1e9 if red_ms > 0 else None
    traffic = None
    tp = ROOT / "profiles" / "batch_kernel_traffic.json"
    if tp.exists():
        try:
            traffic = json.loads(tp.read_text()).get("dram_bytes_per_step")
        except Exception:
            traffic = None
    pair = launches == 3 * steps   # 3 .. 8 sequences per GPU: ONE split launch pair works through them (preparation + cluster kernel + fine-level kernel)
    engine = ("split launch pair of the persistent kernel, sequence after sequence" if pair else "streaming engine")
    return {
        "workload": f"configs[3]: 64 independent synthetic 640x480 sequences, ICP+RGB+SO3, 64 / N per GPU in one batched handle ({engine}), no collective",
        "sequences_total": B * world, "sequences_per_gpu": B, "value": value, "unit": "frames/s", "steps": steps, "warmup": warm, "ms_per_step": ms / steps,
        "scaling": "strong (64 sequences in total)", "gpu_launches_per_step": launches / steps, "e2e": e2e,
        "roofline": {"bound": "hbm", "kernel": "k_gn_persistent pair (one bracket per batched step)" if pair else "kb_phase_a_staged + kb_phase_b (every ICP/RGB reduction launch of a step)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": (achieved / peak) if achieved else None, "traffic": None if pair else traffic, "peak_source": peak_src, "ms_per_step": red_ms / steps,
                     "algorithmic_bytes_per_step": alg, "share_of_step": red_ms / ms if ms else None,
                     "note": "algorithmic bytes = what the reference's icpStep + computeRgbResidual + rgbStep move per pixel-iteration (110 B); the fused kernels "
                             "move less (compacted correspondences, candidate masks), see DESIGN.md"},
        "check": {"max_frame_error_mm": err_mm},
    }


def run_other_configs(args, rank, local_rank, world, odo, dframes, frames, barrier, max_over_ranks):
    """Short measurements of the remaining BASELINE.json configurations (their parity tests are tests/test_gpu_configs.py).

    configs[4]: 256 pose hypotheses per frame scored by the ICP residual reduction, 256 / N per GPU, best pose by one NCCL
                min-allreduce of packed 64-bit keys (slam_b200/relocalise.py);
    configs[2]: 1280x720, 4-level pyramid, ICP only, 10/5/4/4 iterations (rank 0, N = 1 only)."""
    import torch
    from slam_b200 import RGBDOdometry
    from slam_b200.relocalise import score_sharded
    dev = f"cuda:{local_rank}"
    out = {}
    # ---- configs[4]: ONE frame (rank 0's) for all ranks, hypotheses of SURVEY 8(d) (gt o SE3 noise, 5 cm / 3 deg, seed 0xBEEF,
    #      hypothesis 0 unperturbed), score = lastICPError with the inliers >= 1400 guard of lc/Ferns.cpp:275-279 (scaled by the level's
    #      pixel count), sharded over the ranks, winner by one NCCL min-all-reduce of the device-side packed key
    from slam_b200.relocalise import broadcast_frame, connect_peers, perturbed_hypotheses, score_sharded_device, score_sharded_peers
    d = dframes[3]
    fr = frames[3]
    bufs = [d["depth"], d["mv"], d["mn"]]
    meta = torch.from_numpy(np.concatenate([fr["model_pose"].reshape(-1), fr["gt_pose"].reshape(-1)]).astype(np.float64)).to(dev)
    broadcast_frame(bufs + [meta])          # every rank now holds rank 0's frame (outside the timed scoring)
    meta = meta.cpu().numpy()
    model = meta[:16].reshape(4, 4).astype(np.float32)
    gt = meta[16:].reshape(4, 4)
    odo.initICPModel(d["mv"], d["mn"], MODEL_CUTOFF, model)
    odo.initICP(d["depth"], DEPTH_CUTOFF)
    n_hyp = 256
    T, R = perturbed_hypotheses(gt, n_hyp)
    key = torch.full((1,), np.iinfo(np.int64).max, dtype=torch.int64, device=dev)
    ostream = torch.cuda.ExternalStream(odo.stream, device=dev)
    # N > 1: the minimum over the ranks is taken over NVLink peer memory when the ranks can map each other's memory (CUDA IPC), with
    # the NCCL min-all-reduce of the device word timed next to it
    peers = world > 1 and connect_peers(odo, rank, world)
    for level in (0, 2):
        min_inl = 1400 >> (2 * level)

        def timed(fn, reps=20):
            for _ in range(3):
                res = fn()
            barrier()
            t0 = time.perf_counter()
            for _ in range(reps):
                res = fn()
            barrier()
            return max_over_ranks((time.perf_counter() - t0) * 1e3) / reps, res

        nccl_ms, (best, err) = timed(lambda: score_sharded_device(odo, level, model, T, R, key, rank, world, min_inliers=min_inl, stream=ostream))
        ms = nccl_ms
        if peers:
            ms, (pbest, perr) = timed(lambda: score_sharded_peers(odo, level, model, T, R, rank, world, min_inliers=min_inl))
            assert (pbest, np.float32(perr)) == (best, np.float32(err)), f"configs[4] level {level}: peer-memory winner {pbest} / {perr} != NCCL winner {best} / {err}"
        # the winner must not depend on the sharding: every rank scores all hypotheses alone (untimed) and compares
        full_best, full_err, _ = score_sharded(odo, level, model, T, R, 0, 1, min_inliers=min_inl)
        assert (full_best, np.float32(full_err)) == (best, np.float32(err)), f"configs[4] level {level}: sharded winner {best} / {err} != single-GPU winner {full_best} / {full_err}"
        out[f"configs[4] level {level}"] = {"hypotheses": n_hyp, "per_gpu": n_hyp // world, "ms_per_frame": ms, "hypotheses_per_s": n_hyp / (ms * 1e-3),
                                            "best_index": best, "best_error": err, "min_inliers": min_inl, "winner_equals_single_gpu": True,
                                            "perturbation": "sigma_t 5 cm, sigma_r 3 deg, seed 0xBEEF, hypothesis 0 = ground truth",
                                            "collective": ("none" if world == 1 else "minimum over NVLink peer memory: one 16-byte store per peer + poll, in the launch behind the "
                                                           "scoring launch" if peers else "1 NCCL min-all-reduce of one device int64"),
                                            "nccl_ms_per_frame": nccl_ms if world > 1 else None}
    # ---- the depth pre-filter in front of initICP (13x13 bilateral, compute-bound: 169 exp per pixel)
    if world == 1:
        from slam_b200.odometry import load_library
        lib = load_library()
        raw = torch.stack([dframes[(7 * b) % len(dframes)]["depth"] for b in range(64)])
        out_f = torch.zeros_like(raw)
        for nimg in (1, 64):
            for _ in range(3):
                lib.slam_op_depth_bilateral(raw.data_ptr(), H, W, DEPTH_CUTOFF, out_f.data_ptr(), nimg, None)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 50 if nimg == 1 else 10
            e0.record()
            for _ in range(reps):
                lib.slam_op_depth_bilateral(raw.data_ptr(), H, W, DEPTH_CUTOFF, out_f.data_ptr(), nimg, None)
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / reps
            out[f"depth pre-filter 640x480 x{nimg}"] = {"us_per_launch": us, "frames_per_s": nimg / (us * 1e-6), "exp_per_s": nimg * W * H * 169 / (us * 1e-6)}
        try:
            from oracle.cpu_oracle import depth_bilateral
            host = frames[0]["depth"]
            depth_bilateral(host, DEPTH_CUTOFF)
            t0 = time.perf_counter()
            for _ in range(3):
                depth_bilateral(host, DEPTH_CUTOFF)
            out["depth pre-filter 640x480 x1"]["cpu_port_us"] = (time.perf_counter() - t0) / 3 * 1e6
        except Exception:
            pass
        del raw, out_f
    # ---- the fern relocaliser around the ICP path (key-frame database in HBM)
    if world == 1:
        try:
            from slam_b200.ferns import Ferns
            fe = Ferns(500, int(DEPTH_CUTOFF * 1000), 115.0, 319.5, 239.5, 481.20, -480.0, W, H, seed=0x51A7, capacity=1100, device=local_rank)
            nkf = 1024
            for k in range(nkf):
                dk = dframes[k % len(dframes)]
                fe.addFrame(dk["mrgba"], dk["mv"], dk["mn"], frames[k % len(dframes)]["model_pose"], k, -1.0)   # threshold -1: every frame is kept
            dq = dframes[5]
            reps = 50
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(reps):
                fe.encode(dq["mrgba"], dq["mv"], dq["mn"])
            enc_us = (time.perf_counter() - t0) / reps * 1e6
            ms_search = []
            for _ in range(reps):
                fe.search(100000, use_time=True)
                ms_search.append(fe.lastSearchMs())
            cons = []
            t0 = time.perf_counter()
            for _ in range(20):
                fe.findFrame(cons, frames[5]["model_pose"], dq["mv"], dq["mn"], dq["mrgba"], 100000, True)
            ff_ms = (time.perf_counter() - t0) / 20 * 1e3
            us = float(np.median(ms_search)) * 1e3
            out["fern relocaliser, 1024 key frames"] = {"encode_us_incl_readback": enc_us, "search_kernel_us": us, "search_GBps": nkf * 512 / (us * 1e-6) / 1e9,
                                                         "find_frame_ms": ff_ms, "accepted": int(fe.lastClosest >= 0), "icp_count": float(fe.lastMatch.icp_count)}
            fe.close()
        except Exception as e:   # the extra must never take the headline down
            out["fern relocaliser, 1024 key frames"] = {"error": str(e)[:200]}
    # ---- the model-prediction producer (surfel splat + fill-in) that makes the tracker's model maps
    if world == 1:
        try:
            from slam_b200.predict import ModelPredictor
            from slam_b200.synth import Scene, surfels_from_frame
            scene = Scene(seed=0x51A7)
            poses = scene.trajectory(1000, seed=0x51A7)
            mp = ModelPredictor(W, H, 319.5, 239.5, 481.20, -480.0, device=local_rank)
            for nview in (1, 4):
                model = np.concatenate([surfels_from_frame(scene, poses[40 * k], seed=k) for k in range(nview)])
                d_model = torch.from_numpy(model).to(dev)
                dq = dframes[5]
                ts = []
                for it in range(25):
                    mp.predict(poses[5 + it % 3], d_model, len(model), MODEL_CUTOFF, 10.0, 1, 200, dq["depth"], dq["rgba"])
                    ts.append(mp.lastMs())
                a = np.array(ts[5:]) * 1e3
                frags = mp.lastFragments()
                splat_us, resolve_us = float(a[:, 0].mean()), float(a[:, 1].mean())
                # resolve launch, per pixel: z-buffer 8 r + 8 w, winners 8 w, view ray 16 r, winner's surfel 48 r, FillIn textures 36 w, raw rgba 4 r
                out[f"model prediction, {len(model)} surfels"] = {
                    "splat_us": splat_us, "resolve_fill_us": resolve_us, "frames_per_s": 1e6 / (splat_us + resolve_us), "fragments": frags,
                    "fragments_per_s": frags / (splat_us * 1e-6), "surfel_stream_GBps": len(model) * 48 / (splat_us * 1e-6) / 1e9,
                    "resolve_GBps": W * H * 128 / (resolve_us * 1e-6) / 1e9, "covered": float((mp.winners()[1] >= 0).mean())}
                del d_model
            mp.close()
            # the reference's frame loop minus fusion: predict at the last estimated pose, then track from it (one stream, device pointers)
            mp = ModelPredictor(W, H, 319.5 + 0.5, 239.5 + 0.5, 481.20, -480.0, device=local_rank, stream=odo.stream)
            model = np.concatenate([surfels_from_frame(scene, poses[k], conf=25.0, seed=k) for k in (95, 110, 125, 140, 160)])
            d_model = torch.from_numpy(model).to(dev)
            up16 = lambda a: torch.from_numpy(a.view(np.int16).copy()).to(dev)
            seq = [scene.render_frame(poses[k]) for k in range(101, 141)]
            dseq = [(up16(d), torch.from_numpy(c).to(dev)) for d, c in seq]
            tex = mp.fillIn

            def loop(n, pose):
                for i in range(n):
                    dd, dc = dseq[i]
                    mp.predict(pose, d_model, len(model), MODEL_CUTOFF, 10.0, i, 1000, dd, dc)
                    fr = odo.make_frame(dd, dc, tex.vertexTexture, tex.normalTexture, tex.imageTexture, pose, DEPTH_CUTOFF, MODEL_CUTOFF)
                    t, R = odo.track_device(fr, pose[:3, 3].copy(), pose[:3, :3].copy())
                    pose = np.eye(4, dtype=np.float32)
                    pose[:3, :3], pose[:3, 3] = R.reshape(3, 3), t
                return pose
            loop(5, poses[100].copy())
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            last = loop(len(dseq), poses[100].copy())
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            out["closed loop: predict at the estimated pose + track, %d surfels" % len(model)] = {
                "frames_per_s": len(dseq) / dt, "ms_per_frame": dt / len(dseq) * 1e3,
                "final_error_mm": float(np.linalg.norm(last[:3, 3] - poses[140][:3, 3]) * 1e3)}
            mp.close()
        except Exception as e:   # the extra must never take the headline down
            out["model prediction"] = {"error": str(e)[:200]}
    # ---- configs[2]
    if world == 1:
        from slam_b200.synth import Scene
        W2, H2 = 1280, 720
        sc = W2 / 640.0
        scene = Scene(width=W2, height=H2, fx=481.20 * sc, fy=-480.0 * sc, cx=(319.5 + 0.5) * sc - 0.5, cy=(239.5 + 0.5) * sc - 0.5, seed=0x51A7)
        poses = scene.trajectory(1000, seed=0x51A7)
        hd = RGBDOdometry(W2, H2, (319.5 + 0.5) * sc - 0.5, (239.5 + 0.5) * sc - 0.5, 481.20 * sc, -480.0 * sc, num_levels=4, iterations=(10, 5, 4, 4), device=local_rank)
        fl = []
        for i in range(8):
            k = 50 + 100 * i
            depth, rgba = scene.render_frame(poses[k])
            mv, mn, mrgba = scene.render_model(poses[k - 1])
            t = lambda a: torch.from_numpy(a.view(np.int16) if a.dtype == np.uint16 else a).to(dev)
            dd = dict(depth=t(depth), rgba=t(rgba), mv=t(mv), mn=t(mn), mrgba=t(mrgba))
            P = poses[k - 1].astype(np.float32)
            fl.append((dd, P, hd.make_frame(dd["depth"], dd["rgba"], dd["mv"], dd["mn"], dd["mrgba"], P, DEPTH_CUTOFF, MODEL_CUTOFF), poses[k][:3, 3]))
        torch.cuda.synchronize()
        run = lambda i: hd.track_device(fl[i % 8][2], fl[i % 8][1][:3, 3].copy(), fl[i % 8][1][:3, :3].copy(), False, 100.0, True, False, False)
        for i in range(5):
            run(i)
        torch.cuda.synchronize()
        n2 = 100
        t0 = time.perf_counter()
        for i in range(n2):
            last = run(i)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        out["configs[2] 1280x720 4-level ICP-only"] = {"frames_per_s": n2 / dt, "ms_per_frame": dt / n2 * 1e3,
                                                        "last_frame_error_mm": float(np.linalg.norm(last[0] - fl[(n2 - 1) % 8][3]) * 1e3)}
        hd.close()
    return out


def pin_rank_to_cores(local_rank: int, world: int) -> str:
    """One process per GPU: give every rank its own slice of the host cores, so that the per-frame host work of N ranks (launch
    calls, the result poll, the clock sampler) does not migrate or pile up on the same cores.  The slice starts on the NUMA node of
    the rank's GPU when sysfs tells which one that is."""
    try:
        cores = sorted(os.sched_getaffinity(0))
        if world <= 1 or len(cores) < 2 * world:
            return f"{len(cores)} cores, not pinned"
        node_cores = None
        try:
            import torch
            bus = torch.cuda.get_device_properties(local_rank).pci_bus_id if hasattr(torch.cuda.get_device_properties(local_rank), "pci_bus_id") else None
            if bus is not None:
                base = Path("/sys/bus/pci/devices")
                for d in base.iterdir():
                    if d.name.lower().endswith(f"{bus:02x}:00.0"):
                        node = int((d / "numa_node").read_text())
                        if node >= 0:
                            txt = Path(f"/sys/devices/system/node/node{node}/cpulist").read_text().strip()
                            node_cores = []
                            for part in txt.split(","):
                                a, _, b = part.partition("-")
                                node_cores += list(range(int(a), int(b or a) + 1))
                            node_cores = [c for c in node_cores if c in cores]
                        break
        except Exception:
            node_cores = None
        per = len(cores) // world
        mine = cores[local_rank * per:(local_rank + 1) * per]
        if node_cores and len(node_cores) >= per:
            # ranks whose GPUs share a node split that node's cores among them
            share = max(1, len(node_cores) // max(1, world))
            k = (local_rank * share) % max(1, len(node_cores) - share + 1)
            mine = node_cores[k:k + max(share, 2)]
        os.sched_setaffinity(0, set(mine))
        return f"cores {mine[0]}-{mine[-1]} ({len(mine)})"
    except Exception as e:   # pragma: no cover
        return f"not pinned ({e})"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-baselines", action="store_true", help="skip the cpu_baseline / ref_cuda context measurements")
    ap.add_argument("--no-batched", action="store_true", help="skip the batched extra (configs[3])")
    args = ap.parse_args()
    args.warmup = max(3, args.warmup)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    divert_stdout()
    if args.impl == "reference":
        if args.steps > 200:
            args.steps = 200   # bounded sample: ~0.15 s per frame on 8 cores
        run_reference_arm(args, rank, world)
        return
    log(f"[rank {rank}] host cores: {pin_rank_to_cores(local_rank, world)}")
    if world == 1 and args.gpus > 1:
        log(f"--gpus {args.gpus} without torchrun: running a single rank (launch with torch.distributed.run for N > 1)")
    run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
