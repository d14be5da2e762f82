"""Build the native pieces in-tree with nvcc / gcc (no JIT cache: the .so files travel with the repo).

  slam_b200/libslam_odom.so    the product: CUDA kernels + C ABI (include/slam_odom.h), sm_100a
  slam_b200/libslam_synth.so   synthetic ICL-NUIM-shaped scene ray-caster (test/bench data only)
  oracle/liboracle.so          CPU restatement of the reference math (test infrastructure)
  oracle/_ref/libslam_ref.so   the reference's own src/odom kernels compiled for sm_100a where they
                               lie under /root/reference + our replay harness (test infrastructure;
                               only built when /root/reference is present)

Numeric flags of the product are the reference's (src/CMakeLists.txt:115-116): the parity contract
(bit-exact masks and pyramids) depends on them.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
CSRC = ROOT / "slam_b200" / "csrc"
REF = Path(os.environ.get("SLAM_REFERENCE_DIR", "/root/reference"))

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
REF_NUMERIC_FLAGS = ["--ftz=true", "--prec-div=false", "--prec-sqrt=false"]
NVCC_COMMON = ARCH + ["-std=c++17", "-O3", "-lineinfo"] + REF_NUMERIC_FLAGS + ["-Xcompiler", "-fPIC"]
# predict.cu restates GLSL shaders, not the reference's CUDA: IEEE division / square root, no FMA contraction, no flush to
# zero, so that the CPU restatement (oracle/predict_oracle.c) reproduces it bit for bit
IEEE_NUMERIC_FLAGS = ["--ftz=false", "--prec-div=true", "--prec-sqrt=true", "--fmad=false"]
NVCC_IEEE = ARCH + ["-std=c++17", "-O3", "-lineinfo"] + IEEE_NUMERIC_FLAGS + ["-Xcompiler", "-fPIC"]
PER_FILE_FLAGS = {"predict.cu": NVCC_IEEE}


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found")


def _run(cmd, **kw):
    print("+", " ".join(str(c) for c in cmd), flush=True)
    subprocess.run([str(c) for c in cmd], check=True, **kw)


def _stale(target: Path, sources) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(s).stat().st_mtime > t for s in sources)


def build_product(force: bool = False) -> Path:
    out = ROOT / "slam_b200" / "libslam_odom.so"
    cus = [CSRC / n for n in ("odom_api.cu", "gn_kernel.cu", "batch_engine.cu", "reduce_kernels.cu", "prep_kernels.cu", "ferns.cu", "predict.cu")]
    deps = cus + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.hpp")) + [ROOT / "include" / n for n in ("slam_odom.h", "slam_ferns.h", "slam_predict.h")]
    if not force and not _stale(out, deps):
        return out
    objdir = ROOT / "build" / "product"
    objdir.mkdir(parents=True, exist_ok=True)
    objs = []
    procs = []
    for cu in cus:
        obj = objdir / (cu.stem + ".o")
        objs.append(obj)
        cmd = [_nvcc()] + PER_FILE_FLAGS.get(cu.name, NVCC_COMMON) + ["-I", ROOT / "include", "-c", cu, "-o", obj]
        print("+", " ".join(str(c) for c in cmd), flush=True)
        procs.append(subprocess.Popen([str(c) for c in cmd]))
    for p in procs:
        if p.wait() != 0:
            raise RuntimeError("nvcc failed")
    _run([_nvcc()] + ARCH + ["-shared", "-o", out] + objs)
    return out


def build_synth(force: bool = False) -> Path:
    out = ROOT / "slam_b200" / "libslam_synth.so"
    src = ROOT / "slam_b200" / "synth" / "synth.c"
    if force or _stale(out, [src]):
        _run(["gcc", "-O3", "-march=x86-64-v2", "-fopenmp", "-fPIC", "-shared", "-o", out, src, "-lm"])
    return out


def build_hostmath(force: bool = False) -> Path:
    out = ROOT / "slam_b200" / "libslam_hostmath.so"
    src = CSRC / "hostmath_capi.cpp"
    if force or _stale(out, [src, CSRC / "small_math.hpp"]):
        _run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", out, src])
    return out


def build_oracle(force: bool = False) -> Path:
    out = ROOT / "oracle" / "liboracle.so"
    srcs = [ROOT / "oracle" / "odom_oracle.c", ROOT / "oracle" / "depth_filter_oracle.c", ROOT / "oracle" / "predict_oracle.c"]
    if force or _stale(out, srcs):
        _run(["gcc", "-O3", "-march=x86-64-v2", "-ffp-contract=off", "-fopenmp", "-fPIC", "-shared", "-o", out] + srcs + ["-lm"])
    return out


def build_reference(force: bool = False) -> Path | None:
    """oracle/_ref/libslam_ref.so: reference kernels (unmodified, compiled from /root/reference) + harness."""
    out = ROOT / "oracle" / "_ref" / "libslam_ref.so"
    refsrc = REF / "src"
    if not (refsrc / "odom" / "reduce.cu").exists():
        return out if out.exists() else None
    shim = ROOT / "oracle" / "ref_shim.cuh"
    harness = ROOT / "oracle" / "ref_harness.cu"
    deps = [shim, harness, CSRC / "small_math.hpp", ROOT / "include" / "slam_odom.h", refsrc / "odom" / "reduce.cu", refsrc / "odom" / "utils.cu"]
    if not force and not _stale(out, deps):
        return out
    out.parent.mkdir(parents=True, exist_ok=True)
    objdir = ROOT / "build" / "ref"
    objdir.mkdir(parents=True, exist_ok=True)
    # the reference's own nvcc flags (src/CMakeLists.txt:115-116) + sm_100a
    flags = ARCH + REF_NUMERIC_FLAGS + ["-D_FORCE_INLINES", "-O3", "-lineinfo", "-Xcompiler", "-fPIC", "-I", refsrc]
    procs = []
    objs = []
    for name, path, extra in (
        ("reduce", refsrc / "odom" / "reduce.cu", ["-include", shim]),
        ("utils", refsrc / "odom" / "utils.cu", ["-include", shim]),
        ("harness", harness, ["-include", shim, "-I", CSRC, "-I", ROOT / "include", "-std=c++17"]),
    ):
        obj = objdir / (name + ".o")
        objs.append(obj)
        cmd = [_nvcc()] + flags + extra + ["-c", path, "-o", obj]
        print("+", " ".join(str(c) for c in cmd), flush=True)
        procs.append(subprocess.Popen([str(c) for c in cmd]))
    dm = objdir / "device_memory.o"
    objs.append(dm)
    _run(["g++", "-O2", "-fPIC", "-I", "/usr/local/cuda/include", "-I", refsrc, "-c", refsrc / "cuda" / "containers" / "device_memory.cpp", "-o", dm])
    for p in procs:
        if p.wait() != 0:
            raise RuntimeError("nvcc failed (reference)")
    _run([_nvcc()] + ARCH + ["-shared", "-o", out] + objs)
    return out


def build_all(force: bool = False) -> None:
    build_product(force)
    build_synth(force)
    build_hostmath(force)
    build_oracle(force)
    build_reference(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
