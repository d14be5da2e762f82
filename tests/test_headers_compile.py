"""The public headers are what a maintainer of the reference would compile against: they must be valid C99 (the three C-ABI headers)
and C++14 (the RGBDOdometryef shim), and a program written against the reference's method names must link with libslam_odom.so.
Without a GPU that program has to fail loudly in the constructor (no CPU fallback); with one it tracks a tiny blank frame."""
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent

C_SOURCE = r"""
#include "slam_odom.h"
#include "slam_ferns.h"
#include "slam_predict.h"
int main(void)
{
    slam_odom_params p = {0};
    slam_predict_params q = {0};
    slam_ferns_params f = {0};
    slam_predict_textures t = {0};
    (void)p; (void)q; (void)f; (void)t;
    return SLAM_OK;
}
"""

CPP_SOURCE = r"""
#include <cstdio>
#include <cuda_runtime_api.h>
#include "RGBDOdometryef.hpp"
#include "ModelPrediction.hpp"
int main()
{
    try
    {
        // the reference's constructor arguments (RGBDOdometryef.h:32-36)
        RGBDOdometryef odom(128, 96, 63.5f, 47.5f, 96.f, -96.f);
        unsigned short * depth = nullptr;
        unsigned char * rgba = nullptr;
        float * verts = nullptr, * norms = nullptr;
        cudaMalloc((void **)&depth, 128 * 96 * 2);
        cudaMalloc((void **)&rgba, 128 * 96 * 4);
        cudaMalloc((void **)&verts, 128 * 96 * 16);
        cudaMalloc((void **)&norms, 128 * 96 * 16);
        cudaMemset(depth, 0, 128 * 96 * 2);
        cudaMemset(rgba, 0, 128 * 96 * 4);
        cudaMemset(verts, 0, 128 * 96 * 16);
        cudaMemset(norms, 0, 128 * 96 * 16);
        const float pose[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
        odom.initFirstRGB(rgba);
        odom.initICPModel(verts, norms, 20.f, pose);
        odom.initRGBModel(rgba);
        odom.initICP(depth, 3.f);
        odom.initRGB(rgba);
        float trans[3] = {0, 0, 0}, rot[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        odom.getIncrementalTransformation(trans, rot, false, 10.f, true, false, true);
        // nothing to align: the pose must come back unchanged
        std::printf("tracked %g %g %g count %g\n", trans[0], trans[1], trans[2], odom.lastICPCount);
        // the producer and the relocaliser with the reference's class and method names: an empty map predicts nothing, the fill-in
        // passes then take everything from the (empty) raw frame, and a relocaliser without key frames finds nothing
        slam_b200::IndexMap indexMap(128, 96, 63.5f, 47.5f, 96.f, -96.f);
        slam_b200::FillIn fillIn(indexMap);
        indexMap.combinedPredict(pose, std::make_pair((const float *)nullptr, 0), 20.f, 10.f, 1, 1, 200, slam_b200::IndexMap::ACTIVE);
        fillIn.vertex(indexMap.vertexTex(), depth, false);
        fillIn.normal(indexMap.normalTex(), depth, false);
        fillIn.image(indexMap.imageTex(), rgba, false);
        odom.initICPModel(fillIn.vertexTexture(), fillIn.normalTexture(), 20.f, pose);
        slam_b200::Ferns ferns(500, 3000, 115.f, 63.5f, 47.5f, 96.f, -96.f, 128, 96, 7u);
        std::vector<slam_b200::Ferns::SurfaceConstraint> constraints;
        float est[16];
        ferns.findFrame(constraints, pose, fillIn.vertexTexture(), fillIn.normalTexture(), fillIn.imageTexture(), 1000, true, est);
        std::printf("frames %d closest %d constraints %d\n", ferns.numFrames(), ferns.lastClosest, (int)constraints.size());
        if(ferns.numFrames() != 0 || ferns.lastClosest != -1 || !constraints.empty()) return 4;
        return (std::fabs(trans[0]) < 1e-6f && std::fabs(rot[0] - 1.f) < 1e-6f) ? 0 : 3;
    }
    catch(const std::exception & e)
    {
        std::printf("exception: %s\n", e.what());
        return 2;
    }
}
"""


def test_c_headers_are_valid_c99(tmp_path):
    src = tmp_path / "abi.c"
    src.write_text(C_SOURCE)
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-fsyntax-only", "-I", str(ROOT / "include"), str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def _build_cpp(tmp_path):
    src = tmp_path / "caller.cpp"
    src.write_text(CPP_SOURCE)
    exe = tmp_path / "caller"
    lib = ROOT / "slam_b200"
    cmd = ["g++", "-std=c++14", "-Wall", "-I", str(ROOT / "include"), "-I", "/usr/local/cuda/include", str(src), "-o", str(exe), "-L", str(lib), "-lslam_odom",
           "-L", "/usr/local/cuda/lib64", "-lcudart", f"-Wl,-rpath,{lib}", "-Wl,-rpath,/usr/local/cuda/lib64"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_cpp_shim_links_and_fails_loudly_without_a_gpu(built, tmp_path):
    import torch
    exe = _build_cpp(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("a GPU is present (see the gpu test)")
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 2 and "exception" in r.stdout and "CUDA" in r.stdout, (r.returncode, r.stdout, r.stderr)


@pytest.mark.gpu
def test_cpp_caller_with_the_reference_method_names_runs(built, tmp_path):
    exe = _build_cpp(tmp_path)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "tracked" in r.stdout, (r.returncode, r.stdout, r.stderr)
