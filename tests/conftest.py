import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
os.environ.setdefault("OMP_NUM_THREADS", str(min(8, os.cpu_count() or 1)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def built():
    """Native libraries are built in-tree (idempotent); the GPU box only uses the prebuilt files."""
    from slam_b200 import build
    build.build_synth()
    if not (ROOT / "slam_b200" / "libslam_odom.so").exists():
        build.build_product()
    return True


@pytest.fixture(scope="session")
def icl_sequence(built):
    """Scene + 40-pose slice of the 1000-frame orbit at 640x480 (ICL-NUIM intrinsics)."""
    from tests.support import make_scene
    scene, intr = make_scene(640, 480)
    poses = scene.trajectory(1000)
    return scene, intr, poses


@pytest.fixture(scope="session")
def ref_lib(built):
    from oracle import ref_cuda
    if not ref_cuda.available():
        pytest.skip("oracle/_ref/libslam_ref.so not built (needs /root/reference at build time)")
    return ref_cuda.load()
