"""slam_b200 -- B200-native (sm_100a) RGB-D camera tracking hot path.

A from-scratch implementation of ONE path of siw-engineering/slam: the ``RGBDOdometryef``
frame-to-model tracker (ICP + RGB + SO3 pre-alignment), behind the reference's own interface.

  include/slam_odom.h        C ABI of libslam_odom.so (CUDA kernels + host loop)
  include/RGBDOdometryef.hpp header-only C++ class with the reference's method names
  slam_b200.odometry         ctypes mirror of the same interface (tests, bench)
  slam_b200.synth            synthetic ICL-NUIM-shaped data (tests, bench)

There is no CPU fallback: without the CUDA library and a GPU the tracker raises.
"""
from .odometry import RGBDOdometry, OdometryError, load_library, Tap  # noqa: F401

__all__ = ["RGBDOdometry", "OdometryError", "load_library", "Tap"]
