"""stereo_svo_slam_b200 — B200 (sm_100a) implementation of stereo-svo-slam's tracking hot path.

Package layout (only what the path needs):
  csrc/      CUDA kernels, the svo_* device C-ABI and the svo_slam_* host facade (C++)
  capi.py    ctypes binding of include/svo_cuda.h
  slam.py    Python mirror of the reference's wrapper API (StereoSlam, Frame, KeyFrame, ...)
  synth.py   seeded synthetic stereo sequences of the BASELINE configs
  build.py   in-tree nvcc build
"""
from .capi import CameraSettings, SvoError  # noqa: F401
from .slam import StereoSlam, run_many  # noqa: F401
