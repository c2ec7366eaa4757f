"""The OpenCV-typed drop-in header (include/stereo_slam_b200.hpp) compiles against the C-ABI library and keeps the
reference's signatures. Built with a stand-in for the OpenCV types (no OpenCV C++ SDK in this image)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    exe = str(tmp_path / "facade_check")
    lib = os.path.join(ROOT, "stereo_svo_slam_b200")
    cmd = ["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "tests", "cpp"),
           os.path.join(ROOT, "tests", "cpp", "facade_check.cpp"), "-o", exe, "-L", lib, "-lstereosvo_b200", f"-Wl,-rpath,{lib}"]
    subprocess.check_call(cmd)
    return exe


def test_facade_header_compiles_and_fails_loudly_without_gpu(tmp_path):
    from stereo_svo_slam_b200 import capi
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True)
    if capi.device_count() == 0:
        assert r.returncode == 10 and "no CUDA device" in r.stdout, (r.returncode, r.stdout, r.stderr)
    else:
        assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)


@pytest.mark.gpu
def test_facade_header_tracks_on_gpu(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "tracked" in r.stdout, (r.returncode, r.stdout, r.stderr)


def test_signatures_match_reference_header():
    src = open(os.path.join(ROOT, "include", "stereo_slam_b200.hpp")).read()
    # src/include/stereo_slam.hpp:35-64
    for sig in ["StereoSlam(const CameraSettings &camera_settings)",
                "void new_image(const cv::Mat &left, const cv::Mat &right, const float time_stamp)",
                "void get_keyframe(KeyFrame &keyframe)", "void get_keyframes(std::vector<KeyFrame> &keyframes)",
                "bool get_frame(Frame &frame)", "void get_trajectory(std::vector<Pose> &trajectory)",
                "Pose update_pose(const Pose &pose, const cv::Vec6f &speed, const cv::Vec6f &pose_variance, const cv::Vec6f &speed_variance,"]:
        assert sig in src, sig
