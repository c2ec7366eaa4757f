"""Golden vectors from the REFERENCE ITSELF (oracle/_ref = the unmodified /root/reference/src/lib sources compiled against
oracle/cvshim, see oracle/Makefile).  Run in the build container, where /root/reference exists:

    python tests/golden/make_ref_golden.py

writes tests/golden/ref_vectors.npz: for every case the reference's trajectory, its final frame and all its keyframes
(2-D / 3-D keypoints, level, type, keyframe_id, keypoint_index, flags, inlier / outlier counters, score, depth-filter
state).  tests/test_ref_pin.py checks the CPU oracle against these vectors everywhere, and against a live run of
oracle/_ref wherever that library is present.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import oracle as orc  # noqa: E402
from stereo_svo_slam_b200 import synth  # noqa: E402

# (name, synth config, frames, settings overrides, IMU-style update_pose calls between frames)
CASES = [
    ("S", "S", 120, {}, False),
    ("SF", "SF", 100, {}, False),          # three keyframes: merge with old keypoints, several origin keyframes
    ("SF_imu", "SF", 80, {}, True),        # StereoSlam::update_pose called between frames with dt > 0
    ("S_dist", "S", 30, dict(k1=-0.12, k2=0.05, p1=0.001, p2=-0.0015, k3=0.01), False),   # full projectPoints model
    ("S_blender_grid", "S", 30, dict(grid_width=25, grid_height=16, max_pyramid_levels=5), False),  # odd grid, 5 levels
    ("C3", "C3", 64, {}, False),           # BASELINE configs[2]; keyframe #2 at frame 46
    ("C4", "C4", 12, {}, False),           # BASELINE configs[3]: 1280x720, 5 levels, 16x14 grid (3 000 keypoints per frame)
    # the algorithm settings of the shipped src/app/EuRoC.yaml and Blender.yaml (BASELINE configs[0]/[1]; their videos are missing)
    ("C3_euroc_yaml", "C3", 24, dict(grid_width=54, grid_height=48, search_x=60, search_y=6, max_pyramid_levels=6,
                                     min_pyramid_level_pose_estimation=2), False),
    ("C3_blender_yaml", "C3", 24, dict(grid_width=75, grid_height=48, search_x=50, search_y=6, max_pyramid_levels=5,
                                       min_pyramid_level_pose_estimation=2), False),
]
FIELDS = ("kps2d", "kps3d", "score", "kf_state", "kf_cov") + orc.INFO_COLS[:-1]   # colour is rand(): not compared


def imu_call(seq, k):
    """A deterministic IMU-like sample for the gap after frame k: the pose half a frame ahead and the velocity of the synthetic
    trajectory, both with noise, the app's kind of variances, dt = 0.025 s (pose, speed, variances, dt)."""
    rng = np.random.default_rng(9000 + k)
    p0, p1 = seq.pose(k), seq.pose(k + 1)
    pose = (0.5 * (p0 + p1) + rng.standard_normal(6) * 2e-3).astype(np.float32)
    speed = ((p1 - p0) / 0.05 + rng.standard_normal(6) * 0.05).astype(np.float32)
    return pose, speed, np.full(6, 0.5, np.float32), np.full(6, 2.0, np.float32), 0.025


def run_case(slam_cls, cfg, frames, over, imu, on_frame=None):
    c = synth.CONFIGS[cfg]
    d = synth.settings_dict(cfg)
    d.update(over)
    seq = synth.make_sequence(cfg)
    slam = slam_cls(orc.CameraSettings(**d), c["width"], c["height"])
    imu_out = []
    for k in range(frames):
        L, R = seq.render(k)
        slam.new_image(L, R, k / 20.0)
        if on_frame:
            on_frame(k, slam)
        if imu and k % 3 == 2:
            imu_out.append(slam.update_pose(*imu_call(seq, k)))
    return slam, np.array(imu_out, np.float32).reshape(-1, 6)


def collect(slam, imu_out):
    out = {"trajectory": slam.trajectory(), "imu_out": imu_out, "n_keyframes": np.array([slam.n_keyframes()])}
    f = slam.frame()
    for key in FIELDS + ("pose",):
        out["frame_" + key] = f[key]
    for k in range(slam.n_keyframes()):
        kf = slam.keyframe_full(k)
        for key in FIELDS + ("pose",):
            out[f"kf{k}_{key}"] = kf[key]
    return out


def main():
    assert orc.have_ref(), "oracle/_ref is not built (needs /root/reference)"
    vec = {}
    for name, cfg, frames, over, imu in CASES:
        slam, imu_out = run_case(orc.RefSlam, cfg, frames, over, imu)
        for k, v in collect(slam, imu_out).items():
            vec[f"{name}/{k}"] = v
        print(name, "frames", frames, "keyframes", slam.n_keyframes(), "kps", slam.n_kps())
        slam.close()
    path = os.path.join(ROOT, "tests", "golden", "ref_vectors.npz")
    np.savez_compressed(path, **vec)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
