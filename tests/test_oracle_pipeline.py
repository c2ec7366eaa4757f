"""The CPU oracle's full pipeline (restated StereoSlam::new_image) on the seeded synthetic sequences:
sanity of the restatement itself — it tracks, creates keyframes, and reproduces the reference's quirks."""
import numpy as np

from oracle import oracle as orc
from stereo_svo_slam_b200 import synth


def run(cfg, frames, **over):
    d = synth.settings_dict(cfg)
    d.update(over)
    c = synth.CONFIGS[cfg]
    seq = synth.make_sequence(cfg)
    slam = orc.OracleSlam(orc.CameraSettings(**d), c["width"], c["height"], tracing=True)
    errs = []
    for k in range(frames):
        L, R = seq.render(k)
        slam.new_image(L, R, k / 20.0)
        gt, p = seq.pose(k), slam.pose()
        errs.append((np.abs(p[:3] - gt[:3]).max(), np.abs(p[3:] - gt[3:]).max()))
    return slam, seq, np.array(errs)


def test_oracle_tracks_small_sequence():
    slam, seq, errs = run("S", 20)
    assert errs[:, 0].max() < 0.03 and errs[:, 1].max() < 0.006
    assert slam.n_keyframes() >= 1 and slam.n_kps() > 30
    assert slam.trajectory().shape == (20, 6)
    ev = slam.trace("align_evals").reshape(8, 2)
    assert ev[2, 0] >= 2 and ev[3, 0] >= 2 and ev[:, 0].max() <= 51   # <= 50 evaluations + the initial one (Q3)


def test_oracle_depths_match_ground_truth():
    slam, seq, _ = run("S", 1)
    pose, k2, k3 = slam.keyframe(0)
    z_gt = np.array([seq.depth_at(0, u, v) for (u, v) in k2])
    rel = np.abs(k3[:, 2] - z_gt) / z_gt
    assert np.median(rel) < 0.05   # disparity is integer-quantised


def test_pose_prior_is_one_frame_stale():
    # the prior used for frame k+1 is the filtered pose of frame k-1
    slam, seq, _ = run("S", 6)
    traj = slam.trajectory()
    assert np.allclose(slam.trace("align_pose_in"), traj[3], atol=1e-6)


def test_timestamps_must_increase():
    # dt == 0 poisons the motion filter with inf/NaN
    d = synth.settings_dict("S")
    c = synth.CONFIGS["S"]
    seq = synth.make_sequence("S")
    slam = orc.OracleSlam(orc.CameraSettings(**d), c["width"], c["height"])
    L, R = seq.render(0)
    slam.new_image(L, R, 1.0)
    L, R = seq.render(1)
    slam.new_image(L, R, 1.0)
    assert not np.isfinite(slam.pose()).all()
