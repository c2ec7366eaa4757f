"""Host stages of the library (svo_host_*, svo_motion_filter_*: the DepthCalculator / KeyFrameManager / StereoSlam bookkeeping
that stays on the CPU, the same functions the facade calls) against the oracle — which tests/test_ref_pin.py pins bit for bit
to the reference — and, for the motion filter, against the reference itself.  No GPU needed.

Reference: select_best_keypoints depth_calculator.cpp:37-65, find_bad_keypoints :67-86, merge_keypoints :88-130 (called with
swapped grid arguments :179-180), keyframe_needed keyframe_manager.cpp:47-74, update_pose stereo_slam.cpp:296-359."""
import numpy as np
import pytest

from oracle import oracle as orc
from stereo_svo_slam_b200 import capi, synth


def rand_points(rng, n, w, h, on_lines=0.3, grid=(30, 24)):
    """points inside / outside the image, many exactly on cell borders and image edges (the strict inequalities matter)"""
    p = np.stack([rng.uniform(-20, w + 20, n), rng.uniform(-20, h + 20, n)], 1).astype(np.float32)
    m = rng.random(n) < on_lines
    p[m, 0] = (rng.integers(0, w // grid[1] + 2, m.sum()) * grid[1]).astype(np.float32)   # multiples of either grid size
    m = rng.random(n) < on_lines
    p[m, 1] = (rng.integers(0, h // grid[0] + 2, m.sum()) * grid[0]).astype(np.float32)
    m = rng.random(n) < 0.1
    p[m] = np.floor(p[m])
    return p


@pytest.mark.parametrize("seed", range(6))
def test_select_best_keypoints(seed):
    rng = np.random.default_rng(seed)
    n0 = int(rng.integers(1, 400))
    levels = []
    for lv in range(int(rng.integers(1, 4))):
        n = n0 if lv == 0 else int(rng.integers(max(1, n0 - 40), n0 + 40))   # coarser lists may be shorter or longer (Q7)
        xy = rng.integers(0, 700, (n, 2)).astype(np.float32)
        score = rng.integers(0, 40, n).astype(np.float32)                      # many equal scores: `>` vs `>=`
        levels.append((xy, score, rng.integers(0, 2, n).astype(np.int32)))
    if any(len(l[1]) < n0 for l in levels[1:]):
        # the reference reads past the end of a shorter coarse list (undefined behaviour); both sides skip those entries
        pass
    a, b = capi.host_select_best_keypoints(levels), orc.select_best(levels)
    for x, y in zip(a, b):
        assert x.tobytes() == y.tobytes()


@pytest.mark.parametrize("seed", range(4))
def test_find_bad_keypoints_and_keyframe_needed(seed):
    rng = np.random.default_rng(100 + seed)
    w, h = 752, 480
    p = rand_points(rng, 600, w, h)
    p[:8] = [[0, 0], [w, h], [w, 0], [0, h], [-0.0, 5], [w + 0.001, 5], [5, h + 0.001], [w - 0.5, h - 0.5]]
    flags = rng.integers(0, 8, len(p)).astype(np.uint8)
    assert (capi.host_find_bad_keypoints(w, h, p, flags) == orc.find_bad(w, h, p, flags)).all()
    for n in (0, 100, 329, 330, 331, 600):      # 0.66 * 25 * 20 = 330: around the threshold
        q = p[:n].copy()
        q[:, 0] = np.clip(q[:, 0], 1, w - 1)
        q[:, 1] = np.clip(q[:, 1], 1, h - 1)
        f = np.zeros(n, np.uint8)
        assert capi.host_keyframe_needed(w, h, 30, 24, q, f) == orc.keyframe_needed(w, h, 30, 24, q, f)
    assert capi.host_keyframe_needed(w, h, 30, 24, p, flags) == orc.keyframe_needed(w, h, 30, 24, p, flags)


@pytest.mark.parametrize("w,h,gw,gh", [(752, 480, 30, 24), (752, 480, 75, 48), (1280, 720, 16, 14), (320, 240, 32, 24), (101, 77, 10, 7)])
def test_merge_keypoints(w, h, gw, gh):
    rng = np.random.default_rng(w + gw)
    for trial in range(4):
        n_old = int(rng.integers(0, 2 * (w // gw) * (h // gh)))
        old = rand_points(rng, n_old, w, h, grid=(gh, gw))
        # new keypoints: one per detection cell, integer valued, like CornerDetector's output after select_best
        cx, cy = np.meshgrid(np.arange(w // gw), np.arange(h // gh))
        new = np.stack([cx.ravel() * gw + rng.integers(0, gw, cx.size), cy.ravel() * gh + rng.integers(0, gh, cx.size)], 1).astype(np.float32)
        if trial == 3:
            old = old[:0]                                                        # first keyframe: every cell is free
        a, b = capi.host_merge_keypoints(w, h, gw, gh, old, new), orc.merge(w, h, gw, gh, old, new)
        assert a.tolist() == b.tolist()
        if trial == 3:
            assert len(a) > 0.5 * len(new)


def test_merge_grid_arguments_are_swapped():
    # the cells walked are grid_height wide and grid_width tall.  An old keypoint at (35, 10) lies in the swapped
    # cell x in (24, 48), y in (0, 30): a new keypoint at (40, 20) is NOT added, although it lies in another 30x24 cell row.
    old = np.array([[35, 10]], np.float32)
    new = np.array([[40, 20], [10, 20]], np.float32)
    a = capi.host_merge_keypoints(60, 48, 30, 24, old, new)
    assert a.tolist() == orc.merge(60, 48, 30, 24, old, new).tolist()
    assert 0 not in a.tolist() and 1 in a.tolist()


def imu_samples(n, seed=3):
    rng = np.random.default_rng(seed)
    for k in range(n):
        yield ((rng.standard_normal(6) * 0.05 * (k + 1)).astype(np.float32), (rng.standard_normal(6) * 0.2).astype(np.float32),
               rng.uniform(0.05, 0.5, 6).astype(np.float32), rng.uniform(0.5, 2.0, 6).astype(np.float32), float(rng.choice([0.0, 0.05, 0.011, 1.0])))


def test_motion_filter_matches_oracle_and_reference():
    """StereoSlam::update_pose (stereo_slam.cpp:296-359) through the library's host code, the oracle and the reference itself.
    The library solves the 12x12 gain system by Gaussian elimination in double, OpenCV by a float Jacobi SVD: float noise."""
    cs = orc.CameraSettings(**synth.settings_dict("S"))
    o = orc.OracleSlam(cs, 320, 240, tracing=False)
    r = orc.RefSlam(cs, 320, 240) if orc.have_ref() else None
    g = capi.MotionFilter()
    worst = worst_ref = 0.0
    for pose, speed, pv, sv, dt in imu_samples(40):
        a, pre = g.update(pose, speed, pv, sv, dt)
        b = o.update_pose(pose, speed, pv, sv, dt)
        scale = max(1.0, np.abs(b).max())
        worst = max(worst, np.abs(a - b).max() / scale)
        if r is not None:
            c = r.update_pose(pose, speed, pv, sv, dt)
            assert c.tobytes() == b.tobytes()          # oracle == reference, bit for bit
            worst_ref = max(worst_ref, np.abs(a - c).max() / scale)
    assert worst <= 2e-6 and worst_ref <= 2e-6, (worst, worst_ref)
    if r is not None:
        r.close()
