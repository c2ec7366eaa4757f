"""The pin of the CPU oracle against the REFERENCE ITSELF.

oracle/_ref/libstereosvo_ref.so is built from the unmodified reference sources (/root/reference/src/lib/*.cpp) against
oracle/cvshim (oracle/Makefile).  tests/golden/ref_vectors.npz holds its outputs (tests/golden/make_ref_golden.py):
trajectories, final frames and every keyframe of nine runs — small, fast-motion (three keyframes), fast-motion with IMU
updates, lens distortion, an odd grid with five pyramid levels, BASELINE configs[2] across keyframe #2, BASELINE configs[3]
(1280x720, five levels, 3 000 keypoints) and the algorithm settings of the shipped EuRoC.yaml and Blender.yaml.

 * everywhere: the oracle's restatement (oracle/svo_oracle.cpp) must reproduce those vectors BIT FOR BIT — positions,
   3-D points, levels, types, origin keyframe / index, flags, vote counters, scores, depth-filter states, IMU outputs;
 * where oracle/_ref exists (always in the build container; it also travels to the GPU box): a live run of the reference,
   frame by frame against the oracle, the vectors re-derived, and the reference's stage functions against the oracle's.
"""
import os

import numpy as np
import pytest

from oracle import oracle as orc
from stereo_svo_slam_b200 import synth
from tests.golden import make_ref_golden as mg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VEC = np.load(os.path.join(ROOT, "tests", "golden", "ref_vectors.npz"))
needs_ref = pytest.mark.skipif(not orc.have_ref(), reason="oracle/_ref is not built and /root/reference is absent")


def same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and a.dtype == b.dtype and a.tobytes() == b.tobytes()


def check_against_vectors(name, got):
    keys = [k for k in VEC.files if k.startswith(name + "/")]
    assert keys and len(keys) == len(got), (len(keys), len(got))
    bad = [k for k in keys if not same(VEC[k], got[k[len(name) + 1:]])]
    assert not bad, bad


@pytest.mark.parametrize("case", [c[0] for c in mg.CASES])
def test_oracle_reproduces_the_reference_bit_for_bit(case):
    name, cfg, frames, over, imu = next(c for c in mg.CASES if c[0] == case)
    slam, imu_out = mg.run_case(lambda cs, w, h: orc.OracleSlam(cs, w, h, tracing=False), cfg, frames, over, imu)
    check_against_vectors(name, mg.collect(slam, imu_out))


def test_vectors_cover_the_keyframe_logic():
    # what the vectors exercise: old + new keypoints in one keyframe, keypoints of three origin keyframes in one frame,
    # every flag, both vote counters
    assert VEC["SF/n_keyframes"][0] == 3 and VEC["SF_imu/n_keyframes"][0] >= 2 and VEC["C3/n_keyframes"][0] == 2
    assert set(np.unique(VEC["SF/frame_keyframe_id"])) == {0, 1, 2}
    kf2 = VEC["SF/kf2_keyframe_id"]
    assert (kf2 < 2).any() and (kf2 == 2).any()                      # merged: survivors first, then the new keypoints
    assert (np.diff(np.nonzero(kf2 == 2)[0]) == 1).all()
    flags = np.concatenate([VEC[f"SF/kf{k}_flags"] for k in range(3)])
    assert (flags & 2).any() and (flags & 4).any()
    assert VEC["SF/frame_inlier_count"].max() > 20 and VEC["SF/frame_outlier_count"].max() > 0
    assert VEC["C3/trajectory"].shape == (64, 6) and VEC["SF_imu/imu_out"].shape[0] >= 10


@needs_ref
def test_live_reference_frame_by_frame():
    """Every frame of the three-keyframe run: pose, keypoints and all bookkeeping identical; also the vectors re-derived."""
    o_holder = {}

    def oracle_factory(cs, w, h):
        return orc.OracleSlam(cs, w, h, tracing=False)

    name, cfg, frames, over, imu = next(c for c in mg.CASES if c[0] == "SF")
    c, d = synth.CONFIGS[cfg], synth.settings_dict(cfg)
    o = oracle_factory(orc.CameraSettings(**d), c["width"], c["height"])
    seq = synth.make_sequence(cfg)

    def on_frame(k, ref):
        L, R = seq.render(k)
        o.new_image(L, R, k / 20.0)
        fr, fo = ref.frame(), o.frame()
        assert fr["id"] == fo["id"] == k and fr["ts"] == fo["ts"]
        assert ref.n_keyframes() == o.n_keyframes()
        for key in mg.FIELDS + ("pose",):
            assert same(fr[key], fo[key]), (k, key)
        o_holder["n"] = k + 1

    ref, imu_out = mg.run_case(orc.RefSlam, cfg, frames, over, imu, on_frame)
    assert o_holder["n"] == frames
    check_against_vectors(name, mg.collect(ref, imu_out))
    for k in range(ref.n_keyframes()):
        a, b = ref.keyframe_full(k), o.keyframe_full(k)
        for key in mg.FIELDS + ("pose",):
            assert same(a[key], b[key]), (k, key)
    ref.close()


@needs_ref
def test_reference_stage_functions(fixture_images):
    d = synth.settings_dict("S")
    d.update(k1=-0.1, k2=0.03, p1=0.002, p2=-0.001, k3=0.004)
    cs = orc.CameraSettings(**d)
    ref = orc.RefSlam(cs, 320, 240)
    # exponential_map.hpp:12-37 — the known answer of src/test/test_exponential_map.cpp:37-47 and random twists
    kat = ref.expmap([0.1, 0.2, 0.3, 0.4, 0.5, 0.6])
    assert np.abs(kat[:3] - [0.12187591059875308, 0.173369312443241, 0.30760829923146377]).max() < 1e-6
    assert same(kat[3:], np.array([0.4, 0.5, 0.6], np.float32))
    rng = np.random.default_rng(5)
    for s in (1e-3, 1e-2, 0.3):
        for _ in range(300):
            tw = (rng.standard_normal(6) * s).astype(np.float32)
            assert same(ref.expmap(tw), orc.expmap(tw))
    # transform_keypoints.cpp:11-47 (full distortion model)
    pts = (rng.uniform(-2, 2, (500, 3)) + [0, 0, 5]).astype(np.float32)
    pose = np.array([0.1, -0.2, 0.05, 0.02, -0.03, 0.015], np.float32)
    assert same(ref.project(pose, pts), orc.project(cs, pose, pts))
    # corner_detector.cpp:13-79 on the reference's own fixture images (src/test/left.png, testimage0.png), three grids
    for img in (fixture_images["left"], fixture_images["testimage0"]):
        for gw, gh in ((30, 24), (75, 48), (16, 14)):
            a, b = ref.detect_keypoints(img, gw, gh), orc.detect_keypoints(img, gw, gh)
            assert all(same(x, y) for x, y in zip(a, b))
    ref.close()


@needs_ref
def test_reference_keyframe_pyramids(fixture_images):
    """stereo_slam.cpp:135-139: the reference's own half-sample and LK pyramids (images + Scharr planes) of a keyframe."""
    L, R = fixture_images["left"], fixture_images["right"]
    h, w = L.shape
    cs = orc.CameraSettings(**dict(synth.settings_dict("C3")))
    ref = orc.RefSlam(cs, w, h)
    ref.new_image(L, R, 0.0)
    lv = L
    for i in range(cs.max_pyramid_levels):
        assert same(ref.keyframe_image(0, i), lv)
        lv = orc.half_sample(lv)
    assert same(ref.keyframe_image(1, 0), R) and ref.keyframe_image(1, 1) is None
    lv = L
    for i in range(3):
        assert same(ref.keyframe_image(2, i), lv)
        assert same(ref.keyframe_image(3, i), orc.scharr(lv))
        lv = orc.pyr_down(lv)
    ref.close()
