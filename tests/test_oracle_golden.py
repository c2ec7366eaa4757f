"""Pins the CPU oracle's restatement of the OpenCV-owned primitives (oracle/ocv_prims.hpp) against golden
vectors produced by the real library (cv2 4.13.0, tests/golden/make_golden.py), and the reference's own
known-answer for exponential_map (src/test/test_exponential_map.cpp:37-47)."""
import hashlib

import numpy as np

from oracle import oracle as orc


def sha(a):
    return np.frombuffer(hashlib.sha1(np.ascontiguousarray(a).tobytes()).digest(), np.uint8)


def test_fast_bit_exact(fixture_images, cv2_vectors):
    for nm in ("left", "testimage0"):
        got = orc.fast(fixture_images[nm], 6)
        want = cv2_vectors[f"fast_{nm}"]
        assert got.shape == want.shape
        assert (got == want.astype(np.int32)).all()  # same list, raster order, same scores


def test_sobel_bit_exact(fixture_images, cv2_vectors):
    e = orc.sobel_x(fixture_images["left"])
    assert (sha(e) == cv2_vectors["sobel_left_sha"]).all()
    assert (e[100:104] == cv2_vectors["sobel_left_rows"]).all()


def test_lk_pyramid_bit_exact(fixture_images, cv2_vectors):
    lv = fixture_images["left"]
    for i in range(3):
        assert (sha(lv) == cv2_vectors[f"lkpyr_img{i}_sha"]).all(), f"pyrDown level {i}"
        assert (sha(orc.scharr(lv)) == cv2_vectors[f"lkpyr_der{i}_sha"]).all(), f"Scharr level {i}"
        if i == 2:
            assert (lv == cv2_vectors["lkpyr_img2"]).all()
        lv = orc.pyr_down(lv)


def test_lk_matches_opencv(fixture_images, cv2_vectors):
    v = cv2_vectors
    nxt, st, err = orc.lk(fixture_images["testimage0"], v["lk_img_next"], v["lk_prev"], v["lk_init"], 31)
    assert (st == v["lk_status"]).all()
    ok = st == 1
    assert ok.sum() > 100
    # float summation order differs (OpenCV uses 4-lane SIMD partial sums): tolerance, not bits
    assert np.abs(nxt[ok] - v["lk_next"][ok]).max() < 2e-3
    assert np.median(np.abs(nxt[ok] - v["lk_next"][ok])) < 1e-4
    assert np.abs(err[ok] - v["lk_err"][ok]).max() < 1e-3
    # failed points keep whatever position the pyramid descent left behind — identical in OpenCV
    assert np.abs(nxt[~ok] - v["lk_next"][~ok]).max() < 2e-3


def test_ssd_rule_matches_opencv_at_corners(fixture_images, cv2_vectors):
    cs = orc.CameraSettings(baseline=28.2, fx=470, fy=470, cx=376, cy=240, grid_height=48, grid_width=75, search_x=50,
                            search_y=6, window_size_pose_estimator=4, window_size_opt_flow=31,
                            window_size_depth_calculator=31, max_pyramid_levels=5, min_pyramid_level_pose_estimation=2)
    d = orc.ssd_disparity(fixture_images["left"], fixture_images["right"], cs, cv2_vectors["ssd_pts"], 0)
    want = cv2_vectors["ssd_cv2"][:, 0]
    # cv2 returns a float32 SSD map with DFT noise; at textured (FAST) points the arg-min rule agrees
    agree = np.abs(d - want) < 1e-6
    assert agree.mean() >= 0.97, agree.mean()


def test_ssd_map_exact_vs_numpy(fixture_images):
    L, R = fixture_images["left"], fixture_images["right"]
    tpl = L[100:131, 200:231]
    roi = R[94:137, 200:281]
    m = orc.ssd_map(roi, tpl)
    ref = np.empty_like(m)
    for k in range(m.shape[0]):
        for j in range(m.shape[1]):
            dd = roi[k:k + 31, j:j + 31].astype(np.int64) - tpl.astype(np.int64)
            ref[k, j] = (dd * dd).sum()
    assert (m == ref).all()


def test_rodrigues_project_bit_exact(cv2_vectors):
    v = cv2_vectors
    for r, R in zip(v["rod_in"], v["rod_out_f32"]):
        assert (orc.rodrigues(r) == R).all()
    cs = orc.CameraSettings(fx=v["proj_K"][0, 0], fy=v["proj_K"][1, 1], cx=v["proj_K"][0, 2], cy=v["proj_K"][1, 2],
                            k1=v["proj_D"][0], k2=v["proj_D"][1], p1=v["proj_D"][2], p2=v["proj_D"][3], k3=v["proj_D"][4])
    # project_keypoints passes -rvec and pre-subtracts t (transform_keypoints.cpp:33-45)
    pose = np.concatenate([np.zeros(3, np.float32), -v["proj_r"]]).astype(np.float32)
    got = orc.project(cs, pose, v["proj_P"])
    assert (got == v["proj_out"]).all()
    cs0 = orc.CameraSettings(fx=cs.fx, fy=cs.fy, cx=cs.cx, cy=cs.cy)
    assert (orc.project(cs0, pose, v["proj_P"]) == v["proj_out_nodist"]).all()


def test_svd_inverse_and_solve(cv2_vectors):
    v = cv2_vectors
    ok, Hi = orc.invert_svd(v["inv_H"])
    assert ok == 1
    rel = np.abs(Hi - v["inv_Hinv"]).max() / np.abs(v["inv_Hinv"]).max()
    assert rel < 1e-5, rel
    x = orc.solve_svd(v["solve_A"], v["solve_b"])
    assert np.abs(x - v["solve_x"]).max() < 1e-6
    ok, Z = orc.invert_svd(np.zeros((6, 6), np.float32))
    assert ok == 0 and (Z == 0).all()  # Matx::inv returns zeros when cv::invert reports singular


def test_kalman_matches_opencv(cv2_vectors):
    v = cv2_vectors
    kf = orc.Kalman(12, 12)
    kf.mat("H")[:] = np.eye(12, dtype=np.float32)
    kf.mat("Q")[:] = np.eye(12, dtype=np.float32) * 100
    kf.mat("Ppost")[:] = np.eye(12, dtype=np.float32)
    for i in range(12):
        dt = 0.0 if i % 3 else 0.005
        A = kf.mat("A")
        for k in range(6):
            A[k, k + 6] = dt
        kf.predict()
        R = np.eye(12, dtype=np.float32)
        R[:6, :6] *= 0.1
        kf.mat("R")[:] = R
        kf.correct(v["kf12_z"][i])
        assert np.abs(kf.mat("xpre") - v["kf12_pre"][i]).max() < 1e-5
        assert np.abs(kf.mat("xpost") - v["kf12_post"][i]).max() < 1e-5
    assert np.abs(kf.mat("Ppost") - v["kf12_P"]).max() < 1e-4
    k1 = orc.Kalman(1, 1)
    k1.mat("H")[:] = 1
    k1.mat("Q")[:] = 0.0001
    k1.mat("Ppost")[:] = np.float32((0.5 / (47.9064 / 435.2047)) ** 2)
    k1.mat("xpost")[:] = 0.25
    for Rn, z, x, P in v["kf1_seq"]:
        k1.mat("R")[:] = Rn
        k1.predict()
        k1.correct(np.array([z], np.float32))
        assert abs(k1.mat("xpost")[0] - x) <= 1e-7 * abs(x) + 1e-9
        assert abs(k1.mat("Ppost")[0, 0] - P) <= 2e-7 * abs(P)


def test_exponential_map_known_answer():
    # src/test/test_exponential_map.cpp:37-47 (values; closed form with norm == 1)
    out = orc.expmap([0.1, 0.2, 0.3, 0.4, 0.5, 0.6])
    assert np.abs(out[:3] - np.array([0.12187591, 0.17336931, 0.30760830], np.float32)).max() < 2e-7
    assert (out[3:] == np.array([0.4, 0.5, 0.6], np.float32)).all()


def test_half_sample_truncates():
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (481, 753), dtype=np.uint8)  # odd sizes: trailing row/col dropped
    out = orc.half_sample(img)
    a = img[:480, :752].astype(np.int32)
    want = (a[0::2, 0::2] + a[0::2, 1::2] + a[1::2, 0::2] + a[1::2, 1::2]) // 4
    assert out.shape == (240, 376) and (out == want).all()


def test_rectification_bit_exact(fixture_images, cv2_vectors):
    # cv::initUndistortRectifyMap + cv::remap(INTER_LINEAR) with the EuRoC.yaml calibration (euroc_input.cpp:48-49, :69-73)
    for side, name in (("LEFT", "left"), ("RIGHT", "right")):
        p = cv2_vectors[f"rect_{side}_params"]
        K, D, R, P = p[:9].reshape(3, 3), p[9:14], p[14:23].reshape(3, 3), p[23:32].reshape(3, 3)
        m1, m2 = orc.rectify_map(K, D, R, P, 752, 480)
        assert (sha(np.stack([m1, m2])) == cv2_vectors[f"rect_{side}_map_sha"]).all()
        rect = orc.remap(fixture_images[name], m1, m2)
        assert (sha(rect) == cv2_vectors[f"rect_{side}_img_sha"]).all()
        assert (rect[236:240] == cv2_vectors[f"rect_{side}_rows"]).all()
