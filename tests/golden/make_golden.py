"""Generate the committed golden fixtures (run in the build container, where /root/reference and cv2 exist).

    python tests/golden/make_golden.py

Outputs (committed):
  tests/golden/stereo_fixture.npz   grayscale copies of the reference's own test fixtures
                                    (src/test/left.png, right.png, testimage0.png — 752x480, 3 equal channels)
  tests/golden/cv2_vectors.npz      outputs of the REAL OpenCV (cv2 4.13.0) primitives the reference calls on
                                    its hot path, on those fixtures: the pins for oracle/ocv_prims.hpp.

The `-m gpu` tests and bench.py never read /root/reference; they read these files.
"""
import hashlib
import os

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/src/test"


def sha(a):
    return np.frombuffer(hashlib.sha1(np.ascontiguousarray(a).tobytes()).digest(), np.uint8)


def main():
    cv2.setNumThreads(1)
    imgs = {}
    for name in ("left", "right", "testimage0"):
        im = cv2.imread(os.path.join(REF, name + ".png"), cv2.IMREAD_UNCHANGED)
        assert im.shape == (480, 752, 3) and (im[..., 0] == im[..., 1]).all() and (im[..., 0] == im[..., 2]).all()
        imgs[name] = np.ascontiguousarray(im[..., 0])
    np.savez_compressed(os.path.join(HERE, "stereo_fixture.npz"), **imgs)

    L, R, T = imgs["left"], imgs["right"], imgs["testimage0"]
    out = {}
    # --- FAST(6) + NMS (corner_detector.cpp:19-22)
    for nm, im in (("left", L), ("testimage0", T)):
        det = cv2.FastFeatureDetector_create(6)
        kps = det.detect(im)
        out[f"fast_{nm}"] = np.array([[k.pt[0], k.pt[1], k.response] for k in kps], np.float32)
    # --- Sobel x 8U (corner_detector.cpp:25)
    out["sobel_left_sha"] = sha(cv2.Sobel(L, -1, 1, 0))
    out["sobel_left_rows"] = cv2.Sobel(L, -1, 1, 0)[100:104].copy()
    # --- buildOpticalFlowPyramid (stereo_slam.cpp:139)
    _, pyr = cv2.buildOpticalFlowPyramid(L, (31, 31), 2)
    for lv in range(3):
        out[f"lkpyr_img{lv}_sha"] = sha(pyr[2 * lv])
        out[f"lkpyr_der{lv}_sha"] = sha(pyr[2 * lv + 1])
    out["lkpyr_img2"] = np.ascontiguousarray(pyr[4])
    out["lkpyr_der2_rows"] = np.ascontiguousarray(pyr[5][50:54])
    # --- calcOpticalFlowPyrLK (optical_flow.cpp:41-44): testimage0 vs affine-warped copy
    rng = np.random.default_rng(7)
    M = np.array([[1.004, 0.006, 1.7], [-0.005, 0.997, -2.3]], np.float64)
    T2 = cv2.warpAffine(T, M, (752, 480), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT_101)
    fk = out["fast_testimage0"]
    sel = rng.choice(len(fk), 120, replace=False)
    prev = fk[sel, :2].astype(np.float32)
    extra = np.array([[2.0, 3.0], [750.5, 478.2], [5.5, 470.0], [376.0, 2.0], [10.0, 10.0], [-20.0, 100.0],
                      [700.0, 200.0], [740.0, 20.0]], np.float32)
    prev = np.vstack([prev, extra]).astype(np.float32)
    true_next = (prev @ M[:, :2].T + M[:, 2]).astype(np.float32)
    init = (true_next + rng.uniform(-3, 3, true_next.shape)).astype(np.float32)
    nxt, st, err = cv2.calcOpticalFlowPyrLK(T, T2, prev, init.copy(), winSize=(31, 31), maxLevel=2,
                                            criteria=(cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 30, 0.01),
                                            flags=cv2.OPTFLOW_USE_INITIAL_FLOW, minEigThreshold=1e-4)
    out["lk_img_next"] = T2
    out["lk_prev"], out["lk_init"], out["lk_next"] = prev, init, nxt.reshape(-1, 2)
    out["lk_status"], out["lk_err"] = st.reshape(-1), err.reshape(-1)
    # --- matchTemplate SQDIFF + minMaxLoc at FAST keypoints (depth_calculator.cpp:201-239)
    fl = out["fast_left"]
    sel = rng.choice(len(fl), 200, replace=False)
    pts = fl[sel, :2]
    res = []
    for (x, y) in pts.astype(int):
        wb, wa, sx, sy = 15, 16, 50, 6
        x11, x12 = max(0, x - wb), min(752 - 1, x + wa)
        y11, y12 = max(0, y - wb), min(480, y + wa)
        x21, x22 = max(0, x - wb), min(752 - 1, x + wa + sx)
        y21, y22 = max(0, y - wb - sy), min(480 - 1, y + wa + sy)
        m = cv2.matchTemplate(R[y21:y22, x21:x22], L[y11:y12, x11:x12], cv2.TM_SQDIFF)
        mn, _, loc, _ = cv2.minMaxLoc(m)
        sub = m[loc[1]:, loc[0]:]
        js = np.nonzero(sub <= mn)[1] + loc[0]
        res.append([float(np.float32(js.astype(np.float32).sum()) / np.float32(len(js))), mn, loc[0], loc[1]])
    out["ssd_pts"] = pts.astype(np.float32)
    out["ssd_cv2"] = np.array(res, np.float64)
    # --- Rodrigues / projectPoints (pose_manager.cpp:15-16, transform_keypoints.cpp:45), Econ-like distortion
    rv = rng.normal(0, 0.3, (50, 3)).astype(np.float32)
    rv[0] = 0
    out["rod_in"] = rv
    out["rod_out"] = np.stack([cv2.Rodrigues(r.reshape(3, 1))[0].astype(np.float64) for r in rv])
    out["rod_out_f32"] = np.stack([cv2.Rodrigues(r)[0] for r in rv]).astype(np.float32)
    P = (rng.normal(0, 1, (200, 3)) * [2, 1.5, 1] + [0, 0, 5]).astype(np.float32)
    K = np.array([[435.2047, 0, 367.4517], [0, 435.2047, 252.2009], [0, 0, 1]], np.float32)
    D = np.array([-0.28, 0.07, 0.0002, 1.7e-5, 0.01], np.float32)
    r = np.array([0.02, -0.03, 0.01], np.float32)
    out["proj_P"], out["proj_K"], out["proj_D"], out["proj_r"] = P, K, D, r
    out["proj_out"] = cv2.projectPoints(P, r, np.zeros(3, np.float32), K, D)[0].reshape(-1, 2).astype(np.float32)
    out["proj_out_nodist"] = cv2.projectPoints(P, r, np.zeros(3, np.float32), K, np.zeros(5, np.float32))[0].reshape(-1, 2).astype(np.float32)
    # --- invert / solve SVD (pose_estimator.cpp:405, depth_filter.cpp:200)
    J = rng.normal(0, 1, (300, 6)).astype(np.float32) * np.array([25, 25, 8, 120, 120, 60], np.float32)
    Hm = (J.T @ J).astype(np.float32)
    out["inv_H"] = Hm
    out["inv_Hinv"] = cv2.invert(Hm, flags=cv2.DECOMP_SVD)[1]
    A = rng.normal(0, 1, (3, 2)).astype(np.float32)
    b = rng.normal(0, 1, (3, 1)).astype(np.float32)
    out["solve_A"], out["solve_b"] = A, b
    out["solve_x"] = cv2.solve(A, b, flags=cv2.DECOMP_SVD)[1]
    # --- KalmanFilter 12-state as configured by StereoSlam (stereo_slam.cpp:33-39, :296-359)
    kf = cv2.KalmanFilter(12, 12)
    kf.transitionMatrix = np.eye(12, dtype=np.float32)
    kf.measurementMatrix = np.eye(12, dtype=np.float32)
    kf.processNoiseCov = np.eye(12, dtype=np.float32) * 100
    kf.errorCovPost = np.eye(12, dtype=np.float32)
    kf.statePost = np.zeros((12, 1), np.float32)
    zs, posts, pres = [], [], []
    for i in range(12):
        dt = 0.0 if i % 3 else 0.005
        A_ = np.eye(12, dtype=np.float32)
        for k in range(6):
            A_[k, k + 6] = dt
        kf.transitionMatrix = A_
        kf.predict()
        Rm = np.eye(12, dtype=np.float32)
        Rm[:6, :6] *= 0.1
        kf.measurementNoiseCov = Rm
        z = rng.normal(0, 1, (12, 1)).astype(np.float32)
        kf.correct(z)
        zs.append(z.reshape(-1)); posts.append(kf.statePost.reshape(-1).copy()); pres.append(kf.statePre.reshape(-1).copy())
    out["kf12_z"], out["kf12_post"], out["kf12_pre"] = np.array(zs), np.array(posts), np.array(pres)
    out["kf12_P"] = kf.errorCovPost.copy()
    # 1x1 depth filter (depth_calculator.cpp:277-289, depth_filter.cpp:202-215)
    k1 = cv2.KalmanFilter(1, 1)
    k1.transitionMatrix = np.eye(1, dtype=np.float32)
    k1.measurementMatrix = np.eye(1, dtype=np.float32)
    k1.processNoiseCov = np.eye(1, dtype=np.float32) * 0.0001
    k1.errorCovPost = np.array([[(0.5 / (47.9064 / 435.2047)) ** 2]], np.float32)
    k1.statePost = np.array([[0.25]], np.float32)
    seq = []
    for i in range(20):
        Rn = np.float32(rng.uniform(0.5, 30))
        k1.measurementNoiseCov = np.array([[Rn]], np.float32)
        k1.predict()
        z = np.array([[0.25 + rng.normal(0, 0.02)]], np.float32)
        k1.correct(z)
        seq.append([Rn, z[0, 0], k1.statePost[0, 0], k1.errorCovPost[0, 0]])
    out["kf1_seq"] = np.array(seq, np.float32)
    # --- EuRoC rectification front end (src/app/euroc_input.cpp:48-49, :69-73), calibration of src/app/EuRoC.yaml
    cal = {
        "LEFT": dict(K=[458.654, 0.0, 367.215, 0.0, 457.296, 248.375, 0.0, 0.0, 1.0], D=[-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05, 0.0],
                     R=[0.999966347530033, -0.001422739138722922, 0.008079580483432283, 0.001365741834644127, 0.9999741760894847,
                        0.007055629199258132, -0.008089410156878961, -0.007044357138835809, 0.9999424675829176],
                     P=[435.2046959714599, 0, 367.4517211914062, 0, 0, 435.2046959714599, 252.2008514404297, 0, 0, 0, 1, 0]),
        "RIGHT": dict(K=[457.587, 0.0, 379.999, 0.0, 456.134, 255.238, 0.0, 0.0, 1], D=[-0.28368365, 0.07451284, -0.00010473, -3.555907e-05, 0.0],
                      R=[0.9999633526194376, -0.003625811871560086, 0.007755443660172947, 0.003680398547259526, 0.9999684752771629,
                         -0.007035845251224894, -0.007729688520722713, 0.007064130529506649, 0.999945173484644],
                      P=[435.2046959714599, 0, 367.4517211914062, -47.90639384423901, 0, 435.2046959714599, 252.2008514404297, 0, 0, 0, 1, 0]),
    }
    for side, img in (("LEFT", L), ("RIGHT", R)):
        c = cal[side]
        K_, D_, R_, P_ = np.array(c["K"]).reshape(3, 3), np.array(c["D"]), np.array(c["R"]).reshape(3, 3), np.array(c["P"]).reshape(3, 4)
        m1, m2 = cv2.initUndistortRectifyMap(K_, D_, R_, P_[:3, :3], (752, 480), cv2.CV_32F)
        rect = cv2.remap(img, m1, m2, cv2.INTER_LINEAR)
        out[f"rect_{side}_params"] = np.concatenate([K_.ravel(), D_.ravel(), R_.ravel(), P_[:3, :3].ravel()])
        out[f"rect_{side}_map_sha"] = sha(np.stack([m1, m2]))
        out[f"rect_{side}_img_sha"] = sha(rect)
        out[f"rect_{side}_rows"] = rect[236:240].copy()
    out["cv2_version"] = np.frombuffer(cv2.__version__.encode(), np.uint8)
    np.savez_compressed(os.path.join(HERE, "cv2_vectors.npz"), **out)
    print("wrote golden fixtures:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
