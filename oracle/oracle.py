"""ctypes binding of the CPU oracle (oracle/_build/libsvo_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs. The product package (stereo_svo_slam_b200) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libsvo_oracle.so")


class CameraSettings(C.Structure):
    """Field-for-field the reference's CameraSettings (src/include/stereo_slam_types.hpp:16-36)."""
    _fields_ = [(n, C.c_float) for n in ("baseline", "fx", "fy", "cx", "cy", "k1", "k2", "k3", "p1", "p2")] + \
               [(n, C.c_int) for n in ("grid_height", "grid_width", "search_x", "search_y",
                                       "window_size_pose_estimator", "window_size_opt_flow",
                                       "window_size_depth_calculator", "max_pyramid_levels",
                                       "min_pyramid_level_pose_estimation")]


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ("svo_oracle.cpp", "ocv_prims.hpp", "Makefile")]
    if force or not os.path.exists(_LIB_PATH) or any(
            os.path.exists(s) and os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_slam_create.restype = C.c_void_p
        _lib.orc_kf_create.restype = C.c_void_p
        _lib.orc_kf_mat.restype = C.POINTER(C.c_float)
        _lib.orc_align.restype = C.c_float
        _lib.orc_align_cost.restype = C.c_float
        _lib.orc_refine.restype = C.c_float
    return _lib


def _p(a, t=None):
    return a.ctypes.data_as(C.c_void_p)


def _u8(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def half_sample(img):
    img = _u8(img)
    h, w = img.shape
    out = np.empty((h // 2, w // 2), np.uint8)
    lib().orc_half_sample(_p(img), w, h, w, _p(out))
    return out


def pyr_down(img):
    img = _u8(img)
    h, w = img.shape
    out = np.empty(((h + 1) // 2, (w + 1) // 2), np.uint8)
    lib().orc_pyr_down(_p(img), w, h, _p(out))
    return out


def scharr(img):
    img = _u8(img)
    h, w = img.shape
    out = np.empty((h, w, 2), np.int16)
    lib().orc_scharr(_p(img), w, h, _p(out))
    return out


def sobel_x(img):
    img = _u8(img)
    h, w = img.shape
    out = np.empty((h, w), np.uint8)
    lib().orc_sobel_x(_p(img), w, h, _p(out))
    return out


def fast(img, thr=6, max_kps=200000):
    img = _u8(img)
    h, w = img.shape
    out = np.empty((max_kps, 3), np.int32)
    n = lib().orc_fast(_p(img), w, h, thr, max_kps, _p(out))
    return out[:min(n, max_kps)].copy()


def detect_keypoints(img, grid_w, grid_h, level=0):
    img = _u8(img)
    h, w = img.shape
    mx = (w // max(1, grid_w) + 1) * (h // max(1, grid_h) + 1) + 8
    xy = np.empty((mx, 2), np.float32)
    sc = np.empty(mx, np.float32)
    ty = np.empty(mx, np.int32)
    n = lib().orc_detect_keypoints(_p(img), w, h, grid_w, grid_h, level, mx, _p(xy), _p(sc), _p(ty))
    return xy[:n].copy(), sc[:n].copy(), ty[:n].copy()


def rodrigues(r):
    r = _f32(r)
    R = np.empty((3, 3), np.float32)
    lib().orc_rodrigues(_p(r), _p(R))
    return R


def project(cs, pose6, pts3):
    pts3 = _f32(pts3).reshape(-1, 3)
    pose6 = _f32(pose6)
    out = np.empty((pts3.shape[0], 2), np.float32)
    lib().orc_project(C.byref(cs), _p(pose6), _p(pts3), pts3.shape[0], _p(out))
    return out


def invert_svd(A):
    A = _f32(A)
    n = A.shape[0]
    out = np.empty((n, n), np.float32)
    ok = lib().orc_invert_svd(_p(A), n, _p(out))
    return ok, out


def solve_svd(A, B):
    A = _f32(A)
    B = _f32(B)
    m, n = A.shape
    B2 = B.reshape(m, -1)
    X = np.empty((n, B2.shape[1]), np.float32)
    lib().orc_solve_svd(_p(A), m, n, _p(B2), B2.shape[1], _p(X))
    return X


def expmap(tw):
    tw = _f32(tw)
    out = np.empty(6, np.float32)
    lib().orc_expmap(_p(tw), _p(out))
    return out


def rectify_map(K, D, R, P, w, h):
    a = [np.ascontiguousarray(x, dtype=np.float64).reshape(-1) for x in (K, D, R, np.asarray(P, np.float64)[:3, :3])]
    m1, m2 = np.empty((h, w), np.float32), np.empty((h, w), np.float32)
    lib().orc_rectify_map(_p(a[0]), _p(a[1]), _p(a[2]), _p(a[3]), w, h, _p(m1), _p(m2))
    return m1, m2


def remap(src, m1, m2):
    src = _u8(src)
    m1, m2 = _f32(m1), _f32(m2)
    h, w = m1.shape
    dst = np.empty((h, w), np.uint8)
    lib().orc_remap(_p(src), src.shape[1], src.shape[0], _p(m1), _p(m2), w, h, _p(dst))
    return dst


class Kalman:
    NAMES = {"A": 0, "H": 1, "Q": 2, "R": 3, "Ppre": 4, "Ppost": 5, "xpre": 6, "xpost": 7, "K": 8}

    def __init__(self, n, m):
        self.n, self.m = n, m
        self.h = C.c_void_p(lib().orc_kf_create(n, m))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_kf_destroy(self.h)
            self.h = None

    def _shape(self, name):
        n, m = self.n, self.m
        return {"A": (n, n), "H": (m, n), "Q": (n, n), "R": (m, m), "Ppre": (n, n), "Ppost": (n, n),
                "xpre": (n,), "xpost": (n,), "K": (n, m)}[name]

    def mat(self, name):
        shp = self._shape(name)
        ptr = lib().orc_kf_mat(self.h, self.NAMES[name])
        return np.ctypeslib.as_array(ptr, shape=(int(np.prod(shp)),)).reshape(shp)

    def predict(self):
        lib().orc_kf_predict(self.h)

    def correct(self, z):
        z = _f32(z)
        lib().orc_kf_correct(self.h, _p(z))


def lk(prev_img, next_img, prev_pts, init_pts, win=31):
    prev_img, next_img = _u8(prev_img), _u8(next_img)
    h, w = prev_img.shape
    prev_pts = _f32(prev_pts).reshape(-1, 2)
    nxt = _f32(init_pts).reshape(-1, 2).copy()
    n = prev_pts.shape[0]
    status = np.empty(n, np.uint8)
    err = np.empty(n, np.float32)
    lib().orc_lk(_p(prev_img), _p(next_img), w, h, win, _p(prev_pts), _p(nxt), n, _p(status), _p(err))
    return nxt, status, err


def ssd_disparity(left, right, cs, kps2d, mode):
    left, right = _u8(left), _u8(right)
    h, w = left.shape
    kps2d = _f32(kps2d).reshape(-1, 2)
    out = np.empty(kps2d.shape[0], np.float32)
    lib().orc_ssd_disparity(_p(left), _p(right), w, h, C.byref(cs), _p(kps2d), kps2d.shape[0], mode, _p(out))
    return out


def ssd_map(roi, tpl):
    roi, tpl = _u8(roi), _u8(tpl)
    mh, mw = roi.shape[0] - tpl.shape[0] + 1, roi.shape[1] - tpl.shape[1] + 1
    out = np.empty((mh, mw), np.uint32)
    lib().orc_ssd_map(_p(roi), roi.shape[1], roi.shape[0], roi.shape[1], _p(tpl), tpl.shape[1], tpl.shape[0],
                      tpl.shape[1], _p(out), mh * mw)
    return out


def align(prev_left, cur_left, cs, kps2d, kps3d, pose_in):
    prev_left, cur_left = _u8(prev_left), _u8(cur_left)
    h, w = prev_left.shape
    kps2d, kps3d, pose_in = _f32(kps2d).reshape(-1, 2), _f32(kps3d).reshape(-1, 3), _f32(pose_in)
    pose_out = np.empty(6, np.float32)
    evals = np.zeros(16, np.int32)
    cost = lib().orc_align(_p(prev_left), _p(cur_left), w, h, C.byref(cs), _p(kps2d), _p(kps3d), kps2d.shape[0],
                           _p(pose_in), _p(pose_out), _p(evals))
    return pose_out, float(cost), evals.reshape(8, 2)


def align_cost(prev_left, cur_left, cs, kps2d, kps3d, pose, level, want_grad=True):
    prev_left, cur_left = _u8(prev_left), _u8(cur_left)
    h, w = prev_left.shape
    kps2d, kps3d, pose = _f32(kps2d).reshape(-1, 2), _f32(kps3d).reshape(-1, 3), _f32(pose)
    grad = np.zeros(6, np.float32)
    cost = lib().orc_align_cost(_p(prev_left), _p(cur_left), w, h, C.byref(cs), _p(kps2d), _p(kps3d), kps2d.shape[0],
                                _p(pose), level, _p(grad) if want_grad else None)
    return float(cost), grad


def refine(cs, kps2d, kps3d, flags, pose_in):
    kps2d, kps3d, pose_in = _f32(kps2d).reshape(-1, 2), _f32(kps3d).reshape(-1, 3), _f32(pose_in)
    flags = np.ascontiguousarray(flags, dtype=np.int32)
    pose_out = np.empty(6, np.float32)
    ev = np.zeros(2, np.int32)
    cost = lib().orc_refine(C.byref(cs), _p(kps2d), _p(kps3d), _p(flags), kps2d.shape[0], _p(pose_in), _p(pose_out),
                            _p(ev))
    return pose_out, float(cost), ev


class OracleSlam:
    """CPU restatement of the reference's StereoSlam (new_image / pose / trajectory / keyframes)."""

    def __init__(self, cs, width, height, tracing=True):
        self.cs, self.w, self.h_ = cs, width, height
        self.h = C.c_void_p(lib().orc_slam_create(C.byref(cs), width, height))
        lib().orc_slam_set_tracing(self.h, 1 if tracing else 0)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_slam_destroy(self.h)
            self.h = None

    def new_image(self, left, right, ts):
        left, right = _u8(left), _u8(right)
        assert left.shape == (self.h_, self.w) and right.shape == left.shape
        lib().orc_slam_new_image(self.h, _p(left), _p(right), self.w, C.c_float(ts))

    def update_pose(self, pose, speed, pose_var, speed_var, dt):
        out = np.empty(6, np.float32)
        a = [_f32(x) for x in (pose, speed, pose_var, speed_var)]
        lib().orc_slam_update_pose(self.h, _p(a[0]), _p(a[1]), _p(a[2]), _p(a[3]), C.c_double(dt), _p(out))
        return out

    def pose(self):
        p = np.empty(6, np.float32)
        lib().orc_slam_get_pose(self.h, _p(p))
        return p

    def stage_ms(self):
        """[pyramids, alignment, klt, refine-gn, ssd, filter, keyframe, total] of the last new_image (ms)."""
        out = np.zeros(8, np.float64)
        lib().orc_slam_stage_ms(self.h, _p(out))
        return out

    def n_kps(self):
        return lib().orc_slam_n_kps(self.h)

    def n_keyframes(self):
        return lib().orc_slam_n_keyframes(self.h)

    def trajectory(self):
        n = lib().orc_slam_trajectory(self.h, None, 0)
        out = np.empty((n, 6), np.float32)
        lib().orc_slam_trajectory(self.h, _p(out), n)
        return out

    def trace(self, name):
        n = lib().orc_slam_trace(self.h, name.encode(), None, 0)
        if n < 0:
            return None
        out = np.empty(n, np.float32)
        lib().orc_slam_trace(self.h, name.encode(), _p(out), n)
        return out

    def keyframe(self, k):
        n = lib().orc_slam_keyframe(self.h, k, None, None, None, 0)
        if n < 0:
            return None
        pose = np.empty(6, np.float32)
        k2 = np.empty((n, 2), np.float32)
        k3 = np.empty((n, 3), np.float32)
        lib().orc_slam_keyframe(self.h, k, _p(pose), _p(k2), _p(k3), n)
        return pose, k2, k3
