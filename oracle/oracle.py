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
# the reference itself (unmodified /root/reference/src/lib/*.cpp against oracle/cvshim), see oracle/Makefile
_REF_PATH = os.path.join(_HERE, "_ref", "libstereosvo_ref.so")
_REF_SOURCES = "/root/reference/src/lib"


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
        subprocess.check_call(["make", "-C", _HERE, "-s", "_build/libsvo_oracle.so"] + (["-B"] if force else []))
    build_ref(force)
    return _LIB_PATH


def build_ref(force=False):
    """oracle/_ref/libstereosvo_ref.so from the reference's own sources; only where /root/reference exists (this
    container) — the GPU box uses the prebuilt file that travelled with the snapshot. Returns the path or None."""
    if os.path.isdir(_REF_SOURCES):
        src = [os.path.join(_HERE, f) for f in ("ref_capi.cpp", "ocv_prims.hpp", "Makefile",
                                                os.path.join("cvshim", "opencv2", "opencv.hpp"))]
        if force or not os.path.exists(_REF_PATH) or any(
                os.path.getmtime(s) > os.path.getmtime(_REF_PATH) for s in src):
            subprocess.check_call(["make", "-C", _HERE, "-s", "ref"] + (["-B"] if force else []))
    return _REF_PATH if os.path.exists(_REF_PATH) else None


def have_ref():
    return build_ref() is not None


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


def lk_iters(prev_img, next_img, prev_pts, init_pts, win=31):
    """lk() plus, per point, the largest iteration count of any level (30: that level did not converge)."""
    prev_img, next_img = _u8(prev_img), _u8(next_img)
    h, w = prev_img.shape
    prev_pts = _f32(prev_pts).reshape(-1, 2)
    nxt = _f32(init_pts).reshape(-1, 2).copy()
    n = prev_pts.shape[0]
    status, err, it = np.empty(n, np.uint8), np.empty(n, np.float32), np.zeros(n, np.int32)
    lib().orc_lk_iters(_p(prev_img), _p(next_img), w, h, win, _p(prev_pts), _p(nxt), n, _p(status), _p(err), _p(it))
    return nxt, status, err, it


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


def _pp(arrs, ctype):
    """array of pointers to the rows of a ragged list"""
    return (C.POINTER(ctype) * len(arrs))(*[a.ctypes.data_as(C.POINTER(ctype)) for a in arrs])


def select_best(levels):
    """select_best_keypoints (depth_calculator.cpp:37-65); levels = [(xy (n,2) f32, score (n) f32, type (n) i32), ...]."""
    xy = [_f32(l[0]).reshape(-1, 2) for l in levels]
    sc = [_f32(l[1]) for l in levels]
    ty = [np.ascontiguousarray(l[2], dtype=np.int32) for l in levels]
    n0 = len(sc[0])
    cnt = np.array([len(x) for x in sc], np.int32)
    k2, so, to, lo = np.zeros((n0, 2), np.float32), np.zeros(n0, np.float32), np.zeros(n0, np.int32), np.zeros(n0, np.int32)
    n = lib().orc_select_best(len(levels), _p(cnt), _pp(xy, C.c_float), _pp(sc, C.c_float), _pp(ty, C.c_int), _p(k2), _p(so), _p(to), _p(lo))
    assert n == n0
    return k2, so, to, lo


def find_bad(w, h, kps2d, flags):
    kps2d = _f32(kps2d).reshape(-1, 2)
    flags = np.ascontiguousarray(flags, dtype=np.uint8)
    keep = np.zeros(len(flags), np.uint8)
    lib().orc_find_bad(w, h, len(flags), _p(kps2d), _p(flags), _p(keep))
    return keep


def merge(w, h, grid_width, grid_height, old2d, new2d):
    old2d, new2d = _f32(old2d).reshape(-1, 2), _f32(new2d).reshape(-1, 2)
    app = np.zeros(max(1, 64 * len(new2d) + 64), np.int32)
    n = lib().orc_merge(w, h, grid_width, grid_height, len(old2d), _p(old2d), len(new2d), _p(new2d), _p(app))
    return app[:n].copy()


def keyframe_needed(w, h, grid_width, grid_height, kps2d, flags):
    kps2d = _f32(kps2d).reshape(-1, 2)
    flags = np.ascontiguousarray(flags, dtype=np.uint8)
    return bool(lib().orc_keyframe_needed(w, h, grid_width, grid_height, len(flags), _p(kps2d), _p(flags)))


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

    def frame(self):
        """Current frame as a dict (pose, id, ts, kps2d, kps3d, info columns) — same shape as RefSlam.frame()."""
        return _dump(lambda *a: lib().orc_slam_frame(self.h, *a), True)

    def keyframe_full(self, k):
        return _dump(lambda *a: lib().orc_slam_keyframe_full(self.h, k, *a), False)

    def keyframe(self, k):
        n = lib().orc_slam_keyframe(self.h, k, None, None, None, 0)
        if n < 0:
            return None
        pose = np.empty(6, np.float32)
        k2 = np.empty((n, 2), np.float32)
        k3 = np.empty((n, 3), np.float32)
        lib().orc_slam_keyframe(self.h, k, _p(pose), _p(k2), _p(k3), n)
        return pose, k2, k3


INFO_COLS = ("level", "type", "keyframe_id", "keypoint_index", "flags", "inlier_count", "outlier_count", "color")


def _dump(call, is_frame):
    """Two-pass flat dump (count, then arrays) shared by the oracle and the reference handle APIs."""
    pose = np.zeros(6, np.float32)
    fid = C.c_uint64(0)
    ts = C.c_double(0)
    head = (_p(pose), C.byref(fid)) + ((C.byref(ts),) if is_frame else ())
    n = call(*head, None, None, None, None, 0)
    if n < 0:
        return None
    k2, k3 = np.zeros((n, 2), np.float32), np.zeros((n, 3), np.float32)
    info, finfo = np.zeros((n, 8), np.int32), np.zeros((n, 3), np.float32)
    call(*head, _p(k2), _p(k3), _p(info), _p(finfo), n)
    d = {"pose": pose, "id": int(fid.value), "kps2d": k2, "kps3d": k3, "score": finfo[:, 0].copy(),
         "kf_state": finfo[:, 1].copy(), "kf_cov": finfo[:, 2].copy()}
    if is_frame:
        d["ts"] = float(ts.value)
    for i, c in enumerate(INFO_COLS):
        d[c] = info[:, i].copy()
    return d


class RefSlam:
    """The reference's own StereoSlam (oracle/_ref, built from the unmodified /root/reference/src/lib sources).

    The library keeps process-global counters (depth_calculator.cpp:135, keyframe_manager.cpp:8), so every instance
    loads a private copy of the shared object."""

    def __init__(self, cs, width, height, quiet=True):
        import shutil
        import tempfile
        path = build_ref()
        if path is None:
            raise RuntimeError("oracle/_ref/libstereosvo_ref.so is not built and /root/reference is absent")
        fd, self._tmp = tempfile.mkstemp(prefix="libstereosvo_ref_", suffix=".so")
        os.close(fd)
        shutil.copyfile(path, self._tmp)
        self.l = C.CDLL(self._tmp)
        self.l.ref_slam_create.restype = C.c_void_p
        assert self.l.ref_sizeof_camera_settings() == C.sizeof(CameraSettings)
        self.l.ref_set_quiet(1 if quiet else 0)
        self.cs, self.w, self.h_ = cs, width, height
        self.h = C.c_void_p(self.l.ref_slam_create(C.byref(cs), width, height))

    def close(self):
        if getattr(self, "h", None):
            self.l.ref_slam_destroy(self.h)
            self.h = None
        if getattr(self, "_tmp", None):
            try:
                os.unlink(self._tmp)
            except OSError:
                pass
            self._tmp = None

    __del__ = close

    def new_image(self, left, right, ts):
        left, right = _u8(left), _u8(right)
        assert left.shape == (self.h_, self.w) and right.shape == left.shape
        if self.l.ref_slam_new_image(self.h, _p(left), _p(right), self.w, C.c_float(ts)) != 0:
            raise RuntimeError("reference new_image threw")

    def update_pose(self, pose, speed, pose_var, speed_var, dt):
        out = np.empty(6, np.float32)
        a = [_f32(x) for x in (pose, speed, pose_var, speed_var)]
        self.l.ref_slam_update_pose(self.h, _p(a[0]), _p(a[1]), _p(a[2]), _p(a[3]), C.c_double(dt), _p(out))
        return out

    def pose(self):
        p = np.zeros(6, np.float32)
        self.l.ref_slam_get_pose(self.h, _p(p))
        return p

    def frame(self):
        return _dump(lambda *a: self.l.ref_slam_frame(self.h, *a), True)

    def n_kps(self):
        return self.l.ref_slam_frame(self.h, None, None, None, None, None, None, None, 0)

    def n_keyframes(self):
        return self.l.ref_slam_n_keyframes(self.h)

    def keyframe_full(self, k):
        return _dump(lambda *a: self.l.ref_slam_keyframe(self.h, k, *a), False)

    def keyframe_image(self, which, level):
        """which: 0 left halfSample level, 1 right, 2 LK pyramid image level, 3 LK Scharr level (int16, 2 channels)."""
        w, h = C.c_int(0), C.c_int(0)
        nb = self.l.ref_slam_keyframe_image(self.h, which, level, None, 0, C.byref(w), C.byref(h))
        if nb < 0:
            return None
        buf = np.empty(nb, np.uint8)
        self.l.ref_slam_keyframe_image(self.h, which, level, _p(buf), nb, C.byref(w), C.byref(h))
        if which == 3:
            return buf.view(np.int16).reshape(h.value, w.value, 2)
        return buf.reshape(h.value, w.value)

    def trajectory(self):
        n = self.l.ref_slam_trajectory(self.h, None, 0)
        out = np.empty((n, 6), np.float32)
        self.l.ref_slam_trajectory(self.h, _p(out), n)
        return out

    # stage entry points of the reference (stateless)
    def expmap(self, tw):
        tw = _f32(tw)
        out = np.empty(6, np.float32)
        self.l.ref_expmap(_p(tw), _p(out))
        return out

    def project(self, pose6, pts3):
        pts3 = _f32(pts3).reshape(-1, 3)
        pose6 = _f32(pose6)
        out = np.empty((pts3.shape[0], 2), np.float32)
        self.l.ref_project(C.byref(self.cs), _p(pose6), _p(pts3), pts3.shape[0], _p(out))
        return out

    def detect_keypoints(self, img, grid_w, grid_h, level=0):
        img = _u8(img)
        h, w = img.shape
        mx = (w // max(1, grid_w) + 1) * (h // max(1, grid_h) + 1) + 8
        xy = np.empty((mx, 2), np.float32)
        sc = np.empty(mx, np.float32)
        ty = np.empty(mx, np.int32)
        n = self.l.ref_detect_keypoints(_p(img), w, h, grid_w, grid_h, level, mx, _p(xy), _p(sc), _p(ty))
        return xy[:n].copy(), sc[:n].copy(), ty[:n].copy()
