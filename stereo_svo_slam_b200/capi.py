"""ctypes binding of the C-ABI in include/svo_cuda.h (libstereosvo_b200.so).

This is the product path: it loads the CUDA library and fails loudly if it is missing or if there is no
CUDA device — there is no CPU fallback and nothing here touches oracle/.
"""
import ctypes as C
import os
import re

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libstereosvo_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "svo_cuda.h")

SVO_OK, SVO_ERR_INVALID, SVO_ERR_CUDA, SVO_ERR_NO_DEVICE, SVO_ERR_CAPACITY, SVO_ERR_STATE = range(6)
FLAG_IGNORE_REFINEMENT, FLAG_IGNORE_COMPLETELY, FLAG_IGNORE_TEMPORARY = 1, 2, 4
KP_FAST, KP_EDGELET = 0, 1


class SvoError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"svo error {code}: {msg}")
        self.code = code


class CameraSettings(C.Structure):
    """== struct CameraSettings of the reference (src/include/stereo_slam_types.hpp:16-36)."""
    _fields_ = [(n, C.c_float) for n in ("baseline", "fx", "fy", "cx", "cy", "k1", "k2", "k3", "p1", "p2")] + \
               [(n, C.c_int) for n in ("grid_height", "grid_width", "search_x", "search_y",
                                       "window_size_pose_estimator", "window_size_opt_flow",
                                       "window_size_depth_calculator", "max_pyramid_levels",
                                       "min_pyramid_level_pose_estimation")]


class Pose(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("x", "y", "z", "rx", "ry", "rz")]

    def vec(self):
        return np.array([self.x, self.y, self.z, self.rx, self.ry, self.rz], np.float32)


class KeyPointInfo(C.Structure):
    _fields_ = [("score", C.c_float), ("level", C.c_int), ("type", C.c_int), ("keyframe_id", C.c_uint64),
                ("keypoint_index", C.c_uint64), ("color", C.c_uint8 * 3), ("ignore_during_refinement", C.c_uint8),
                ("ignore_completely", C.c_uint8), ("ignore_temporary", C.c_uint8), ("outlier_count", C.c_int),
                ("inlier_count", C.c_int), ("kf_inv_depth", C.c_float), ("kf_variance", C.c_float)]


KPINFO_DTYPE = np.dtype(KeyPointInfo)


class TrackIO(C.Structure):
    _fields_ = [("n", C.c_int), ("prev_kps2d", C.c_void_p), ("kps3d", C.c_void_p), ("ref_kps2d", C.c_void_p),
                ("keyframe_id", C.c_void_p), ("flags", C.c_void_p), ("inlier_count", C.c_void_p),
                ("outlier_count", C.c_void_p), ("kf_state", C.c_void_p), ("pose_prior", C.c_float * 6),
                ("kps2d", C.c_void_p), ("pose_aligned", C.c_float * 6), ("pose_refined", C.c_float * 6),
                ("align_cost", C.c_float), ("refine_cost", C.c_float), ("align_evals", C.c_int * 16),
                ("refine_evals", C.c_int * 2), ("klt_pts", C.c_void_p), ("klt_err", C.c_void_p),
                ("klt_status", C.c_void_p), ("disparity", C.c_void_p), ("kps2d_refine_in", C.c_void_p),
                ("klt_iters", C.c_void_p), ("keypoint_index", C.c_void_p)]


def declared_symbols():
    """Every function the header declares (used by the CPU test that the library exports all of them)."""
    src = open(HEADER_PATH).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(svo_[a-z0-9_]+)\s*\(", src)))


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SvoError(SVO_ERR_NO_DEVICE, f"{LIB_PATH} is missing: run `python -m stereo_svo_slam_b200.build` "
                                              "(the CUDA extension is mandatory; there is no CPU fallback)")
        _lib = C.CDLL(LIB_PATH)
        _lib.svo_last_error.restype = C.c_char_p
        _lib.svo_last_error.argtypes = [C.c_void_p]
        _lib.svo_slam_last_error.restype = C.c_char_p
        _lib.svo_slam_last_error.argtypes = [C.c_void_p]
        _lib.svo_slam_ctx.restype = C.c_void_p
        _lib.svo_slam_ctx.argtypes = [C.c_void_p]
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a if shape is None else a.reshape(shape)


def settings_from_dict(d):
    return CameraSettings(**d)


class Context:
    """Device layer: svo_ctx_* / svo_upload_stereo / per-stage entry points."""

    def __init__(self, settings, width, height, device=0, max_keypoints=0, _borrowed=None):
        self.settings, self.w, self.h = settings, width, height
        self._owned = _borrowed is None
        if _borrowed is not None:
            self.h_ctx = C.c_void_p(_borrowed)
            return
        out = C.c_void_p()
        rc = lib().svo_ctx_create(C.byref(settings), device, width, height, max_keypoints, C.byref(out))
        if rc:
            raise SvoError(rc, lib().svo_last_error(None).decode())
        self.h_ctx = out

    def _ck(self, rc):
        if rc:
            raise SvoError(rc, lib().svo_last_error(self.h_ctx).decode())

    def close(self):
        if getattr(self, "h_ctx", None) and self._owned:
            lib().svo_ctx_destroy(self.h_ctx)
        self.h_ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- image sets
    def upload(self, left, right):
        assert left.dtype == np.uint8 and right.dtype == np.uint8 and left.shape == (self.h, self.w) and right.shape == left.shape
        assert left.strides[1] == 1 and right.strides[1] == 1
        slot = C.c_int()
        self._ck(lib().svo_upload_stereo(self.h_ctx, _p(left), C.c_size_t(left.strides[0]), _p(right),
                                         C.c_size_t(right.strides[0]), C.byref(slot)))
        return slot.value

    def set_rectification(self, which, K, D, R, P):
        """EuRoC front end (euroc_input.cpp:48-49): rectify every image uploaded as left (which=0) / right (which=1)."""
        a = [np.ascontiguousarray(np.asarray(x, np.float64).reshape(-1)) for x in (K, D, R, np.asarray(P, np.float64).reshape(3, -1)[:, :3])]
        assert a[0].size == 9 and a[1].size == 5 and a[2].size == 9 and a[3].size == 9
        self._ck(lib().svo_set_rectification(self.h_ctx, which, _p(a[0]), _p(a[1]), _p(a[2]), _p(a[3])))

    def clear_rectification(self):
        self._ck(lib().svo_clear_rectification(self.h_ctx))

    def rectification_maps(self, which):
        m1, m2 = np.empty((self.h, self.w), np.float32), np.empty((self.h, self.w), np.float32)
        self._ck(lib().svo_rectification_maps(self.h_ctx, which, _p(m1), _p(m2)))
        return m1, m2

    def set_align_cluster(self, ctas):
        """SMs sharing the alignment solve of a frame: 8 = lowest latency (default), 1-2 = lowest SM time per frame."""
        self._ck(lib().svo_set_align_cluster(self.h_ctx, ctas))

    def set_solver_width(self, wide):
        """Line searches of the two solvers: 1 = several trial poses per round on idle SMs (one sequence alone), 0 = sequential,
        -1 = automatic (svo_set_solver_width); the same bits either way."""
        self._ck(lib().svo_set_solver_width(self.h_ctx, wide))

    def release(self, slot):
        self._ck(lib().svo_slot_release(self.h_ctx, slot))

    def retain(self, slot):
        self._ck(lib().svo_slot_retain(self.h_ctx, slot))

    def level_size(self, kind, level):
        w, h = C.c_int(), C.c_int()
        self._ck(lib().svo_slot_level_size(self.h_ctx, kind, level, C.byref(w), C.byref(h)))
        return w.value, h.value

    def download(self, slot, kind, level):
        """kind 0/1/2: u8 level; 3: Scharr level (h, w, 2) int16; 4: LK image level with its 32-px frame; 5: Scharr level with its frame."""
        w, h = self.level_size(kind, level)
        if kind in (3, 5):
            out = np.empty((h, w, 2), np.int16)
            self._ck(lib().svo_download_level(self.h_ctx, slot, kind, level, _p(out), C.c_size_t(4 * w)))
            return out
        out = np.empty((h, w), np.uint8)
        self._ck(lib().svo_download_level(self.h_ctx, slot, kind, level, _p(out), C.c_size_t(w)))
        return out

    def sync(self):
        self._ck(lib().svo_sync(self.h_ctx))

    # ---- stages
    def detect_keypoints(self, slot, level, grid_w, grid_h):
        w, h = self.level_size(0, level)
        cap = (w // grid_w + 1) * (h // grid_h + 1) + 4
        xy, sc, ty = np.empty((cap, 2), np.float32), np.empty(cap, np.float32), np.empty(cap, np.int32)
        n = C.c_int()
        self._ck(lib().svo_detect_keypoints(self.h_ctx, slot, level, grid_w, grid_h, cap, _p(xy), _p(sc), _p(ty), C.byref(n)))
        return xy[:n.value].copy(), sc[:n.value].copy(), ty[:n.value].copy()

    def fast_corners(self, slot, level, max_out=200000):
        out = np.empty((max_out, 3), np.int32)
        n = C.c_int()
        self._ck(lib().svo_fast_corners(self.h_ctx, slot, level, max_out, _p(out), C.byref(n)))
        return out[:min(n.value, max_out)].copy()

    def stereo_match(self, slot, kps2d, mode):
        kps2d = _f32(kps2d, (-1, 2))
        out = np.empty(kps2d.shape[0], np.float32)
        self._ck(lib().svo_stereo_match(self.h_ctx, slot, _p(kps2d), kps2d.shape[0], mode, _p(out)))
        return out

    def align(self, prev_slot, cur_slot, kps2d, kps3d, pose_in, flags=None):
        kps2d, kps3d, pose_in = _f32(kps2d, (-1, 2)), _f32(kps3d, (-1, 3)), _f32(pose_in)
        fl = None if flags is None else np.ascontiguousarray(flags, dtype=np.uint8)
        pose_out, cost, ev = np.empty(6, np.float32), C.c_float(), np.zeros(16, np.int32)
        self._ck(lib().svo_align(self.h_ctx, prev_slot, cur_slot, _p(kps2d), _p(kps3d), _p(fl), kps2d.shape[0], _p(pose_in),
                                 _p(pose_out), C.byref(cost), _p(ev)))
        return pose_out, cost.value, ev.reshape(8, 2)

    def align_probe(self, prev_slot, cur_slot, kps2d, kps3d, level, pose):
        kps2d, kps3d, pose = _f32(kps2d, (-1, 2)), _f32(kps3d, (-1, 3)), _f32(pose)
        cost, grad = C.c_float(), np.empty(6, np.float32)
        self._ck(lib().svo_align_probe(self.h_ctx, prev_slot, cur_slot, _p(kps2d), _p(kps3d), kps2d.shape[0], level, _p(pose),
                                       C.byref(cost), _p(grad)))
        return cost.value, grad

    def klt_slots(self, prev_slot, cur_slot, prev_pts, init_pts):
        prev_pts, init_pts = _f32(prev_pts, (-1, 2)), _f32(init_pts, (-1, 2))
        n = prev_pts.shape[0]
        nxt, st, err = np.empty((n, 2), np.float32), np.empty(n, np.uint8), np.empty(n, np.float32)
        self._ck(lib().svo_klt_slots(self.h_ctx, prev_slot, cur_slot, _p(prev_pts), _p(init_pts), n, _p(nxt), _p(st), _p(err)))
        return nxt, st, err

    def klt(self, keyframe_ids, cur_slot, prev_pts, init_pts):
        prev_pts, init_pts = _f32(prev_pts, (-1, 2)), _f32(init_pts, (-1, 2))
        ids = np.ascontiguousarray(keyframe_ids, dtype=np.int32)
        n = prev_pts.shape[0]
        nxt, st, err = np.empty((n, 2), np.float32), np.empty(n, np.uint8), np.empty(n, np.float32)
        self._ck(lib().svo_klt(self.h_ctx, _p(ids), cur_slot, _p(prev_pts), _p(init_pts), n, _p(nxt), _p(st), _p(err)))
        return nxt, st, err

    def reproj_refine(self, kps2d, kps3d, flags, pose_in):
        kps2d, kps3d, pose_in = _f32(kps2d, (-1, 2)), _f32(kps3d, (-1, 3)), _f32(pose_in)
        fl = np.ascontiguousarray(flags, dtype=np.uint8)
        pose_out, cost, ev = np.empty(6, np.float32), C.c_float(), np.zeros(2, np.int32)
        self._ck(lib().svo_reproj_refine(self.h_ctx, _p(kps2d), _p(kps3d), _p(fl), kps2d.shape[0], _p(pose_in), _p(pose_out),
                                         C.byref(cost), _p(ev)))
        return pose_out, cost.value, ev

    def depth_filter_update(self, slot, kps2d, ref_kps2d, keyframe_id, kps3d, flags, inlier, outlier, kf_state, pose):
        """DepthFilter::update_depth + post-processing as one stage (svo_depth_filter_update); inputs are not modified."""
        a = dict(kps2d=_f32(kps2d, (-1, 2)).copy(), ref_kps2d=_f32(ref_kps2d, (-1, 2)).copy(),
                 keyframe_id=np.ascontiguousarray(keyframe_id, dtype=np.int32).copy(), kps3d=_f32(kps3d, (-1, 3)).copy(),
                 flags=np.ascontiguousarray(flags, dtype=np.uint8).copy(), inlier=np.ascontiguousarray(inlier, dtype=np.int32).copy(),
                 outlier=np.ascontiguousarray(outlier, dtype=np.int32).copy(), kf_state=_f32(kf_state, (-1, 2)).copy())
        n = len(a["flags"])
        a["disparity"], a["kps2d_out"] = np.empty(n, np.float32), np.empty((n, 2), np.float32)
        pose = _f32(pose)
        self._ck(lib().svo_depth_filter_update(self.h_ctx, slot, n, _p(a["kps2d"]), _p(a["ref_kps2d"]), _p(a["keyframe_id"]), _p(a["kps3d"]),
                                               _p(a["flags"]), _p(a["inlier"]), _p(a["outlier"]), _p(a["kf_state"]), _p(pose),
                                               _p(a["disparity"]), _p(a["kps2d_out"])))
        return a

    def project(self, pose, kps3d):
        kps3d, pose = _f32(kps3d, (-1, 3)), _f32(pose)
        out = np.empty((kps3d.shape[0], 2), np.float32)
        self._ck(lib().svo_project(self.h_ctx, _p(pose), _p(kps3d), kps3d.shape[0], _p(out)))
        return out

    def keyframe_commit(self, slot, pose):
        pose = _f32(pose)
        out = C.c_int()
        self._ck(lib().svo_keyframe_commit(self.h_ctx, slot, _p(pose), C.byref(out)))
        return out.value

    def keyframe_set_templates(self, keyframe_id, kps2d, first=0):
        """LK template cache of the keypoints a keyframe introduced (svo_keyframe_set_templates)."""
        kps2d = _f32(kps2d, (-1, 2))
        self._ck(lib().svo_keyframe_set_templates(self.h_ctx, keyframe_id, _p(kps2d), first, kps2d.shape[0]))

    def track_frame(self, prev_slot, cur_slot, prev_kps2d, kps3d, ref_kps2d, keyframe_id, flags, inlier, outlier, kf_state,
                    pose_prior, keypoint_index=None):
        """Fused per-frame tracking (svo_track_frame). Returns a dict of outputs; inputs are not modified."""
        n = len(flags)
        a = dict(prev_kps2d=_f32(prev_kps2d, (-1, 2)).copy(), kps3d=_f32(kps3d, (-1, 3)).copy(),
                 ref_kps2d=_f32(ref_kps2d, (-1, 2)).copy(),
                 keyframe_id=np.ascontiguousarray(keyframe_id, dtype=np.int32).copy(),
                 flags=np.ascontiguousarray(flags, dtype=np.uint8).copy(),
                 inlier=np.ascontiguousarray(inlier, dtype=np.int32).copy(),
                 outlier=np.ascontiguousarray(outlier, dtype=np.int32).copy(),
                 kf_state=_f32(kf_state, (-1, 2)).copy(),
                 kps2d=np.empty((n, 2), np.float32), klt_pts=np.empty((n, 2), np.float32), klt_err=np.empty(n, np.float32),
                 klt_status=np.empty(n, np.uint8), disparity=np.empty(n, np.float32), kps2d_refine_in=np.empty((n, 2), np.float32),
                 klt_iters=np.zeros(n, np.int32))
        io = TrackIO()
        io.n = n
        io.prev_kps2d, io.kps3d, io.ref_kps2d, io.keyframe_id = _p(a["prev_kps2d"]), _p(a["kps3d"]), _p(a["ref_kps2d"]), _p(a["keyframe_id"])
        io.flags, io.inlier_count, io.outlier_count, io.kf_state = _p(a["flags"]), _p(a["inlier"]), _p(a["outlier"]), _p(a["kf_state"])
        io.kps2d, io.klt_pts, io.klt_err, io.klt_status = _p(a["kps2d"]), _p(a["klt_pts"]), _p(a["klt_err"]), _p(a["klt_status"])
        io.disparity, io.kps2d_refine_in = _p(a["disparity"]), _p(a["kps2d_refine_in"])
        io.klt_iters = _p(a["klt_iters"])
        if keypoint_index is not None:
            a["keypoint_index"] = np.ascontiguousarray(keypoint_index, dtype=np.int32).copy()
            io.keypoint_index = _p(a["keypoint_index"])
        for k, v in enumerate(_f32(pose_prior)):
            io.pose_prior[k] = float(v)
        self._ck(lib().svo_track_frame(self.h_ctx, prev_slot, cur_slot, C.byref(io)))
        a["pose_aligned"] = np.array(list(io.pose_aligned), np.float32)
        a["pose_refined"] = np.array(list(io.pose_refined), np.float32)
        a["align_cost"], a["refine_cost"] = io.align_cost, io.refine_cost
        a["align_evals"] = np.array(list(io.align_evals), np.int32).reshape(8, 2)
        a["refine_evals"] = np.array(list(io.refine_evals), np.int32)
        return a

    def last_track_timing(self):
        ms, n = C.c_float(), C.c_int()
        self._ck(lib().svo_last_track_timing(self.h_ctx, C.byref(ms), C.byref(n)))
        return ms.value, n.value


def device_count():
    return lib().svo_device_count()


# ---- host stages (svo_host_*, svo_motion_filter_*): pure host code of the library, no device needed --------------------
def _pp(arrs, ctype):
    return (C.POINTER(ctype) * len(arrs))(*[a.ctypes.data_as(C.POINTER(ctype)) for a in arrs])


def _host_ck(rc):
    if rc:
        raise SvoError(rc, "host stage: invalid argument")


def host_select_best_keypoints(levels):
    """select_best_keypoints (depth_calculator.cpp:37-65); levels = [(xy (n,2), score (n), type (n)), ...] finest first."""
    xy = [_f32(l[0], (-1, 2)) for l in levels]
    sc = [_f32(l[1]) for l in levels]
    ty = [np.ascontiguousarray(l[2], dtype=np.int32) for l in levels]
    cnt = np.array([len(x) for x in sc], np.int32)
    n0 = int(cnt[0])
    k2, so, to, lo = np.zeros((n0, 2), np.float32), np.zeros(n0, np.float32), np.zeros(n0, np.int32), np.zeros(n0, np.int32)
    n = C.c_int()
    _host_ck(lib().svo_host_select_best_keypoints(len(levels), _p(cnt), _pp(xy, C.c_float), _pp(sc, C.c_float), _pp(ty, C.c_int), n0,
                                                  _p(k2), _p(so), _p(to), _p(lo), C.byref(n)))
    assert n.value == n0
    return k2, so, to, lo


def host_find_bad_keypoints(width, height, kps2d, flags):
    kps2d = _f32(kps2d, (-1, 2))
    flags = np.ascontiguousarray(flags, dtype=np.uint8)
    keep = np.zeros(len(flags), np.uint8)
    _host_ck(lib().svo_host_find_bad_keypoints(width, height, len(flags), _p(kps2d), _p(flags), _p(keep)))
    return keep


def host_merge_keypoints(width, height, grid_width, grid_height, old_kps2d, new_kps2d):
    old, new = _f32(old_kps2d, (-1, 2)), _f32(new_kps2d, (-1, 2))
    n = C.c_int()
    _host_ck(lib().svo_host_merge_keypoints(width, height, grid_width, grid_height, len(old), _p(old), len(new), _p(new), 0, None, C.byref(n)))
    app = np.zeros(max(1, n.value), np.int32)
    _host_ck(lib().svo_host_merge_keypoints(width, height, grid_width, grid_height, len(old), _p(old), len(new), _p(new), len(app), _p(app),
                                            C.byref(n)))
    return app[:n.value].copy()


def host_keyframe_needed(width, height, grid_width, grid_height, kps2d, flags):
    kps2d = _f32(kps2d, (-1, 2))
    flags = np.ascontiguousarray(flags, dtype=np.uint8)
    out = C.c_int()
    _host_ck(lib().svo_host_keyframe_needed(width, height, grid_width, grid_height, len(flags), _p(kps2d), _p(flags), C.byref(out)))
    return bool(out.value)


class MotionFilter:
    """The 12-state cv::KalmanFilter of StereoSlam and StereoSlam::update_pose (stereo_slam.cpp:29-41, :296-359)."""

    def __init__(self):
        self._h = C.c_void_p()
        _host_ck(lib().svo_motion_filter_create(C.byref(self._h)))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().svo_motion_filter_destroy(self._h)
            self._h = None

    def update(self, pose, speed, pose_variance, speed_variance, dt):
        """-> (filtered pose (6), kf.statePre (12))"""
        p = Pose(*[float(v) for v in pose])
        a = [_f32(x) for x in (speed, pose_variance, speed_variance)]
        out, pre = Pose(), np.zeros(12, np.float32)
        _host_ck(lib().svo_motion_filter_update(self._h, C.byref(p), _p(a[0]), _p(a[1]), _p(a[2]), C.c_double(dt), C.byref(out), _p(pre)))
        return out.vec(), pre
