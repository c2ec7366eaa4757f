"""Python mirror of the reference's Python wrapper (src/python/wrapper/slam_accelerator.pyx:50-91):
StereoSlam(camera_settings).new_image(left, right, time_stamp) / get_frame() / get_keyframe() /
get_keyframes() / get_trajectory() / update_pose(...), backed by the svo_slam_* C-ABI (CUDA, sm_100a).
"""
import ctypes as C

import numpy as np

from . import capi
from .capi import CameraSettings, KeyPointInfo, Pose, SvoError  # noqa: F401  (re-exported, same names as the reference)

_KP_DTYPE = np.dtype(KeyPointInfo)


class KeyPoints:
    """== struct KeyPoints (stereo_slam_types.hpp:108-112): kps2d (n,2), kps3d (n,3), info (structured array)."""

    def __init__(self, kps2d, kps3d, info):
        self.kps2d, self.kps3d, self.info = kps2d, kps3d, info

    def __len__(self):
        return len(self.info)


class Frame:
    """== struct Frame (stereo_slam_types.hpp:119-125). stereo_image levels are fetched lazily from the device."""

    def __init__(self, slam, index, fid, pose, time_stamp, kps):
        self._slam, self._index = slam, index
        self.id, self.pose, self.time_stamp, self.kps = fid, pose, time_stamp, kps

    def image(self, kind="left", level=0):
        """Pixels of THIS frame (the reference's getters return value copies): a keyframe is addressed by its absolute index;
        the current frame's images are only on the device until the next new_image, after which this raises."""
        k = {"left": 0, "right": 1, "opt_flow": 2}[kind]
        return self._slam._image(self._index, k, level, self.id)


class KeyFrame(Frame):
    pass


class StereoSlam:
    def __init__(self, camera_settings, width, height, device=0, max_keypoints=0):
        if isinstance(camera_settings, dict):
            camera_settings = CameraSettings(**camera_settings)
        self.camera_settings, self.width, self.height = camera_settings, width, height
        out = C.c_void_p()
        rc = capi.lib().svo_slam_create_with_capacity(C.byref(camera_settings), device, width, height, max_keypoints, C.byref(out))
        if rc:
            raise SvoError(rc, capi.lib().svo_slam_last_error(None).decode())
        self._h = out
        self._keep = None

    def close(self):
        if getattr(self, "_h", None):
            capi.lib().svo_slam_destroy(self._h)
            self._h = None

    def reset(self):
        """A fresh StereoSlam on the same camera that keeps every device resource (svo_slam_reset): the next image is a first image."""
        self._ck(capi.lib().svo_slam_reset(self._h))

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc:
            raise SvoError(rc, capi.lib().svo_slam_last_error(self._h).decode())

    @staticmethod
    def _img(a):
        if a.dtype != np.uint8 or a.ndim != 2:
            raise SvoError(capi.SVO_ERR_INVALID, "images must be 2-D uint8 (CV_8U single channel, stereo_slam.cpp:96)")
        if a.strides[1] != 1:
            a = np.ascontiguousarray(a)
        return a

    def set_rectification(self, which, K, D, R, P):
        """EurocInput's initUndistortRectifyMap + remap (src/app/euroc_input.cpp:48-49, :69-73) moved onto the device:
        from now on the image passed as left (which=0) / right (which=1) is a raw camera image."""
        a = [np.ascontiguousarray(np.asarray(x, np.float64).reshape(-1)) for x in (K, D, R, np.asarray(P, np.float64).reshape(3, -1)[:, :3])]
        if (a[0].size, a[1].size, a[2].size, a[3].size) != (9, 5, 9, 9):
            raise SvoError(capi.SVO_ERR_INVALID, "K, R: 3x3; D: 5 coefficients; P: 3x3 or 3x4")
        self._ck(capi.lib().svo_slam_set_rectification(self._h, which, *[x.ctypes.data_as(C.c_void_p) for x in a]))

    def set_align_cluster(self, ctas):
        """Latency/throughput knob of the alignment kernel (svo_set_align_cluster): 8 SMs per frame for one live sequence,
        1-2 when many sequences share the GPU."""
        self.context().set_align_cluster(ctas)

    def set_solver_width(self, wide):
        """1: wide line searches (several trial poses per round on idle SMs: lowest latency of one sequence), 0: sequential
        (fewest SMs per frame), -1: automatic (svo_set_solver_width).  Same results bit for bit."""
        self.context().set_solver_width(wide)

    # ---- StereoSlam::new_image (stereo_slam.cpp:123)
    def new_image(self, left, right, time_stamp):
        left, right = self._img(left), self._img(right)
        if left.shape != (self.height, self.width) or right.shape != left.shape:
            raise SvoError(capi.SVO_ERR_INVALID, "image size differs from the size the instance was created with")
        self._ck(capi.lib().svo_slam_new_image(self._h, left.ctypes.data_as(C.c_void_p), C.c_size_t(left.strides[0]),
                                               right.ctypes.data_as(C.c_void_p), C.c_size_t(right.strides[0]),
                                               C.c_float(time_stamp)))

    # pipelined form for many independent sequences (one CUDA stream per instance)
    def new_image_begin(self, left, right, time_stamp):
        left, right = self._img(left), self._img(right)
        self._keep = (left, right)
        self._ck(capi.lib().svo_slam_new_image_begin(self._h, left.ctypes.data_as(C.c_void_p), C.c_size_t(left.strides[0]),
                                                     right.ctypes.data_as(C.c_void_p), C.c_size_t(right.strides[0]),
                                                     C.c_float(time_stamp)))

    def new_image_end(self):
        self._ck(capi.lib().svo_slam_new_image_end(self._h))
        self._keep = None

    def _image(self, index, kind, level, frame_id=None):
        ctx = capi.Context(self.camera_settings, self.width, self.height, _borrowed=capi.lib().svo_slam_ctx(self._h))
        w, h = ctx.level_size(kind, level)
        out = np.empty((h, w), np.uint8)
        if index is None:
            fid = C.c_uint64()
            self._ck(capi.lib().svo_slam_get_frame(self._h, C.byref(fid), None, None, None))
            if frame_id is not None and fid.value != frame_id:
                raise SvoError(capi.SVO_ERR_STATE, f"images of frame {frame_id} are gone: the sequence is at frame {fid.value}")
            self._ck(capi.lib().svo_slam_get_frame_image(self._h, kind, level, out.ctypes.data_as(C.c_void_p), C.c_size_t(w)))
        else:
            self._ck(capi.lib().svo_slam_get_keyframe_image(self._h, index, kind, level, out.ctypes.data_as(C.c_void_p), C.c_size_t(w)))
        return out

    def _fetch(self, index):
        fid, pose, ts, n = C.c_uint64(), Pose(), C.c_double(), C.c_int()
        if index is not None and index < 0:   # "latest keyframe": pin the absolute index, later keyframes must not change what this object shows
            index = capi.lib().svo_slam_keyframe_count(self._h) - 1
            if index < 0:
                return None
        if index is None:
            rc = capi.lib().svo_slam_get_frame(self._h, C.byref(fid), C.byref(pose), C.byref(ts), C.byref(n))
        else:
            rc = capi.lib().svo_slam_get_keyframe(self._h, index, C.byref(fid), C.byref(pose), C.byref(ts), C.byref(n))
        if rc == capi.SVO_ERR_STATE:
            return None
        self._ck(rc)
        k2, k3 = np.empty((n.value, 2), np.float32), np.empty((n.value, 3), np.float32)
        info = np.zeros(n.value, _KP_DTYPE)
        args = (n.value, k2.ctypes.data_as(C.c_void_p), k3.ctypes.data_as(C.c_void_p), info.ctypes.data_as(C.c_void_p))
        if index is None:
            self._ck(capi.lib().svo_slam_get_frame_keypoints(self._h, *args))
        else:
            self._ck(capi.lib().svo_slam_get_keyframe_keypoints(self._h, index, *args))
        cls = Frame if index is None else KeyFrame
        return cls(self, index, fid.value, pose.vec(), ts.value, KeyPoints(k2, k3, info))

    def get_frame(self):
        """StereoSlam::get_frame (stereo_slam.cpp:278-284): None before the first image."""
        return self._fetch(None)

    def get_keyframe(self):
        return self._fetch(-1)

    def get_keyframes(self):
        return [self._fetch(i) for i in range(capi.lib().svo_slam_keyframe_count(self._h))]

    def keyframe_count(self):
        return capi.lib().svo_slam_keyframe_count(self._h)

    def get_trajectory(self):
        n = capi.lib().svo_slam_get_trajectory(self._h, 0, None)
        out = np.empty((n, 6), np.float32)
        if n:
            capi.lib().svo_slam_get_trajectory(self._h, n, out.ctypes.data_as(C.c_void_p))
        return out

    def pose(self):
        fid, pose, ts, n = C.c_uint64(), Pose(), C.c_double(), C.c_int()
        self._ck(capi.lib().svo_slam_get_frame(self._h, C.byref(fid), C.byref(pose), C.byref(ts), C.byref(n)))
        return pose.vec()

    def update_pose(self, pose, speed, pose_variance, speed_variance, dt):
        p = Pose(*[float(v) for v in pose])
        arrs = [np.ascontiguousarray(a, dtype=np.float32) for a in (speed, pose_variance, speed_variance)]
        out = Pose()
        self._ck(capi.lib().svo_slam_update_pose(self._h, C.byref(p), *[a.ctypes.data_as(C.c_void_p) for a in arrs],
                                                 C.c_double(dt), C.byref(out)))
        return out.vec()

    def last_stats(self):
        ms, n, kf = C.c_float(), C.c_int(), C.c_int()
        self._ck(capi.lib().svo_slam_last_stats(self._h, C.byref(ms), C.byref(n), C.byref(kf)))
        return dict(gpu_ms=ms.value, launches=n.value, keyframe_created=bool(kf.value))

    def dropped_keypoints(self):
        """New keypoints keyframes could not take because the device keypoint block was full (0 unless max_keypoints is tiny)."""
        out = C.c_longlong()
        self._ck(capi.lib().svo_slam_dropped_keypoints(self._h, C.byref(out)))
        return int(out.value)

    def total_counters(self):
        out = (C.c_longlong * 8)()
        self._ck(capi.lib().svo_slam_total_counters(self._h, out))
        names = ("frames", "tracking_frames", "keyframes", "align_patches", "klt_windows", "keypoints", "align_evaluations", "refine_evaluations")
        return dict(zip(names, [int(v) for v in out]))

    def last_counters(self):
        out = (C.c_longlong * 8)()
        self._ck(capi.lib().svo_slam_last_counters(self._h, out))
        names = ("keypoints", "align_keypoints", "align_cost_evals", "align_grad_evals", "refine_cost_evals", "refine_grad_evals",
                 "klt_iterations", "klt_tracked")
        return dict(zip(names, [int(v) for v in out]))

    def context(self):
        """The device context behind the facade (stage-level probes)."""
        return capi.Context(self.camera_settings, self.width, self.height, _borrowed=capi.lib().svo_slam_ctx(self._h))


def run_many(slams, left, right, time_stamps, workers=4, on_device=False, strides=None, restart=None):
    """svo_slam_run_many: advance len(slams) independent sequences by F frames each in one native call.

    left / right: per sequence a list of F frames — numpy uint8 arrays (H, W) with a common row stride, or, with
    on_device=True, integer device addresses (row stride = strides or the image width); time_stamps: [n][F] floats.
    `workers` native host threads share the sequences; no Python runs between the frames.  restart: optional [n][F] booleans —
    True makes that frame the first image of a new sequence of the stream (svo_slam_reset before it)."""
    n = len(slams)
    if n == 0:
        return
    F = len(left[0])
    lp, rp = (C.c_void_p * (n * F))(), (C.c_void_p * (n * F))()
    ts = np.empty(n * F, np.float32)
    keep = []
    stride = strides
    for i in range(n):
        if len(left[i]) != F or len(right[i]) != F or len(time_stamps[i]) != F:
            raise SvoError(capi.SVO_ERR_INVALID, "every sequence needs the same number of frames")
        for f in range(F):
            a, b = left[i][f], right[i][f]
            if on_device:
                lp[i * F + f], rp[i * F + f] = int(a), int(b)
            else:
                a, b = StereoSlam._img(a), StereoSlam._img(b)
                if stride is None:
                    stride = a.strides[0]
                if a.strides[0] != stride or b.strides[0] != stride:
                    raise SvoError(capi.SVO_ERR_INVALID, "all frames of one run_many call must share one row stride")
                keep.append((a, b))
                lp[i * F + f], rp[i * F + f] = a.ctypes.data, b.ctypes.data
            ts[i * F + f] = time_stamps[i][f]
    if stride is None:
        stride = slams[0].width
    handles = (C.c_void_p * n)(*[s._h for s in slams])
    bad = C.c_int(-1)
    rs = None if restart is None else np.ascontiguousarray(np.asarray(restart, dtype=np.uint8).reshape(n * F))
    rc = capi.lib().svo_slam_run_many_restart(handles, n, F, lp, rp, C.c_size_t(stride), C.c_size_t(stride), ts.ctypes.data_as(C.c_void_p),
                                              None if rs is None else rs.ctypes.data_as(C.c_void_p), 1 if on_device else 0, int(workers),
                                              C.byref(bad))
    if rc:
        who = slams[bad.value] if 0 <= bad.value < n else slams[0]
        raise SvoError(rc, f"sequence {bad.value}: " + capi.lib().svo_slam_last_error(who._h).decode())
