"""Drop-in for the reference's Cython module `slam_accelerator` (src/python/wrapper/slam_accelerator.pyx, SURVEY.md §8f
rank 4): same class names, constructor arguments and attribute-style access, so `src/python/main.py` and `draw_kps.py`
run against this library with

    import stereo_svo_slam_b200.slam_accelerator as slam_accelerator      # or put the package directory on PYTHONPATH

    cs = CameraSettings(); cs.fx = ...; slam = StereoSlam(cs)             # slam_accelerator.pyx:50-57, :135-349
    slam.new_image(left, right, t); f = slam.get_frame(); kf = slam.get_keyframe()
    f.pose.x, f.kps.kps2d[i].x, f.kps.info[i].color['r'], f.kps.info[i].type == KeyPointType.KP_FAST, kf.stereo_image.left[0]

Everything is a thin view over the ctypes facade in slam.py (which calls the CUDA library); nothing is computed here.
"""
import numpy as np

from . import slam as _slam
from .capi import CameraSettings  # noqa: F401  (default-constructible, fields assigned one by one like the Cython class)


class KeyPointType:
    """== enum KeyPointType (stereo_slam_types.hpp:52-55, slam_accelerator.pyx:711-715)"""
    KP_FAST = 0
    KP_EDGELET = 1


class Pose:
    """slam_accelerator.pyx:637-709"""
    __slots__ = ("x", "y", "z", "rx", "ry", "rz")

    def __init__(self, v=(0, 0, 0, 0, 0, 0)):
        self.x, self.y, self.z, self.rx, self.ry, self.rz = (float(a) for a in v)

    def __repr__(self):
        return f"Pose({self.x}, {self.y}, {self.z}, {self.rx}, {self.ry}, {self.rz})"


class KeyPoint2d:
    __slots__ = ("x", "y")

    def __init__(self, x=0.0, y=0.0):
        self.x, self.y = float(x), float(y)


class KeyPoint3d:
    __slots__ = ("x", "y", "z")

    def __init__(self, x=0.0, y=0.0, z=0.0):
        self.x, self.y, self.z = float(x), float(y), float(z)


class KeyPointInformation:
    """slam_accelerator.pyx:518-574 (+ the fields of stereo_slam_types.hpp:86-100 the wrapper leaves out)"""

    def __init__(self, rec):
        self.score, self.level, self.type = float(rec["score"]), int(rec["level"]), int(rec["type"])
        self.keyframe_id, self.keypoint_index = int(rec["keyframe_id"]), int(rec["keypoint_index"])
        c = rec["color"]
        self.color = {"r": int(c[0]), "g": int(c[1]), "b": int(c[2])}
        self.ignore_during_refinement = bool(rec["ignore_during_refinement"])
        self.ignore_completely, self.ignore_temporary = bool(rec["ignore_completely"]), bool(rec["ignore_temporary"])
        self.inlier_count, self.outlier_count = int(rec["inlier_count"]), int(rec["outlier_count"])


class _Seq:
    """list-like view that builds the small objects on access"""

    def __init__(self, arr, make):
        self._a, self._make = arr, make

    def __len__(self):
        return len(self._a)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self._make(r) for r in self._a[i]]
        return self._make(self._a[i])

    def __iter__(self):
        return (self._make(r) for r in self._a)


class KeyPoints:
    """slam_accelerator.pyx:576-635"""

    def __init__(self, kps):
        self.kps2d = _Seq(kps.kps2d, lambda r: KeyPoint2d(r[0], r[1]))
        self.kps3d = _Seq(kps.kps3d, lambda r: KeyPoint3d(r[0], r[1], r[2]))
        self.info = _Seq(kps.info, KeyPointInformation)


class StereoImage:
    """slam_accelerator.pyx:94-133: `left` = the half-sample pyramid levels, `right` = [level 0]; fetched from the device on
    first access."""

    def __init__(self, frame, levels):
        self._f, self._levels, self._left, self._right = frame, levels, None, None

    @property
    def left(self):
        if self._left is None:
            self._left = [self._f.image("left", l) for l in range(self._levels)]
        return self._left

    @property
    def right(self):
        if self._right is None:
            self._right = [self._f.image("right", 0)]
        return self._right


class Frame:
    """slam_accelerator.pyx:351-398"""

    def __init__(self, f, levels):
        self.id, self.pose, self.kps = int(f.id), Pose(f.pose), KeyPoints(f.kps)
        self.stereo_image = StereoImage(f, levels)


class KeyFrame(Frame):
    """slam_accelerator.pyx:400-447"""


class StereoSlam:
    """slam_accelerator.pyx:50-91.  The reference constructor takes only the camera settings; the image size is taken from
    the first stereo pair."""

    def __init__(self, camera_settings, device=0):
        self._cs, self._device, self._impl = camera_settings, device, None

    def new_image(self, left, right, dt):
        left, right = np.asarray(left), np.asarray(right)
        if self._impl is None:
            self._impl = _slam.StereoSlam(self._cs, left.shape[1], left.shape[0], device=self._device)
        self._impl.new_image(left, right, float(dt))

    def _wrap(self, f, cls):
        return None if f is None else cls(f, self._cs.max_pyramid_levels)

    def get_frame(self):
        return self._wrap(self._impl.get_frame() if self._impl else None, Frame)

    def get_keyframe(self):
        return self._wrap(self._impl.get_keyframe() if self._impl else None, KeyFrame)

    def get_keyframes(self):
        return [self._wrap(k, KeyFrame) for k in (self._impl.get_keyframes() if self._impl else [])]

    def get_trajectory(self):
        return [Pose(p) for p in (self._impl.get_trajectory() if self._impl else [])]
