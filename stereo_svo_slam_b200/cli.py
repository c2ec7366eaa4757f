"""Headless runner — the reference's `slam_app` minus the Qt GUI.

    python -m stereo_svo_slam_b200.cli --settings Blender.yaml --video blender-classroom.mkv --trajectory out.csv
    python -m stereo_svo_slam_b200.cli --settings EuRoC.yaml --pairs /data/mav0 --trajectory out.csv
    python -m stereo_svo_slam_b200.cli --settings EuRoC.yaml --euroc /data/MH_02_easy/mav0/ --trajectory out.csv
    python -m stereo_svo_slam_b200.cli --synthetic C3 --frames 100 --trajectory out.csv

* settings: the reference's OpenCV-YAML keys (src/app/image_input.cpp:13-37; examples src/app/Blender.yaml, EuRoC.yaml, Econ.yaml)
* --video: side-by-side stereo video, right image = left half, left image = right half, BGR -> gray, time += 1/fps
  (src/app/video_input.cpp:23-42); decoding needs cv2
* --pairs: a directory with left/ and right/ (or EuRoC's cam0/data and cam1/data) holding equally named images; images are
  used as they are (already rectified)
* --euroc: the reference's EurocInput (src/app/euroc_input.cpp): file list and timestamps from cam0/data.csv, cam0 -> `right`,
  cam1 -> `left`, raw images rectified ON THE DEVICE with the LEFT.* / RIGHT.* matrices of the settings file
  (initUndistortRectifyMap + remap, fused in front of the pyramid kernels)
* --websocket PORT: the reference app's WebSocket JSON backend (src/app/svo_slam_backend.cpp) for src/qt-viewer
* trajectory CSV: `time,x,y,z,rx,ry,rz`, column 0 = cumulative seconds spent inside new_image, angles re-ordered for
  Blender exactly as src/app/slam_app.cpp:220-246 does (Rodrigues((Ry*Rx)*Rz))
"""
import argparse
import os
import re
import sys
import time

import numpy as np

from . import synth
from .capi import CameraSettings

_KEYS = {"Camera1.fx": "fx", "Camera1.fy": "fy", "Camera1.cx": "cx", "Camera1.cy": "cy", "Camera.baseline": "baseline",
         "Camera.window_size_pose_estimator": "window_size_pose_estimator", "Camera.window_size_opt_flow": "window_size_opt_flow",
         "Camera.window_size_depth_calculator": "window_size_depth_calculator", "Camera.max_pyramid_levels": "max_pyramid_levels",
         "Camera.min_pyramid_level_pose_estimation": "min_pyramid_level_pose_estimation", "Camera1.k1": "k1", "Camera1.k2": "k2",
         "Camera1.k3": "k3", "Camera1.p1": "p1", "Camera1.p2": "p2", "Camera.grid_width": "grid_width",
         "Camera.grid_height": "grid_height", "Camera.search_x": "search_x", "Camera.search_y": "search_y"}
_INT_FIELDS = {n for n, t in CameraSettings._fields_ if t.__name__ == "c_int"}


def read_settings(path):
    """ImageInput::read_settings (src/app/image_input.cpp:13-37): scalar `key: value` entries of an OpenCV FileStorage YAML.
    Missing keys read as 0, exactly like cv::FileStorage's operator[] on an absent node."""
    vals = {}
    for line in open(path):
        line = line.split("#", 1)[0].strip()
        m = re.match(r"^([A-Za-z0-9_.]+)\s*:\s*([-+0-9.eE]+)\s*$", line)
        if m and m.group(1) in _KEYS:
            vals[_KEYS[m.group(1)]] = float(m.group(2))
    out = {}
    for name, _ in CameraSettings._fields_:
        v = vals.get(name, 0.0)
        out[name] = int(v) if name in _INT_FIELDS else float(v)
    return out


def read_rectification(path):
    """LEFT.* / RIGHT.* `!!opencv-matrix` nodes of the settings file (src/app/euroc_input.cpp:24-46).
    Returns {"LEFT": dict(K, D, R, P, width, height), "RIGHT": ...}; raises if a node is missing (the reference only
    prints an error and then fails inside OpenCV)."""
    text = open(path).read()
    out = {}
    for side in ("LEFT", "RIGHT"):
        ent = {}
        for key, n in (("K", 9), ("D", 5), ("R", 9), ("P", 12)):
            m = re.search(rf"^{side}\.{key}\s*:\s*!!opencv-matrix(.*?)data\s*:\s*\[(.*?)\]", text, re.S | re.M)
            if not m:
                raise ValueError(f"{path}: calibration parameter {side}.{key} is missing")
            vals = [float(v) for v in m.group(2).replace("\n", " ").split(",") if v.strip()]
            if len(vals) < n:
                raise ValueError(f"{path}: {side}.{key} has {len(vals)} values, expected {n}")
            ent[key] = np.array(vals[:n], np.float64)
        for key in ("width", "height"):
            m = re.search(rf"^{side}\.{key}\s*:\s*([0-9]+)", text, re.M)
            if not m:
                raise ValueError(f"{path}: {side}.{key} is missing")
            ent[key] = int(m.group(1))
        ent["K"], ent["R"], ent["P"] = ent["K"].reshape(3, 3), ent["R"].reshape(3, 3), ent["P"].reshape(3, 4)
        out[side] = ent
    return out


def _rodrigues(r):
    return synth._rodrigues(r)


def _rodrigues_inv(R):
    """cv::Rodrigues(matrix -> vector) for proper rotations."""
    c = (np.trace(R) - 1) / 2
    c = min(1.0, max(-1.0, c))
    th = np.arccos(c)
    v = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    s = np.linalg.norm(v) / 2
    if s < 1e-12:
        if c > 0:
            return np.zeros(3)
        # theta = pi: axis from the diagonal
        d = np.sqrt(np.maximum((np.diag(R) + 1) / 2, 0))
        if R[0, 1] < 0:
            d[1] = -d[1]
        if R[0, 2] < 0:
            d[2] = -d[2]
        return d / max(np.linalg.norm(d), 1e-12) * th
    return v / (2 * s) * th


def blender_angles(pose):
    """slam_app.cpp:232-240: Rodrigues((Ry*Rx)*Rz) -> rotation vector."""
    Rx, Ry, Rz = _rodrigues([pose[3], 0, 0]), _rodrigues([0, pose[4], 0]), _rodrigues([0, 0, pose[5]])
    return _rodrigues_inv((Ry @ Rx) @ Rz)


def write_trajectory_csv(path, trajectory, time_stamps):
    """slam_app.cpp:220-246 — one row per frame: cumulative algorithm seconds, position, Blender-ordered angles."""
    with open(path, "w") as f:
        for t, pose in zip(time_stamps, trajectory):
            a = blender_angles(pose)
            f.write(",".join(f"{v:.6g}" for v in (t, pose[0], pose[1], pose[2], a[0], a[1], a[2])) + "\n")


def _imread_gray(path):
    if path.endswith(".npy"):
        return np.ascontiguousarray(np.load(path), dtype=np.uint8)
    import cv2
    im = cv2.imread(path, cv2.IMREAD_GRAYSCALE)
    if im is None:
        raise IOError(f"cannot read {path}")
    return im


def iter_pairs(root):
    for l, r in (("left", "right"), ("cam0/data", "cam1/data"), ("mav0/cam0/data", "mav0/cam1/data")):
        dl, dr = os.path.join(root, l), os.path.join(root, r)
        if os.path.isdir(dl) and os.path.isdir(dr):
            break
    else:
        raise IOError(f"{root}: expected left/ + right/ or cam0/data + cam1/data")
    names = sorted(set(os.listdir(dl)) & set(os.listdir(dr)))
    for k, nm in enumerate(names):
        stem = os.path.splitext(nm)[0]
        ts = float(stem) * 1e-9 if stem.isdigit() and len(stem) > 12 else k / 20.0   # EuRoC names are nanoseconds
        yield _imread_gray(os.path.join(dl, nm)), _imread_gray(os.path.join(dr, nm)), ts


def iter_euroc(root):
    """EurocInput::load_images / read (euroc_input.cpp:60-78, :87-116): cam0 is fed as `right`, cam1 as `left`; time stamps
    are seconds since the first row of cam0/data.csv, as float.  Images are RAW: rectification happens on the device."""
    rows = []
    with open(os.path.join(root, "cam0", "data.csv")) as f:
        for s in f:
            s = s.strip().replace("\r", "")
            if not s or s[0] == "#":
                continue
            rows.append((float(s.split(",")[0]) / 1.0e9, s.rsplit(",", 1)[-1]))
    t0 = rows[0][0] if rows else 0.0
    for t, name in rows:
        right = _imread_gray(os.path.join(root, "cam0", "data", name))
        left = _imread_gray(os.path.join(root, "cam1", "data", name))
        yield left, right, float(np.float32(t - t0))


def iter_video(path):
    import cv2
    cap = cv2.VideoCapture(path)
    if not cap.isOpened():
        raise IOError(f"cannot open {path}")
    fps = cap.get(cv2.CAP_PROP_FPS) or 20.0
    ts = 0.0
    while True:
        ok, img = cap.read()
        if not ok:
            return
        if img.ndim == 3:
            img = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)
        w = img.shape[1] // 2
        right, left = img[:, :w], img[:, w:2 * w]       # video_input.cpp:35-36 (strided views, handled by the C-ABI)
        ts += 1.0 / fps
        yield left, right, ts


def iter_synthetic(cfg, frames):
    seq = synth.make_sequence(cfg)
    for k in range(frames):
        left, right = seq.render(k)
        yield left, right, k / 20.0


def run(frames, settings, device=0, trajectory=None, verbose=False, rectification=None, websocket_port=None, websocket_host="127.0.0.1"):
    from .slam import StereoSlam
    slam, t_algo, stamps, server = None, 0.0, [], None
    for left, right, ts in frames:
        if slam is None:
            slam = StereoSlam(settings if isinstance(settings, CameraSettings) else CameraSettings(**settings), left.shape[1], left.shape[0],
                              device=device)
            if rectification:
                # euroc_input.cpp:69-70: the cam0 image (`right`) goes through the LEFT.* maps, cam1 (`left`) through RIGHT.*
                for which, side in ((0, "RIGHT"), (1, "LEFT")):
                    c = rectification[side]
                    slam.set_rectification(which, c["K"], c["D"], c["R"], c["P"])
            if websocket_port is not None:
                from .backend import WebSocketServer     # the reference app's server (src/app/main.cpp:208), for qt-viewer
                server = WebSocketServer(slam, websocket_port, websocket_host)
        t0 = time.perf_counter()
        slam.new_image(left, right, ts)
        t_algo += time.perf_counter() - t0           # the reference's TickMeter brackets exactly new_image (slam_app.cpp:187-190)
        stamps.append(t_algo)
        if server:
            server.serve_pending()                   # same thread as new_image, like the reference's Qt event loop
        if verbose:
            print(f"frame {len(stamps) - 1}: pose {slam.pose()} keypoints {len(slam.get_frame().kps)} keyframes {slam.keyframe_count()}")
    if server:
        server.close()
    traj = slam.get_trajectory() if slam else np.zeros((0, 6), np.float32)
    if trajectory:
        write_trajectory_csv(trajectory, traj, stamps)
    return traj, stamps, slam


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--settings", "-s", help="camera settings YAML (reference format)")
    src = ap.add_mutually_exclusive_group(required=True)
    src.add_argument("--video", "-v")
    src.add_argument("--pairs")
    src.add_argument("--euroc", help="EuRoC mav0/ directory with raw cam0/ and cam1/ (rectified on the device)")
    src.add_argument("--synthetic", choices=sorted(synth.CONFIGS))
    ap.add_argument("--frames", type=int, default=50, help="frames of the synthetic sequence")
    ap.add_argument("--trajectory", "-t", help="trajectory CSV to write")
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--websocket", type=int, metavar="PORT", help="serve the reference's /keyframes, /pose, /trajectory JSON (qt-viewer uses 8001)")
    ap.add_argument("--websocket-host", default="127.0.0.1", help="interface of the WebSocket server (the reference listens on all: 0.0.0.0)")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args(argv)
    if a.synthetic:
        settings = synth.settings_dict(a.synthetic)
        frames = iter_synthetic(a.synthetic, a.frames)
    else:
        if not a.settings:
            ap.error("--settings is required with --video / --pairs / --euroc")
        settings = read_settings(a.settings)
        frames = iter_video(a.video) if a.video else iter_euroc(a.euroc) if a.euroc else iter_pairs(a.pairs)
    rect = read_rectification(a.settings) if a.euroc else None
    traj, stamps, _ = run(frames, settings, a.device, a.trajectory, a.verbose, rect, a.websocket, a.websocket_host)
    if stamps:
        print(f"{len(stamps)} frames, {len(stamps) / stamps[-1]:.1f} frames/s (algorithm time, test/extract_fps.py definition)")
    return 0


if __name__ == "__main__":
    sys.exit(main())
