"""Profiling driver (tools/ only): ONE C3 sequence, synchronous new_image calls, kernel-by-kernel launches (no graph replay,
so every kernel of a frame shows up as a launch of its own).  60 frames: keyframe #2 falls on frame 46.

    python tools/profile_frames.py [frames] [cfg]      # under ncu: see profiles/README_r02.md
"""
import os
import sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("SVO_NO_GRAPHS", "1")
from stereo_svo_slam_b200 import StereoSlam, capi, synth  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 60
cfg = sys.argv[2] if len(sys.argv) > 2 else "C3"
c, d = synth.CONFIGS[cfg], synth.settings_dict(cfg)
seq = synth.make_sequence(cfg)
imgs = [seq.render(k) for k in range(frames)]
g = StereoSlam(capi.CameraSettings(**d), c["width"], c["height"])
for k in range(frames):
    g.new_image(imgs[k][0], imgs[k][1], k / 20.0)
    if k in (0, 10, 46, frames - 1):
        print("frame", k, "keypoints", len(g.get_frame().kps), "keyframes", g.keyframe_count(), "launches", g.last_stats()["launches"], flush=True)
print("pose", g.pose(), "true", seq.pose(frames - 1))
