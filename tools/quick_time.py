"""Quick single-stream timing of the facade on a synthetic config (developer tool)."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stereo_svo_slam_b200 import StereoSlam, capi, synth

cfg = sys.argv[1] if len(sys.argv) > 1 else "C3"
nf = int(sys.argv[2]) if len(sys.argv) > 2 else 40
c = synth.CONFIGS[cfg]
seq = synth.make_sequence(cfg)
frames = [seq.render(k) for k in range(nf)]
g = StereoSlam(capi.CameraSettings(**synth.settings_dict(cfg)), c["width"], c["height"])
ts, gms = [], []
for k, (L, R) in enumerate(frames):
    t0 = time.perf_counter()
    g.new_image(L, R, k / 20.0)
    ts.append(time.perf_counter() - t0)
    st = g.last_stats()
    gms.append(st["gpu_ms"])
    if k < 3 or st["keyframe_created"]:
        print(k, "wall ms %.3f gpu ms %.3f launches %d kf %s n %d" % (ts[-1] * 1e3, st["gpu_ms"], st["launches"], st["keyframe_created"], len(g.get_frame().kps)))
ts, gms = np.array(ts[5:]), np.array(gms[5:])
print("median wall ms %.3f  mean %.3f  median gpu ms %.3f  fps(median) %.1f" % (np.median(ts) * 1e3, ts.mean() * 1e3, np.median(gms), 1 / np.median(ts)))
gt = seq.pose(nf - 1)
print("final pose err", np.abs(g.pose() - gt).max())
