"""Developer tool: host cost of keyframe frames vs tracking frames (run with SVO_TRACE_KF=1 for the stage split)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from stereo_svo_slam_b200 import StereoSlam, capi, synth

cfg = "C3"
c = synth.CONFIGS[cfg]
seq = synth.make_sequence(cfg, seed=1000)
nf = 64
frames = [seq.render(k) for k in range(nf)]
pinned = [(torch.from_numpy(L).pin_memory().numpy(), torch.from_numpy(R).pin_memory().numpy()) for L, R in frames]
g = StereoSlam(capi.CameraSettings(**synth.settings_dict(cfg)), c["width"], c["height"])
def tri(t):
    p = t % (2 * nf - 2)
    return p if p < nf else 2 * nf - 2 - p
for k in range(300):
    L, R = pinned[tri(k)]
    t0 = time.perf_counter()
    g.new_image(L, R, k / 20.0)
    dt = time.perf_counter() - t0
    st = g.last_stats()
    if st["keyframe_created"] or dt > 1e-3:
        print(k, "wall ms %.3f gpu ms %.3f kf %s n %d" % (dt * 1e3, st["gpu_ms"], st["keyframe_created"], len(g.get_frame().kps)), flush=True)
# pinned H2D bandwidth (one stream, 8 MB transfers)
h = torch.empty(64 << 20, dtype=torch.uint8).pin_memory()
d = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
for sz in (361 << 10, 8 << 20, 64 << 20):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = max(4, (256 << 20) // sz)
    e0.record()
    for _ in range(reps):
        d[:sz].copy_(h[:sz], non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    print("H2D %8d B x %d: %.1f GB/s" % (sz, reps, sz * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9))
