"""Per-stage device times of ONE sequence (developer tool; the A/B harness of kernel changes: run it under different
SVO_* environment switches in one gpurun call).

    python tools/stage_times.py [cfg=C3] [frames=60] [tag]

Prints one JSON line: graph-replay device time per frame (median over tracking frames) and the per-stage CUDA-event
medians of the same sequence launched kernel by kernel (svo_set_profiling), plus the solver's evaluation counts.
"""
import ctypes as C
import json
import os
import sys
import time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stereo_svo_slam_b200 import StereoSlam, capi, synth  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "C3"
nf = int(sys.argv[2]) if len(sys.argv) > 2 else 60
tag = sys.argv[3] if len(sys.argv) > 3 else ""
c, d = synth.CONFIGS[cfg], synth.settings_dict(cfg)
seq = synth.make_sequence(cfg)
imgs = [seq.render(k) for k in range(nf)]
lib = capi.lib()
names = ["upload+pyramids", "sparse_align", "klt", "reproj_refine", "stereo_ssd", "depth_filter", "d2h", "total"]


def run(profiling):
    g = StereoSlam(capi.CameraSettings(**d), c["width"], c["height"])
    ctx = C.c_void_p(lib.svo_slam_ctx(g._h))
    if profiling:
        lib.svo_set_profiling(ctx, 1)
    buf = (C.c_float * 8)()
    gpu, wall, stage, cnt = [], [], [], []
    for k in range(nf):
        t0 = time.perf_counter()
        g.new_image(imgs[k][0], imgs[k][1], k / 20.0)
        w = time.perf_counter() - t0
        st = g.last_stats()
        if k >= 5 and not st["keyframe_created"]:
            gpu.append(st["gpu_ms"]); wall.append(w)
            if profiling:
                lib.svo_last_stage_ms(ctx, buf)
                stage.append(list(buf))
            cnt.append(list(g.last_counters().values()))
    pose = g.pose()
    g.close()
    return np.array(gpu), np.array(wall), np.array(stage), np.array(cnt, dtype=np.float64), pose


gpu, wall, _, cnt, pose = run(False)
_, _, stage, _, pose2 = run(True)
st = np.median(stage, axis=0)
cm = cnt.mean(axis=0)
out = {"tag": tag, "cfg": cfg, "frames": nf, "gpu_ms_graph": float(np.median(gpu)), "wall_ms_graph": float(np.median(wall)),
       "stage_us": {n: round(float(st[i]) * 1e3, 2) for i, n in enumerate(names)},
       "keypoints": float(cm[0]), "counters_mean": [round(float(x), 2) for x in cm],
       "pose_err": float(np.abs(pose - seq.pose(nf - 1)).max()), "same_pose_both_modes": bool((pose == pose2).all()),
       "env": {k: v for k, v in os.environ.items() if k.startswith("SVO_")}}
print(json.dumps(out), flush=True)
