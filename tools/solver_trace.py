"""Phase trace of the two Gauss-Newton solver kernels (developer tool): SVO_SOLVER_TRACE=1 python tools/solver_trace.py [cfg] [frame]

Prints, for one tracking frame, the SM-clock stamps thread 0 of CTA 0 took at the phase boundaries of sparse_align_kernel and
reproj_refine_kernel (svo_debug_solver_trace): level setup, every cost round (with the number of trial poses in it) and every
gradient round, in microseconds at the SM clock nvidia-smi reports."""
import ctypes as C
import os
import subprocess
import sys
import numpy as np
os.environ.setdefault("SVO_SOLVER_TRACE", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stereo_svo_slam_b200 import StereoSlam, capi, synth  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "C3"
frame = int(sys.argv[2]) if len(sys.argv) > 2 else 20
c, d = synth.CONFIGS[cfg], synth.settings_dict(cfg)
seq = synth.make_sequence(cfg)
g = StereoSlam(capi.CameraSettings(**d), c["width"], c["height"])
for k in range(frame + 1):
    L, R = seq.render(k)
    g.new_image(L, R, k / 20.0)
try:
    mhz = float(subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm", "--format=csv,noheader,nounits"], capture_output=True, text=True).stdout.split()[0])
except Exception:  # noqa: BLE001
    mhz = 1965.0
lib = capi.lib()
ctx = C.c_void_p(lib.svo_slam_ctx(g._h))
names = {0: "start", 1: "level", 2: "images staged", 3: "reference terms", 4: "cost round", 5: "gradient round", 6: "end", 7: "  grad: keypoints done", 8: "  grad: sums exchanged", 9: "  grad: 6x6 solved", 10: "  grad: expmap + rotation", 11: "  grad: Rodrigues of the trial pose"}
for which, kname in ((0, "sparse_align_kernel"), (1, "reproj_refine_kernel"), (2, "depth_filter_kernel")):
    if which == 2:
        names = {0: "start", 1: "matrices ready", 2: "own keypoint updated", 3: "all keypoints done", 5: "export copies issued", 6: "export copies done in the CTA",
                 7: "system fence passed", 8: "end"}
    buf = (C.c_ulonglong * 1024)()
    lib.svo_debug_solver_trace(ctx, which, buf, 1024)
    n = int(buf[0])
    ent = [(int(buf[1 + k]) >> 16, (int(buf[1 + k]) >> 8) & 255, int(buf[1 + k]) & 255) for k in range(n)]
    print(f"== {kname}: {n} stamps, SM clock {mhz:.0f} MHz, env", {k: v for k, v in os.environ.items() if k.startswith('SVO_')})
    if not ent:
        continue
    t0, prev = ent[0][0], ent[0][0]
    tot = {}
    for t, extra, tag in ent:
        dt = (t - prev) / mhz
        tot.setdefault(tag, []).append(dt)
        print(f"  {(t - t0) / mhz:8.2f} us  +{dt:6.2f}  {names.get(tag, tag)}" + (f" ({extra})" if tag in (1, 2, 3, 4) else ""))
        prev = t
    print("  summary:", {names[k]: (len(v), round(float(np.sum(v)), 2)) for k, v in tot.items()})
g.close()
