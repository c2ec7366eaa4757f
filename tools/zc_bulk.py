"""Developer tool: zero-copy read bandwidth from page-locked host memory — SM loads (16 B per thread) vs TMA bulk copies
(SVO_ZC_BULK=<chunk bytes>, set before the first call).  python tools/zc_bulk.py [chunk]"""
import ctypes as C, os, sys
if len(sys.argv) > 1:
    os.environ["SVO_ZC_BULK"] = sys.argv[1]
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from stereo_svo_slam_b200 import capi, synth
ctx = capi.Context(capi.CameraSettings(**synth.settings_dict("C3")), 752, 480)
h = torch.empty(64 << 20, dtype=torch.uint8).pin_memory()
h[:1 << 20].copy_(torch.arange(1 << 20, dtype=torch.int64).to(torch.uint8))
out = C.c_float()
for nbytes in (722 << 10, 64 << 20):
    for ctas in (4, 8, 16, 32, 74, 148):
        reps = max(2, (1 << 29) // nbytes)
        rc = capi.lib().svo_debug_zero_copy_bandwidth(ctx.h_ctx, C.c_void_p(h.data_ptr()), C.c_size_t(nbytes), ctas, reps, C.byref(out))
        print(f"{os.environ.get('SVO_ZC_BULK', 'sm-loads'):>8s} {nbytes >> 10:7d} KB, {ctas:5d} CTAs: {out.value:6.1f} GB/s (rc {rc})", flush=True)
