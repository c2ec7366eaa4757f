"""Developer tool: zero-copy (SM-driven) read bandwidth from page-locked host memory vs copy-engine DMA."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from stereo_svo_slam_b200 import capi, synth
ctx = capi.Context(capi.CameraSettings(**synth.settings_dict("C3")), 752, 480)
h = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
out = C.c_float()
for nbytes in (722 << 10, 16 << 20, 256 << 20):
    for ctas in (74, 148, 296, 592, 1184):
        reps = max(2, (1 << 30) // nbytes)
        rc = capi.lib().svo_debug_zero_copy_bandwidth(ctx.h_ctx, C.c_void_p(h.data_ptr()), C.c_size_t(nbytes), ctas, reps, C.byref(out))
        print(f"zero-copy {nbytes >> 10:7d} KB, {ctas:5d} CTAs: {out.value:6.1f} GB/s (rc {rc})", flush=True)
d = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for sz in (722 << 10, 16 << 20, 256 << 20):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = max(2, (1 << 30) // sz)
    e0.record()
    for _ in range(reps):
        d[:sz].copy_(h[:sz], non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    print(f"DMA       {sz >> 10:7d} KB: {sz * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9:6.1f} GB/s")
