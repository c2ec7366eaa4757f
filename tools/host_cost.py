"""Host-side cost of begin/end per frame with S sequences driven from one thread (developer tool)."""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import torch
from stereo_svo_slam_b200 import StereoSlam, capi, synth
S = int(sys.argv[1]) if len(sys.argv) > 1 else 16
F = 40
c = synth.CONFIGS["C3"]
seq = synth.make_sequence("C3")
fr = np.stack([np.stack(seq.render(k)) for k in range(F)])
host = torch.from_numpy(fr).pin_memory()
lib = capi.lib()
settings = capi.CameraSettings(**synth.settings_dict("C3"))
slams = [StereoSlam(settings, 752, 480) for _ in range(S)]
img = 752 * 480
tb, te = [], []
for k in range(F):
    t0 = time.perf_counter()
    for sl in slams:
        off = k * 2 * img
        lib.svo_slam_new_image_begin(sl._h, C.c_void_p(host.data_ptr() + off), C.c_size_t(752), C.c_void_p(host.data_ptr() + off + img), C.c_size_t(752), C.c_float(k / 20))
    t1 = time.perf_counter()
    for sl in slams:
        lib.svo_slam_new_image_end(sl._h)
    t2 = time.perf_counter()
    tb.append((t1 - t0) / S); te.append((t2 - t1) / S)
tb, te = np.array(tb[5:]) * 1e6, np.array(te[5:]) * 1e6
print(f"S={S} begin {np.median(tb):.1f} us/frame  end {np.median(te):.1f} us/frame  -> {1e6/ (np.median(tb)+np.median(te)):.0f} fps single host thread")
