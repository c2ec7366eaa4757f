"""Developer tool: how many stereo pairs per second cross PCIe when 32 contexts on 16 threads do NOTHING but ingest page-locked
frames and build the pyramids (the e2e path minus tracking) — the PCIe ceiling of bench.py's e2e number."""
import ctypes as C, os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import numpy as np
import torch
from stereo_svo_slam_b200 import capi, synth

S, T, K = 32, 16, 300
W, H = 752, 480
host = torch.randint(0, 255, (S, 16, 2, H, W), dtype=torch.uint8).pin_memory()
ctxs = [capi.Context(capi.CameraSettings(**synth.settings_dict("C3")), W, H) for _ in range(S)]
lib = capi.lib()
hp, img = host.data_ptr(), W * H


def work(group):
    torch.cuda.set_device(0)
    slot = C.c_int()
    for k in range(K):
        slots = []
        for s in group:
            off = ((s * 16 + k % 16) * 2) * img
            rc = lib.svo_upload_stereo(ctxs[s].h_ctx, C.c_void_p(hp + off), C.c_size_t(W), C.c_void_p(hp + off + img), C.c_size_t(W), C.byref(slot))
            assert rc == 0
            slots.append(slot.value)
        for s, sl in zip(group, slots):
            lib.svo_sync(ctxs[s].h_ctx)
            lib.svo_slot_release(ctxs[s].h_ctx, sl)


for mode in ("warm", "timed"):
    ts = [threading.Thread(target=work, args=(list(range(t, S, T)),)) for t in range(T)]
    t0 = time.perf_counter()
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    dt = time.perf_counter() - t0
    if mode == "timed":
        fps = S * K / dt
        print(f"ingest+pyramids only: {fps:.0f} pairs/s = {fps * 2 * img / 1e9:.1f} GB/s over PCIe ({S} contexts, {T} threads)")
