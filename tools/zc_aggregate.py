"""The box's aggregate host -> GPU ceiling with all GPUs reading at once (developer tool, torchrun):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 tools/zc_aggregate.py

Every rank reads its own page-locked 256 MB buffer over PCIe, all ranks between the same two barriers: zero-copy SM loads (the
ingest kernel's access pattern, svo_debug_zero_copy_bandwidth) and copy-engine DMA (cudaMemcpyAsync), first rank 0 alone, then all
ranks together.  Rank 0 prints one JSON line: GB/s per GPU and summed.  This is the roofline of bench.py's `e2e` at N GPUs:
e2e frames/s x 776 KB per frame against the aggregate figure."""
import ctypes as C
import json
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from stereo_svo_slam_b200 import capi, synth

rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = capi.Context(capi.CameraSettings(**synth.settings_dict("C3")), 752, 480, device=local)
NB = 256 << 20
h = torch.empty(NB, dtype=torch.uint8).pin_memory()
h.fill_(1)
d = torch.empty(NB, dtype=torch.uint8, device="cuda")
out = C.c_float()


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def zero_copy(active):
    barrier()
    t0 = time.perf_counter()
    gbs = 0.0
    if active:
        capi.lib().svo_debug_zero_copy_bandwidth(ctx.h_ctx, C.c_void_p(h.data_ptr()), C.c_size_t(NB), 148, 16, C.byref(out))   # 4 GiB
        gbs = out.value
    dt = time.perf_counter() - t0
    barrier()
    return gbs, dt


def dma(active):
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    gbs = 0.0
    if active:
        e0.record()
        for _ in range(16):
            d.copy_(h, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        gbs = NB * 16 / (e0.elapsed_time(e1) * 1e-3) / 1e9
    barrier()
    return gbs


def gather(v):
    if world == 1:
        return [v]
    t = torch.zeros(world, device="cuda", dtype=torch.float64)
    t[rank] = v
    dist.all_reduce(t)
    return t.tolist()


zero_copy(True)   # warm-up
res = {}
res["zero_copy_rank0_alone"] = gather(zero_copy(rank == 0)[0])
res["zero_copy_all"] = gather(zero_copy(True)[0])
res["dma_rank0_alone"] = gather(dma(rank == 0))
res["dma_all"] = gather(dma(True))
if rank == 0:
    print(json.dumps({"n_gpus": world, "buffer_mb": NB >> 20, "gbs_per_gpu": res, "gbs_sum": {k: sum(v) for k, v in res.items()},
                      "cpu_count": os.cpu_count()}))
if world > 1:
    dist.destroy_process_group()
