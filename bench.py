#!/usr/bin/env python
"""bench.py — tracked stereo frames/s of the tracking hot path (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--streams S] [--frames-per-step F] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload): BASELINE.json configs[2] — EuRoC-shaped synthetic stereo, 752x480, 4-level pyramid,
30x24 grid (~430-500 keypoints/frame) — as S independent sequences per GPU.  configs[0]/[1] (Blender videos) cannot run:
the .mkv files are missing from the reference checkout (.MISSING_LARGE_BLOBS).  A *step* advances every sequence of
every rank by F stereo frames (default 100: `--steps 20` times 64 000 frames per GPU, more than a second) through
StereoSlam::new_image: pyramids -> sparse alignment -> KLT -> reprojection GN -> depth filter -> host bookkeeping,
keyframe creation when the reference would create one (the 64 rendered frames of a sequence cross the second keyframe;
`keyframes_in_timed_region` says how many fell into the timed steps).  One svo_slam_run_many call per step: the per-frame
host work is native code on `--host-threads` threads, Python runs once per step.

  value  : whole-job frames/s with the stereo frames already resident in HBM
  e2e    : the same through the reference-facing call with HOST (pinned) buffers; every frame crosses PCIe inside the timed
           region (the ingest kernel reads the page-locked images over PCIe, the results come back to host memory)
  aggregate : Mpatches/s (alignment) and Mwindows/s (KLT) of the whole job, from the library's work counters
  configs4  : BASELINE configs[4] literally — 64 streams in total, partitioned over the N GPUs (64 / N per GPU)
  roofline / single_stream / kernels : one sequence, per-stage CUDA events on the launching stream
  cpu_baseline : the reference itself (oracle/_ref: its unmodified sources built against oracle/cvshim) on one host core
  --impl reference : the same workload on the reference itself, one process per host core (the library is single-threaded
           and keeps process-global state), every step a bounded sample (F/12 frames per sequence)
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
import multiprocessing as mp

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# one hardware work queue per sequence: with the default of 8, sequences that share a queue serialise behind each
# other's long solver kernels (must be set before the CUDA context exists)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"   # keep stdout to the one JSON line (NCCL prints its version banner there)

from stereo_svo_slam_b200 import synth  # noqa: E402

CFG = "C3"
PROFILE_FILE = "r02_kernel_profile.json"   # per-kernel ncu figures of this round (tools/ncu_profile_json.py)
METRIC = "tracked_stereo_frames_per_s"
UNIT = "frames/s"


def render_sequence(args):
    cfg, seed, n = args
    seq = synth.make_sequence(cfg, seed=seed)
    c = synth.CONFIGS[cfg]
    out = np.empty((n, 2, c["height"], c["width"]), np.uint8)
    for k in range(n):
        out[k, 0], out[k, 1] = seq.render(k)
    return out


def make_frames(cfg, seeds, n):
    """[S][n][2][H][W] uint8; rendered in parallel processes and cached under /tmp."""
    import concurrent.futures as cf
    outs, todo = {}, []
    for s in seeds:
        path = f"/tmp/svo_synth_{cfg}_{s}_{n}.npy"
        if os.path.exists(path):
            try:
                outs[s] = np.load(path)
                continue
            except Exception:
                pass
        todo.append(s)
    if len(todo) == 1:   # rendered in-process (also the only safe way once CUDA is initialised: no fork)
        outs[todo[0]] = render_sequence((cfg, todo[0], n))
        try:
            np.save(f"/tmp/svo_synth_{cfg}_{todo[0]}_{n}.npy", outs[todo[0]])
        except Exception:
            pass
        todo = []
    if todo:
        with cf.ProcessPoolExecutor(max_workers=min(len(todo), os.cpu_count() or 1)) as ex:
            for s, arr in zip(todo, ex.map(render_sequence, [(cfg, s, n) for s in todo])):
                outs[s] = arr
                try:
                    np.save(f"/tmp/svo_synth_{cfg}_{s}_{n}.npy", arr)
                except Exception:
                    pass
    return np.stack([outs[s] for s in seeds])


class ClockSampler:
    """Samples SM clock and throttle reasons through NVML (in-process thread) during the timed region."""

    def __init__(self, dev, period=0.02):
        self.dev, self.period, self.rows, self.t, self.stop_flag = dev, period, [], None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates all GPUs of the box; CUDA_VISIBLE_DEVICES may remap: resolve by UUID when possible
            self.nv = pynvml
            idx = dev
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                ent = vis.split(",")[dev].strip()
                if ent.isdigit():
                    idx = int(ent)
                else:
                    for k in range(pynvml.nvmlDeviceGetCount()):
                        h = pynvml.nvmlDeviceGetHandleByIndex(k)
                        u = pynvml.nvmlDeviceGetUUID(h)
                        u = u.decode() if isinstance(u, bytes) else u
                        if u.startswith(ent) or ent in u:
                            idx = k
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        except Exception as e:  # noqa: BLE001
            self.nv, self.err = None, str(e)

    def _loop(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.rows.append((sm, r))
            except Exception:  # noqa: BLE001
                pass
            time.sleep(self.period)

    def start(self):
        if self.nv:
            self.t = threading.Thread(target=self._loop, daemon=True)
            self.t.start()

    def stop(self):
        if not self.nv:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: " + getattr(self, "err", "")]}
        self.stop_flag = True
        self.t.join()
        nv = self.nv
        try:
            mx = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            mx = None
        bits = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                "hw_power_brake_slowdown": 0x80}
        reasons = sorted({n for _, r in self.rows for n, b in bits.items() if r & b})
        sm = [a for a, _ in self.rows]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "samples": len(sm), "reasons": reasons}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def algorithmic_bytes(c, n_kps, evals_per_level, klt_levels=3):
    """SURVEY.md §8(d) per-unit figures x the units of one frame."""
    w, h, L = c["width"], c["height"], c["max_pyramid_levels"]
    lv = [(w >> i) * (h >> i) for i in range(L)]
    half = sum(lv[:-1]) + sum(lv[1:])
    lk = [w * h, ((w + 1) // 2) * ((h + 1) // 2)]
    lk.append(((lk and (w + 1) // 2) + 1) // 2 * ((((h + 1) // 2) + 1) // 2))
    pyr_lk = (lk[0] + lk[1]) + (lk[0] + lk[1] + lk[2])  # reads of levels 0,1 + padded writes (interior only counted)
    return {
        "pyramids": half + pyr_lk,
        "align": 120 * n_kps * evals_per_level,          # 120 B per (keypoint, level, evaluation) patch
        "klt": (3 * 6144 + 29) * n_kps,                  # 3 levels x (32x32 u8 + 32x32 s16x2 of the keyframe + 32x32 u8 of the frame) + I/O
        "ssd": 4468 * n_kps,
        "refine": 20 * n_kps,
        "filter": 72 * n_kps,
    }


# --------------------------------------------------------------------------------------------- reference arm
def _ref_kind():
    from oracle import oracle as orc
    return "reference" if orc.have_ref() else "port"


def _ref_worker(args):
    """One process = one host core: its share of the sequences on the reference itself (oracle/_ref), frame by frame."""
    cfg, seeds, nframes, F, W, K, start_evt, ready_q, done_q = args
    from oracle import oracle as orc
    c = synth.CONFIGS[cfg]
    frames = make_frames(cfg, seeds, nframes)
    cs = orc.CameraSettings(**synth.settings_dict(cfg))
    use_ref = orc.have_ref()
    slams = [(orc.RefSlam(cs, c["width"], c["height"]) if use_ref else orc.OracleSlam(cs, c["width"], c["height"], tracing=False)) for _ in seeds]

    def fresh(i):
        if use_ref:
            slams[i].close()
            return orc.RefSlam(cs, c["width"], c["height"])
        return orc.OracleSlam(cs, c["width"], c["height"], tracing=False)

    def step(k0):
        for k in range(k0, k0 + F):
            f = k % nframes
            for i in range(len(slams)):
                if f == 0 and k > 0:          # the stream starts its next sequence: a new StereoSlam, like the b200 arm's reset
                    slams[i] = fresh(i)
                slams[i].new_image(frames[i, f, 0], frames[i, f, 1], f / 20.0)
    for w in range(W):
        step(w * F)
    ready_q.put(1)
    start_evt.wait()
    t0 = time.time()
    for k in range(K):
        step((W + k) * F)
    done_q.put((t0, time.time(), len(seeds) * F * K))


def run_reference(cfg, seeds, nframes, F, W, K, procs):
    """frames/s of the reference on `procs` host cores; returns (fps, seconds, kind)."""
    ctx = mp.get_context("spawn")
    groups = [seeds[p::procs] for p in range(procs) if seeds[p::procs]]
    start_evt, ready_q, done_q = ctx.Event(), ctx.Queue(), ctx.Queue()
    ps = [ctx.Process(target=_ref_worker, args=((cfg, g, nframes, F, W, K, start_evt, ready_q, done_q),)) for g in groups]
    for p_ in ps:
        p_.start()
    for _ in ps:
        ready_q.get()
    start_evt.set()
    res = [done_q.get() for _ in ps]
    for p_ in ps:
        p_.join()
    seconds = max(r[1] for r in res) - min(r[0] for r in res)
    return sum(r[2] for r in res) / seconds, seconds, _ref_kind()


def run_reference_one_core(frames, cfg, passes, warm):
    """the reference on ONE host core over the first stream: `passes` passes over its rendered sequence (every pass a fresh
    StereoSlam, keyframe creation included — the b200 arm's stream definition) after `warm` warm-up frames; frames/s"""
    from oracle import oracle as orc
    c = synth.CONFIGS[cfg]
    cs = orc.CameraSettings(**synth.settings_dict(cfg))

    def fresh():
        return orc.RefSlam(cs, c["width"], c["height"]) if orc.have_ref() else orc.OracleSlam(cs, c["width"], c["height"], tracing=False)
    sl = fresh()
    for k in range(warm):
        sl.new_image(frames[0, k, 0], frames[0, k, 1], k / 20.0)
    if orc.have_ref():
        sl.close()
    nf = frames.shape[1]
    t0 = time.perf_counter()
    for _ in range(passes):
        sl = fresh()
        for k in range(nf):
            sl.new_image(frames[0, k, 0], frames[0, k, 1], k / 20.0)
        if orc.have_ref():
            sl.close()
    dt = time.perf_counter() - t0
    return passes * nf / dt, dt, _ref_kind()


def prefer_host_memory_near_gpu(local_rank):
    """Best effort: make this process allocate its (page-locked) frame buffers on the NUMA node its GPU hangs off
    (set_mempolicy(MPOL_PREFERRED)).  With 8 ranks reading 0.78 MB per frame over PCIe, frames that sit on the other socket cross the
    inter-socket link first.  Returns what was done, for the JSON line."""
    import platform
    info = {"numa_node": None, "policy": "default"}
    try:
        import torch
        pr = torch.cuda.get_device_properties(local_rank)
        bus = "%04x:%02x:%02x.0" % (getattr(pr, "pci_domain_id", 0), pr.pci_bus_id, pr.pci_device_id)
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        info["numa_node"] = node
        info["nodes_online"] = open("/sys/devices/system/node/online").read().strip()
        if node >= 0:
            nr = {"x86_64": 238, "aarch64": 237}.get(platform.machine())
            if nr is not None:
                libc = C.CDLL(None, use_errno=True)
                mask = C.c_ulong(1 << node)
                rc = libc.syscall(nr, 1, C.byref(mask), C.c_ulong(C.sizeof(mask) * 8))   # MPOL_PREFERRED = 1
                info["policy"] = f"preferred node {node}" if rc == 0 else f"default (set_mempolicy errno {C.get_errno()})"
    except Exception as e:  # noqa: BLE001
        info["policy"] = "default (" + str(e)[:80] + ")"
    return info


def stream_seeds(rank, world, streams_per_gpu, total_streams=0):
    """Which sequences (by seed) a rank owns.  Weak scaling: streams_per_gpu sequences per GPU, seeds 1000 + rank*S + s.
    BASELINE configs[4] (total_streams > 0): that many sequences in total, stream s on GPU s mod N, seeds 1000 + s.
    Sequences are independent: no rank ever needs another rank's data (replicas only)."""
    if total_streams > 0:
        return [1000 + s for s in range(total_streams) if s % world == rank]
    return [1000 + rank * streams_per_gpu + s for s in range(streams_per_gpu)]


def align_cluster_for(n_seq, override=-1):
    """SMs per alignment solve (svo_set_align_cluster) for n_seq sequences on one GPU: what a frame costs the GPU is the time its
    cluster holds its SMs, so smaller clusters win once there are enough sequences to hide their longer solves (measured, DESIGN.md
    section 4).  override >= 0: the --align-cluster flag (0 = library default: 8, or 16 above 1024 keypoints)."""
    if override >= 0:
        return override
    return 2 if n_seq >= 56 else (4 if n_seq >= 24 else 0)


def aggregate_ranks(seconds, counters, all_reduce=None):
    """Whole-job figures from per-rank ones: the time is the MAX over ranks, the work counters are SUMMED.
    all_reduce(list_of_floats, "max" | "sum") -> list; None for a single process."""
    if all_reduce is None:
        return seconds, list(counters)
    return all_reduce([seconds], "max")[0], all_reduce(list(counters), "sum")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--frames-per-step", type=int, default=100, help="frames every sequence advances per step")
    ap.add_argument("--frames", type=int, default=64, help="rendered frames per sequence (played forward/backward: continuous motion)")
    ap.add_argument("--streams", type=int, default=48, help="independent sequences per GPU (weak scaling)")
    ap.add_argument("--total-streams", type=int, default=0, help="BASELINE configs[4]: this many sequences in total, stream s on GPU s mod N (strong scaling)")
    ap.add_argument("--host-threads", type=int, default=0, help="native host threads driving the sequences of one GPU (0 = min(16, cores / ranks))")
    ap.add_argument("--align-cluster", type=int, default=-1, choices=[-1, 0, 1, 2, 4, 8, 16],
                    help="SMs per alignment solve in the multi-sequence runs (svo_set_align_cluster; 0 = library default 8/16).  -1 (default): by the "
                         "number of sequences on the GPU — 2 from 56 sequences, 4 from 24, else the library default.  Measured on one B200 "
                         "(frames/s, device-resident; gpurun_out/s10_sweep.txt, s11_sweep.txt): 32 sequences 57.8k / 60.5k / 57.1k / 45.9k with 8 / 4 / 2 / 1 "
                         "SMs, 48 sequences 57.8k / 61.6k / 62.9k, 64 sequences 61.1k (4) / 63.5k (2) / 61.7k (1): what a frame costs the GPU is the time its "
                         "cluster holds its SMs, so smaller clusters win as soon as there are enough sequences to hide their longer solves")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs4", action="store_true")
    ap.add_argument("--playback", default="restart", choices=["restart", "sweep"], help="restart: streams of finite sequences (default); "
                    "sweep: one endless sequence per stream, no keyframes in steady state (developer comparison with round 1)")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    W = max(a.warmup, 3)
    K, F = a.steps, max(1, a.frames_per_step)
    if a.host_threads <= 0:
        a.host_threads = max(1, min(16, (os.cpu_count() or 1) // max(1, world)))
    c = synth.CONFIGS[CFG]
    nframes = max(2, a.frames)   # rendered frames; frame t of a sequence shows rendered frame tri(t): forward/backward sweep

    def tri(t):
        p = t % (2 * nframes - 2)
        return p if p < nframes else 2 * nframes - 2 - p

    def seeds_for(rank_, world_, total):
        return stream_seeds(rank_, world_, a.streams, total)
    seeds = seeds_for(rank, world, a.total_streams)
    S = len(seeds)
    config = {"workload": f"BASELINE configs[2] EuRoC-shaped synthetic 752x480 ({CFG}: 4-level pyramid, 30x24 grid), "
                          + (f"configs[4]: {a.total_streams} independent sequences in total, stream s on GPU s mod N" if a.total_streams else
                             f"{a.streams} independent sequences per GPU") + "; configs[0]/[1] blocked: .mkv missing",
              "sequences_per_gpu": S, "frames_per_step_per_sequence": F, "host_threads_per_gpu": a.host_threads,
              "rendered_frames_per_sequence": nframes,
              "align_cluster": "SMs per alignment solve (svo_set_align_cluster): %s for the %d sequences per GPU of this run (0 = library default 8); "
                               "a deployment knob, results within the pose tolerance for every setting" % (
                                   align_cluster_for(S, a.align_cluster), S),
              "playback": "a stream is a succession of finite sequences (BASELINE configs[4] streams are 200 frames long): the rendered frames "
                          "played forward, then the stream restarts with a fresh tracker on the same device resources (svo_slam_reset) — every "
                          "pass creates keyframe #1 at its frame 0 and keyframe #2 around frame 46, like the reference would",
              "seeds": "1000 + s, s mod N == rank" if a.total_streams else "1000 + rank*S + s",
              "l2_hygiene": "every step touches new frames (0.72 MB per frame and sequence: 2.3 GB per step and GPU, far beyond the 126 MB L2) "
                            "and all sequences' pyramids; a single sequence's working set (<3 MB) is L2 resident by nature of the path — "
                            "the kernels are latency/issue bound (DESIGN.md)"}

    if a.impl == "reference":
        # the reference's own CPU implementation of the path: oracle/_ref (its unmodified sources), one process per host core,
        # rank 0 only.  A step is a bounded sample of the b200 arm's step: F/12 frames per sequence.
        if rank != 0:
            return
        procs = min(S, os.cpu_count() or 1)
        f_ref = max(1, F // 12)
        fps, seconds, kind = run_reference(CFG, seeds, nframes, f_ref, W, K, procs)
        sample = (f"{S} sequences x {K} steps x {f_ref} frames (a step of the b200 arm is {F} frames per sequence) after {W} warm-up steps, "
                  f"one process per host core ({procs}); " + ("oracle/_ref = the reference's unmodified sources built against oracle/cvshim"
                                                             if kind == "reference" else "oracle port (oracle/_ref not built)"))
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": a.gpus, "steps": K, "warmup": W,
                          "ms_per_step": 1e3 * seconds / K, "higher_is_better": True, "scaling": "strong" if a.total_streams else "weak",
                          "vs_baseline": None, "dtype": "u8/f32", "data": "synthetic", "config": config,
                          "cpu_baseline": {"value": fps, "unit": UNIT, "cores": procs, "kind": kind, "sample": sample},
                          "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import torch
    import torch.distributed as dist
    from stereo_svo_slam_b200 import StereoSlam, capi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    config["host_memory"] = prefer_host_memory_near_gpu(local_rank)
    H_, W_ = c["height"], c["width"]
    img = H_ * W_
    settings = capi.CameraSettings(**synth.settings_dict(CFG))
    lib = capi.lib()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def all_reduce(vals, op):
        t = torch.tensor(vals, device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
        return t.tolist()

    def load_inputs(seed_list):
        fr = make_frames(CFG, seed_list, nframes)                      # [S][F][2][H][W]
        h = torch.from_numpy(fr).pin_memory()                          # pinned host inputs (e2e)
        return fr, h, h.to("cuda", non_blocking=False)                 # HBM-resident inputs (value)

    def cluster_for(n_seq):
        return align_cluster_for(n_seq, a.align_cluster)

    def run(mode, host, dev, n_seq, clocks=None):
        """W warm-up steps, then K timed steps of F frames per sequence; returns whole-rank counters and the time (max over ranks)"""
        slams = [StereoSlam(settings, W_, H_, device=local_rank) for _ in range(n_seq)]
        for sl in slams:   # many sequences share the GPU: SM time per frame matters, not the latency of one solve
            sl.set_align_cluster(cluster_for(n_seq))
        base = dev.data_ptr() if mode == "device" else host.data_ptr()
        handles = (C.c_void_p * n_seq)(*[sl._h for sl in slams])
        bad = C.c_int(-1)
        step_times = []
        # the frame tables of all steps (pointers into the resident frame tensor, timestamps) are laid out before the clock starts
        seq_off = (np.arange(n_seq, dtype=np.int64) * nframes)[:, None]
        tables = []
        for k in range(W + K):
            t = k * F + np.arange(F, dtype=np.int64)
            if a.playback == "sweep":                         # developer option: one endless sequence, forward/backward over the frames
                fr = np.array([tri(int(x)) for x in t], np.int64)[None, :]
                ts = np.tile((t / 20.0).astype(np.float32), n_seq)
                rs = np.zeros(n_seq * F, np.uint8)
            else:
                fr = (t % nframes)[None, :]                   # a stream = the rendered sequence over and over, each pass a new sequence
                ts = np.tile(((t % nframes) / 20.0).astype(np.float32), n_seq)
                rs = np.tile(((t % nframes == 0) & (t > 0)).astype(np.uint8), n_seq)
            left = (base + (seq_off + fr) * 2 * img).astype(np.uint64).reshape(-1)
            right = left + np.uint64(img)
            tables.append((np.ascontiguousarray(left), np.ascontiguousarray(right), np.ascontiguousarray(ts), np.ascontiguousarray(rs)))

        def step(k):
            lp, rp, ts, rs = tables[k]
            t0 = time.perf_counter()
            rc = lib.svo_slam_run_many_restart(handles, n_seq, F, lp.ctypes.data_as(C.c_void_p), rp.ctypes.data_as(C.c_void_p), C.c_size_t(W_),
                                               C.c_size_t(W_), ts.ctypes.data_as(C.c_void_p), rs.ctypes.data_as(C.c_void_p),
                                               1 if mode == "device" else 0, a.host_threads, C.byref(bad))
            step_times.append(time.perf_counter() - t0)
            if rc:
                raise RuntimeError(f"sequence {bad.value}: " + lib.svo_slam_last_error(slams[max(bad.value, 0)]._h).decode())

        def totals():
            tot = np.zeros(8, np.int64)
            cnt = C.c_longlong()
            nl = 0
            for sl in slams:
                tot += np.array(list(sl.total_counters().values()), np.int64)
                lib.svo_launch_count(C.c_void_p(lib.svo_slam_ctx(sl._h)), C.byref(cnt))
                nl += cnt.value
            return tot, nl
        for k in range(W):
            step(k)
        step_times.clear()
        tot0, nl0 = totals()
        barrier()
        if clocks:
            clocks.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t0 = time.perf_counter()
        for k in range(W, W + K):
            step(k)
        torch.cuda.synchronize()
        e1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        dev_s = e0.elapsed_time(e1) / 1e3
        ck = clocks.stop() if clocks else None
        tot1, nl1 = totals()
        d = tot1 - tot0
        kps = int(np.mean([len(sl.get_frame().kps) for sl in slams]))
        last_gpu_ms = float(np.mean([sl.last_stats()["gpu_ms"] for sl in slams]))   # device time of every sequence's last frame (latency under load)
        for sl in slams:
            sl.close()
        seconds, sums = aggregate_ranks(max(wall, dev_s), [float(x) for x in d] + [float(nl1 - nl0)], all_reduce if world > 1 else None)
        st_ = sorted(step_times)
        trace = {"host_step_ms_p50": 1e3 * st_[len(st_) // 2], "host_step_ms_max": 1e3 * st_[-1], "host_us_per_frame_p50": 1e6 * st_[len(st_) // 2] / (n_seq * F),
                 "frame_latency_ms_under_load": last_gpu_ms} if st_ else {}
        return dict(seconds=seconds, frames=sums[0], tracking_frames=sums[1], keyframes=sums[2], patches=sums[3], windows=sums[4], keypoints=sums[5],
                    launches=int(sums[8]), kps=kps, clocks=ck, trace=trace)

    def summarize(r):
        fps = r["frames"] / r["seconds"]
        return {"value": fps, "mpatches_per_s": r["patches"] / r["seconds"] / 1e6, "mwindows_per_s": r["windows"] / r["seconds"] / 1e6,
                "keyframes_in_timed_region": int(r["keyframes"]), "frames_in_timed_region": int(r["frames"]), "seconds": r["seconds"]}

    frames_np, host, dev = load_inputs(seeds)
    if os.environ.get("BENCH_ONLY") == "host":     # developer switch: end-to-end run only
        r = run("host", host, dev, S)
        if rank == 0:
            print(json.dumps({"e2e": summarize(r), "trace": r["trace"]}))
        return
    r_dev = run("device", host, dev, S, None if os.environ.get("BENCH_NO_NVML") else ClockSampler(local_rank))
    if os.environ.get("BENCH_ONLY") == "device":   # developer switch: whole-job device-resident run only
        if rank == 0:
            print(json.dumps({"value": summarize(r_dev), "trace": r_dev["trace"]}))
        return
    r_e2e = run("host", host, dev, S)
    value, e2e = r_dev["frames"] / r_dev["seconds"], r_e2e["frames"] / r_e2e["seconds"]
    n_kps_multi = r_dev["keypoints"] / max(r_dev["tracking_frames"], 1.0)

    # ---------------- BASELINE configs[4] as stated: 64 streams in total over the N GPUs (strong scaling), same step definition
    c4s = None
    if not a.no_configs4 and not a.total_streams:
        seeds64 = seeds_for(rank, world, 64)
        if len(seeds64) == S and seeds64 == seeds:
            fr64, h64, d64 = frames_np, host, dev
        else:
            del dev
            fr64, h64, d64 = load_inputs(seeds64)
        r64d = run("device", h64, d64, len(seeds64))
        r64h = run("host", h64, d64, len(seeds64))
        c4s = {"workload": "BASELINE configs[4]: 64 independent 752x480 streams in total, stream s on GPU s mod N (seeds 1000..1063)",
               "streams_per_gpu": len(seeds64), "scaling": "strong", "device_resident": summarize(r64d), "e2e": summarize(r64h)}
        del d64, h64
        dev = None

    # ---------------- what the box delivers from page-locked host memory when ALL ranks read at once (the ceiling of e2e at N GPUs:
    # on the 8-GPU boxes of this pool a GPU gets 23-35 GB/s then, against 51 GB/s alone): every rank reads 64 MB x 8 between barriers
    link_all = None
    if world > 1:
        try:
            probe = torch.empty(64 << 20, dtype=torch.uint8).pin_memory()
            pctx = capi.Context(settings, W_, H_, device=local_rank)
            gbs = C.c_float()
            mine = 0.0
            for _ in range(3):
                barrier()
                if lib.svo_debug_zero_copy_bandwidth(pctx.h_ctx, C.c_void_p(probe.data_ptr()), C.c_size_t(64 << 20), 148, 8, C.byref(gbs)) == 0:
                    mine = max(mine, gbs.value)
            barrier()
            pctx.close()
            del probe
            tmin = torch.tensor([mine], device="cuda", dtype=torch.float64)
            tsum = tmin.clone()
            dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
            dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
            link_all = {"slowest_gpu_gbs": float(tmin.item()), "sum_gbs": float(tsum.item())}
        except Exception as e:  # noqa: BLE001
            link_all = {"error": str(e)[:100]}

    out = None
    if rank == 0:
        # ---------------- single sequence, the way an application runs it (one synchronous new_image per frame, page-locked
        # frames, CUDA-graph replay): wall time per call and device time per frame (events around ingest .. D2H)
        host_np = host.numpy()
        n1 = W + 200

        def lone(width):
            sl_ = StereoSlam(settings, W_, H_, device=local_rank)
            sl_.set_solver_width(width)
            walls_, gpu_ = [], []
            for k in range(n1):
                t0 = time.perf_counter()
                sl_.new_image(host_np[0, tri(k), 0], host_np[0, tri(k), 1], k / 20.0)
                walls_.append(time.perf_counter() - t0)
                stt = sl_.last_stats()
                if not stt["keyframe_created"]:
                    gpu_.append(stt["gpu_ms"])
            sl_.close()
            return float(np.median(walls_[W:])), float(np.median(gpu_[W:]))
        wall_graph, gpu_graph = lone(-1)       # library default: a lone sequence gets the wide line searches (svo_set_solver_width)
        wall_seq, gpu_seq = lone(0)            # sequential line searches (what the multi-sequence runs use)
        # ---------------- the same sequence with per-stage CUDA events (kernel-by-kernel launches): stage table, roofline
        sl = StereoSlam(settings, W_, H_, device=local_rank)
        ctxp = C.c_void_p(lib.svo_slam_ctx(sl._h))
        lib.svo_set_profiling(ctxp, 1)
        stage = np.zeros((n1, 8), np.float32)
        cnt = np.zeros((n1, 8), np.float64)
        walls = []
        buf = (C.c_float * 8)()
        for k in range(n1):
            t0 = time.perf_counter()
            sl.new_image(frames_np[0, tri(k), 0], frames_np[0, tri(k), 1], k / 20.0)
            walls.append(time.perf_counter() - t0)
            lib.svo_last_stage_ms(ctxp, buf)
            stage[k] = np.array(list(buf))
            cnt[k] = list(sl.last_counters().values())
        keep = slice(W, n1)
        st = np.median(stage[keep], axis=0)
        cm = cnt[keep].mean(axis=0)   # per-frame means of the work counters
        names = ["upload+pyramids", "sparse_align", "klt", "reproj_refine", "stereo_ssd", "depth_filter", "d2h", "total"]
        n_kps = int(round(cm[0]))
        sl.close()
        patches = cm[1] * (cm[2] + cm[3])          # (keypoint, level, evaluation) 4x4 patches per frame
        windows = cm[6]                            # (keypoint, level, LK iteration) 31x31 windows per frame
        peak, peak_src = peaks()
        prof = {}
        try:
            prof = json.load(open(os.path.join(ROOT, "profiles", PROFILE_FILE)))
        except Exception:  # noqa: BLE001
            pass

        def traffic_of(kernel):
            for name, rec in prof.get("kernels", {}).items():
                if name.startswith(kernel):
                    return rec.get("dram_bytes_per_launch")
            return None
        stage_kernel = {1: "sparse_align_kernel", 2: "klt31w_kernel", 3: "reproj_refine_kernel", 4: "stereo_ssd_mma_kernel", 5: "depth_filter_kernel"}
        med_wall = float(np.median(walls[W:]))
        dom = int(np.argmax(st[1:6])) + 1
        ab = algorithmic_bytes(c, n_kps, evals_per_level=1)
        ab["align"] = int(120 * patches)           # 120 B per patch x patches of one launch
        dom_bytes = {1: ab["align"], 2: ab["klt"], 3: ab["refine"], 4: ab["ssd"], 5: ab["filter"]}[dom]
        achieved = dom_bytes / (st[dom] * 1e-3) / 1e9 if st[dom] > 0 else 0.0
        kernels = []
        for i, nm in enumerate(names[:7]):
            b = {0: ab["pyramids"], 1: ab["align"], 2: ab["klt"], 3: ab["refine"], 4: ab["ssd"], 5: ab["filter"], 6: 0}[i]
            kernels.append({"stage": nm, "ms": float(st[i]), "share": float(st[i] / max(st[7], 1e-9)), "algorithmic_bytes": int(b),
                            "achieved_gbs": float(b / (st[i] * 1e-3) / 1e9) if st[i] > 0 else None})
        klt_fused_bytes = 6965 * n_kps             # SURVEY §8d primary definition (derivative-fused layout)
        kernels[2]["achieved_gbs_survey_fused_layout"] = float(klt_fused_bytes / (st[2] * 1e-3) / 1e9) if st[2] > 0 else None
        kernels[2]["algorithmic_bytes_survey_fused_layout"] = int(klt_fused_bytes)
        # ---------------- the bound that applies: instruction issue.  A frame's warp instructions (ncu smsp__inst_executed.sum of every
        # kernel of a frame, profiles/) over the chip's issue rate (148 SMs x 4 schedulers x 1 instruction per clock) is the least
        # time a frame can take on a saturated GPU however well latencies are hidden; achieved = what a frame takes at saturation.
        sm_clock = (r_dev["clocks"] or {}).get("sm_mhz") or 1965.0
        inst_per_frame = float(prof.get("warp_instructions_per_frame", 8.0e6))
        floor_us = inst_per_frame / (148 * 4 * sm_clock * 1e6) * 1e6
        per_gpu_fps = value / world
        issue = {"bound": "issue", "kernel": "whole frame (8 kernels)", "achieved": per_gpu_fps, "peak": 1e6 / floor_us, "unit": "frames/s per GPU",
                 "frac": per_gpu_fps * floor_us / 1e6, "us_per_frame_at_saturation": 1e6 / per_gpu_fps, "issue_floor_us_per_frame": floor_us,
                 "warp_instructions_per_frame": inst_per_frame, "sm_mhz": sm_clock,
                 "source": f"profiles/{PROFILE_FILE} (ncu smsp__inst_executed.sum per kernel of one C3 tracking frame)" if prof else
                           "profiles/r01 figure (8.0 M warp instructions per frame)"}
        # PCIe context for e2e: what the link delivers to SM-driven (zero-copy) reads of page-locked memory, measured live
        pcie = {}
        try:
            probe = torch.empty(64 << 20, dtype=torch.uint8).pin_memory()
            pctx = capi.Context(settings, W_, H_, device=local_rank)
            gbs = C.c_float()
            best = 0.0
            for _ in range(3):
                if lib.svo_debug_zero_copy_bandwidth(pctx.h_ctx, C.c_void_p(probe.data_ptr()), C.c_size_t(64 << 20), 148, 8, C.byref(gbs)) == 0:
                    best = max(best, gbs.value)
            pctx.close()
            del probe
            per_frame = 2 * H_ * W_ + n_kps * 45 + n_kps * 90
            pcie = {"bytes_per_frame": int(per_frame), "achieved_gbs": float(e2e / world * per_frame / 1e9), "link_peak_gbs": float(best),
                    "all_gpus_reading": None if link_all is None else dict(
                        link_all, frames_per_s_ceiling=(world * link_all["slowest_gpu_gbs"] * 1e9 / per_frame) if "slowest_gpu_gbs" in link_all else None,
                        note="zero-copy read bandwidth with every rank reading its own page-locked buffer at once; every rank has the same work and a "
                             "step ends with the slowest rank, so N x the slowest GPU's share / bytes per frame bounds e2e at N GPUs"),
                    "frac": float(e2e / world * per_frame / 1e9 / best) if best > 0 else None,
                    "note": "per GPU; link_peak = zero-copy read of a 64 MB page-locked buffer by 148 CTAs (one stream, sequential), "
                            "the e2e traffic is the interleaved streams' 722 KB frames plus the keypoint blocks in both directions"}
        except Exception as e:  # noqa: BLE001
            pcie = {"error": str(e)}
        sd, se = summarize(r_dev), summarize(r_e2e)
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
               "ms_per_step": 1e3 * r_dev["seconds"] / K, "higher_is_better": True, "scaling": "strong" if a.total_streams else "weak", "vs_baseline": None,
               "dtype": "u8/f32", "data": "synthetic", "config": config,
               "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(S * F * (2 * H_ * W_ + n_kps_multi * 45)),
                       "d2h_bytes_per_step": int(S * F * n_kps_multi * 90), "ms_per_step": 1e3 * r_e2e["seconds"] / K,
                       "mpatches_per_s": se["mpatches_per_s"], "mwindows_per_s": se["mwindows_per_s"],
                       "keyframes_in_timed_region": se["keyframes_in_timed_region"], "pcie": pcie},
               "aggregate": {"frames_per_s": value, "mpatches_per_s": sd["mpatches_per_s"], "mwindows_per_s": sd["mwindows_per_s"],
                             "n_gpus": world, "note": "whole job, device-resident run; patch = (keypoint, level, evaluation) 4x4 residual patch of the "
                                                      "alignment, window = (keypoint, level, LK iteration) 31x31 window (SURVEY.md §8d)"},
               "keyframes_in_timed_region": sd["keyframes_in_timed_region"], "frames_in_timed_region": sd["frames_in_timed_region"],
               "timed_seconds": r_dev["seconds"],
               "gpu_launches": int(r_dev["launches"]),
               "host_step_trace": {"value": r_dev["trace"], "e2e": r_e2e["trace"]},
               "clocks": r_dev["clocks"],
               "keypoints_per_frame": r_dev["kps"],
               "configs4": c4s,
               "single_stream": {"frames_per_s": 1.0 / wall_graph, "ms_per_frame_wall": 1e3 * wall_graph, "ms_per_frame_gpu": gpu_graph,
                                 "ms_per_frame_gpu_sequential_line_search": gpu_seq, "ms_per_frame_wall_sequential_line_search": 1e3 * wall_seq,
                                 "ms_per_frame_gpu_staged_events": float(st[7]), "ms_per_frame_wall_staged_events": 1e3 * med_wall,
                                 "pose_iter_latency_us": float(1e3 * st[1] / max(cm[2] + cm[3], 1.0)),
                                 "align_evaluations_per_frame": float(cm[2] + cm[3]),
                                 "mpatches_per_s": float(patches / (st[1] * 1e-3) / 1e6) if st[1] > 0 else None,
                                 "mwindows_per_s": float(windows / (st[2] * 1e-3) / 1e6) if st[2] > 0 else None,
                                 "note": "one sequence, synchronous new_image calls with page-locked host frames (what the reference app does), wide line "
                                         "searches (svo_set_solver_width default for a lone sequence; *_sequential_line_search: width 0, same bits); "
                                         "*_staged_events: same with a CUDA event between the stages (kernel-by-kernel launches, source of the stage table)"},
               "roofline": dict(issue, hbm={"bound": "hbm", "kernel": names[dom], "achieved": achieved, "peak": peak, "unit": "GB/s",
                                            "frac": achieved / peak, "traffic": traffic_of(stage_kernel.get(dom, "?")),
                                            "traffic_source": f"profiles/{PROFILE_FILE} (ncu --set full, dram__bytes_read+write per launch)",
                                            "algorithmic_bytes_per_launch": int(dom_bytes), "peak_source": peak_src,
                                            "note": "the dominant kernel of a lone sequence against the HBM roofline, for the record: the per-frame "
                                                    "working set (<3 MB) is L2 resident, the kernel is bound by dependent-issue latency (SURVEY.md §8d)"},
                                traffic=traffic_of(stage_kernel.get(dom, "?"))),
               "kernels": kernels}
        if world == 1 and not os.environ.get("BENCH_NO_C4"):
            # ---------------- BASELINE configs[3]: high-density stress (1280x720, 5 levels, ~3-4k keypoints), one sequence
            c4 = synth.CONFIGS["C4"]
            f4 = make_frames("C4", [c4["seed"]], 20)[0]
            s4 = StereoSlam(capi.CameraSettings(**synth.settings_dict("C4")), c4["width"], c4["height"], device=local_rank)
            ctx4 = C.c_void_p(lib.svo_slam_ctx(s4._h))
            c4_cluster = int(os.environ.get("BENCH_C4_CLUSTER", "0"))   # 0 = library default: 16 SMs per solve above 1024 keypoints
            s4.set_align_cluster(c4_cluster)
            lib.svo_set_profiling(ctx4, 1)
            st4, cn4 = [], []
            for k in range(20):
                s4.new_image(f4[k, 0], f4[k, 1], k / 20.0)
                lib.svo_last_stage_ms(ctx4, buf)
                st4.append(list(buf))
                cn4.append(list(s4.last_counters().values()))
            st4, cn4 = np.median(np.array(st4[4:]), axis=0), np.array(cn4[4:], np.float64).mean(axis=0)
            s4.close()
            out["stress_c4"] = {"workload": "BASELINE configs[3]: 1280x720 synthetic stereo, 5-level pyramid, 16x14 grid, one sequence, per-stage CUDA events",
                                "keypoints_per_frame": int(cn4[0]), "align_cluster": c4_cluster, "ms_per_frame_gpu": float(st4[7]),
                                "stage_ms": {nm: float(st4[i]) for i, nm in enumerate(names[:7])},
                                "pose_iter_latency_us": float(1e3 * st4[1] / max(cn4[2] + cn4[3], 1.0)),
                                "mpatches_per_s": float(cn4[1] * (cn4[2] + cn4[3]) / (st4[1] * 1e-3) / 1e6) if st4[1] > 0 else None,
                                "mwindows_per_s": float(cn4[6] / (st4[2] * 1e-3) / 1e6) if st4[2] > 0 else None,
                                "align_gbs": float(120 * cn4[1] * (cn4[2] + cn4[3]) / (st4[1] * 1e-3) / 1e9) if st4[1] > 0 else None,
                                "klt_gbs_materialised_derivative_layout": float((3 * 6144 + 29) * cn4[0] / (st4[2] * 1e-3) / 1e9) if st4[2] > 0 else None,
                                "klt_gbs_survey_fused_layout": float(6965 * cn4[0] / (st4[2] * 1e-3) / 1e9) if st4[2] > 0 else None}
        if not a.no_cpu_baseline and world == 1:
            pb, wb = int(os.environ.get("BENCH_CPU_PASSES", "6")), 3
            fps, dt, kind = run_reference_one_core(frames_np, CFG, pb, wb)
            out["cpu_baseline"] = {"value": fps, "unit": UNIT, "cores": 1, "kind": kind, "seconds": dt,
                                   "sample": f"1 stream, {pb} passes over its {nframes}-frame sequence ({pb * nframes} frames, {dt:.1f} s; every pass a fresh "
                                             f"tracker with its keyframes) after {wb} warm-up frames on one host core ("
                                             + ("oracle/_ref = the reference's unmodified sources built against oracle/cvshim" if kind == "reference"
                                                else "oracle port; oracle/_ref not built") + "; the reference library is single-threaded)"}
        print(json.dumps(out, default=lambda o: float(o) if isinstance(o, (np.floating,)) else (int(o) if isinstance(o, np.integer) else str(o))))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
