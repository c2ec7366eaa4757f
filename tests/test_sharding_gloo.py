"""Multi-GPU is replicas only: ranks own disjoint sequences, there is no data-path collective;
the only communication is the barrier and the reduction of (time, work counters) for the one JSON line.  Covered here with a
world-size-2 gloo group on CPU, through bench.py's own partition and aggregation helpers (bench.stream_seeds,
bench.aggregate_ranks) — the code the N-GPU bench runs."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _worker(rank, world, port, streams, total, out):
    import bench
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)

    def all_reduce(vals, op):
        t = torch.tensor(vals, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
        return t.tolist()
    mine = bench.stream_seeds(rank, world, streams, total)
    pad = torch.full((64,), -1, dtype=torch.int64)
    pad[:len(mine)] = torch.tensor(mine, dtype=torch.int64)
    gathered = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(gathered, pad)
    dist.barrier()
    # pretend this rank tracked 40 frames per owned sequence in (0.5 + rank) seconds, with 7 keyframes and 1000 patches per frame
    frames = 40.0 * len(mine)
    seconds, sums = bench.aggregate_ranks(0.5 + rank, [frames, 7.0, 1000.0 * frames], all_reduce)
    if rank == 0:
        out.put(([int(v) for g in gathered for v in g.tolist() if v >= 0], seconds, sums))
    dist.destroy_process_group()


def _run(world, streams, total):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, streams, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return res


def test_weak_scaling_partition_and_aggregation():
    world, streams = 2, 4
    seeds, seconds, sums = _run(world, streams, 0)
    assert sorted(seeds) == list(range(1000, 1000 + world * streams)) and len(set(seeds)) == world * streams
    assert seconds == 1.5                                    # whole-job time = slowest rank
    assert sums == [320.0, 14.0, 320000.0]                   # work counters are summed over ranks
    assert abs(sums[0] / seconds - 213.333) < 1e-2           # whole-job value = all ranks' frames / max time


def test_configs4_partition_of_64_streams():
    # BASELINE configs[4]: 64 streams in total, stream s on GPU s mod N
    seeds, seconds, sums = _run(2, 32, 64)
    assert sorted(seeds) == list(range(1000, 1064))
    assert sums[0] == 64 * 40.0 and seconds == 1.5


def test_partition_is_disjoint_and_complete_for_every_world_size():
    import bench
    for world in (1, 2, 4, 8):
        got = [s for r in range(world) for s in bench.stream_seeds(r, world, 32, 64)]
        assert sorted(got) == list(range(1000, 1064))
        assert all(len(bench.stream_seeds(r, world, 32, 64)) == 64 // world for r in range(world))
        got = [s for r in range(world) for s in bench.stream_seeds(r, world, 5, 0)]
        assert sorted(got) == list(range(1000, 1000 + 5 * world))
    assert bench.aggregate_ranks(2.0, [1.0, 2.0]) == (2.0, [1.0, 2.0])


def test_alignment_cluster_policy_follows_the_streams_per_gpu():
    """bench.align_cluster_for: configs[4] puts 64 / 32 / 16 / 8 streams on a GPU at N = 1 / 2 / 4 / 8 — 2, 4, library default, library
    default SMs per alignment solve; the --align-cluster flag overrides."""
    import bench
    assert [bench.align_cluster_for(len(bench.stream_seeds(0, n, 48, 64))) for n in (1, 2, 4, 8)] == [2, 4, 0, 0]
    assert bench.align_cluster_for(48) == 4 and bench.align_cluster_for(56) == 2 and bench.align_cluster_for(23) == 0
    assert bench.align_cluster_for(48, 8) == 8 and bench.align_cluster_for(4, 0) == 0
