"""Multi-GPU is replicas only: ranks own disjoint sequences, no data-path collective; the only
communication is the barrier and the max-over-ranks time. Covered here with a world-size-2 gloo group on CPU."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def seeds_for(rank, streams):
    # bench.py: seeds = [1000 + rank * S + s for s in range(S)]
    return [1000 + rank * streams + s for s in range(streams)]


def _worker(rank, world, port, streams, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = torch.tensor(seeds_for(rank, streams), dtype=torch.int64)
    gathered = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)
    t = torch.tensor([0.5 + rank], dtype=torch.float64)     # pretend per-rank elapsed seconds
    dist.barrier()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        out.put((torch.cat(gathered).tolist(), float(t.item())))
    dist.destroy_process_group()


def test_rank_partition_and_max_time():
    world, streams = 2, 4
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, streams, q)) for r in range(world)]
    for p in procs:
        p.start()
    seeds, tmax = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(seeds) == list(range(1000, 1000 + world * streams)) and len(set(seeds)) == world * streams
    assert tmax == 1.5   # whole-job time = slowest rank
    # whole-job value = all ranks' frames / max time
    frames_per_rank = streams * 40
    assert abs(world * frames_per_rank / tmax - 213.333) < 1e-2
