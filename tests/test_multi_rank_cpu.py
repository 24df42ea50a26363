"""CPU tier, world_size 2 over gloo: the N>1 path of bench.py shards by rank and reduces only two numbers
(no data-path collective).  The shards themselves are checked to be distinct and deterministic."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    buf, off, ln = bench.make_workload(2048, bench.shard_seed(rank))
    ms_local = 10.0 * (rank + 1)  # rank 1 is the slow one
    value, ms = bench.job_throughput(buf.size, ms_local, world, "cpu")
    stats = bench.gather_rank_stats([ms_local - 1, ms_local, ms_local + 2], world, "cpu")
    assert stats == [[9.0, 10.0, 12.0], [19.0, 20.0, 22.0]], stats  # every rank sees every rank's min / med / max
    q.put((rank, int(buf.size), int(np.frombuffer(buf.tobytes(), np.uint8)[:4096].astype(np.uint64).sum()), value, ms))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_reduction():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 200)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, b0, s0, v0, m0), (r1, b1, s1, v1, m1) = res
    assert (r0, r1) == (0, 1)
    assert s0 != s1, "ranks must own different shards"
    assert m0 == m1 == 20.0, "job time is the max over ranks"
    expect = (b0 + b1) / 20e-3 / 1e6
    assert abs(v0 - expect) < 1e-6 * expect and v0 == v1
