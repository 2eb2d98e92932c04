"""Multi-rank host logic on CPU (gloo, world size 2): lane sharding and the single collective of the
path, an all-reduce(SUM) of the MPCB_NSTATS statistics vector."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mpc_arpo_project_b200 import _lib


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    wl = bench.WORKLOADS["config2"]
    B = 8
    x0, noise = bench.make_inputs(wl, B, 1234 + rank)         # every rank draws its own shard
    stats = np.zeros(_lib.NSTATS)
    stats[0] = x0[0].sum()                                     # stand-ins for sum final_dist / lanes / solves
    stats[3] = B
    stats[5] = 100 * (rank + 1)
    t = torch.from_numpy(stats.copy())
    dist.all_reduce(t)
    timing = torch.tensor([10.0 * (rank + 1)], dtype=torch.float64)
    dist.all_reduce(timing, op=dist.ReduceOp.MAX)
    q.put((rank, x0, stats, t.numpy(), float(timing[0])))
    dist.destroy_process_group()


def test_shards_are_distinct_and_stats_all_reduce():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted((q.get(timeout=120) for _ in range(world)), key=lambda o: o[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, x0a, sa, ra, ta), (_, x0b, sb, rb, tb) = out
    assert not np.allclose(x0a, x0b)                           # different seeds -> different lanes
    np.testing.assert_allclose(ra, sa + sb)
    np.testing.assert_allclose(rb, sa + sb)
    assert ra[3] == 16 and ra[5] == 300                        # whole-job lanes and solves
    assert ta == tb == 20.0                                    # timing is the max over ranks
