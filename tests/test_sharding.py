"""N > 1 host logic on CPU: world_size-2 gloo process group, contiguous sharding with no data-path collective."""
import os
import socket

import numpy as np
import pytest


def test_shard_ranges_partition_and_balance():
    from mseetc.sharding import shard_ranges
    rng = np.random.default_rng(11)
    for world in (1, 2, 4, 8):
        for n in (1, 7, 4096, 16384):
            N = rng.choice([100, 200, 300, 400], size=n)
            r = shard_ranges(N, world)
            assert r[0][0] == 0 and r[-1][1] == n and all(a[1] == b[0] for a, b in zip(r[:-1], r[1:]))
            if n >= 64 * world:
                work = np.array([np.sum(N[a:b] + 1) for a, b in r], dtype=float)
                assert work.max() / work.mean() < 1.02
    assert shard_ranges(np.full(4096, 300), 8) == [(i * 512, (i + 1) * 512) for i in range(8)]


def _worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from mseetc.sharding import solve_sharded
    T = 1000.0 + np.arange(1001)

    def fake_solve(a, b):                  # stands in for casadiSolver.solve_batch on this rank's GPU
        return {'cost': T[a:b] * 2.0, 'status': np.zeros(b - a, dtype=np.int32), 'rank': np.full(b - a, rank)}

    res = solve_sharded(fake_solve, len(T), np.full(len(T), 300))
    if rank == 0:
        out.put({k: v.tolist() for k, v in res.items()})
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_gather():
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    cost = np.array(res['cost'])
    assert len(cost) == 1001 and np.array_equal(cost, (1000.0 + np.arange(1001)) * 2.0)     # order preserved
    ranks = np.array(res['rank'])
    assert set(ranks.tolist()) == {0, 1} and np.all(np.diff(ranks) >= 0) and abs(int((ranks == 0).sum()) - 500) <= 1
