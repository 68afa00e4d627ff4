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

    def fake_solve(idx):                   # stands in for casadiSolver.solve_batch on this rank's GPU
        return {'cost': T[idx] * 2.0, 'status': np.zeros(len(idx), dtype=np.int32), 'rank': np.full(len(idx), rank)}

    res = solve_sharded(fake_solve, len(T), np.full(len(T), 300))
    tiled = solve_sharded(fake_solve, len(T), partition='tiles')

    class FakeSolver:                      # the part of casadiSolver that solve_batch_sharded touches
        numIntervals = 300

        def solve_batch(self, T, t0, vN, v0, overrides=None, to_host=True, **kw):
            T = np.atleast_1d(T)
            return {'z': np.stack([T, T + overrides['mass']], axis=1), 'obj': T * 3.0, 'kkt': np.zeros(len(T)), 'iters': np.full(len(T), rank, np.int32),
                    'status': np.zeros(len(T), np.int32), 'lam': None}

    from mseetc.sharding import solve_batch_sharded
    full = solve_batch_sharded(FakeSolver(), T, overrides={'mass': np.arange(len(T), dtype=float)}, partition='tiles')
    if rank == 0:
        d = {k: v.tolist() for k, v in res.items()}
        d['tiled_cost'] = tiled['cost'].tolist(); d['tiled_rank'] = tiled['rank'].tolist()
        d['full_z'] = full['z'].tolist(); d['full_iters'] = full['iters'].tolist(); d['full_counts'] = full['instances_per_rank']
        out.put(d)
    else:
        assert full is None and tiled is None
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
    # tile-dealt partition: caller's order restored, 32-instance tiles alternate between the ranks
    T = 1000.0 + np.arange(1001)
    assert np.array_equal(np.array(res['tiled_cost']), T * 2.0)
    assert np.array_equal(np.array(res['tiled_rank']), (np.arange(1001) // 32) % 2)
    # solve_batch_sharded: per-instance overrides follow their instances, results come back in the caller's order
    z = np.array(res['full_z'])
    assert np.array_equal(z[:, 0], T) and np.array_equal(z[:, 1], T + np.arange(1001))
    assert np.array_equal(np.array(res['full_iters']), (np.arange(1001) // 32) % 2) and sum(res['full_counts']) == 1001


def test_shard_tiles_balance_a_sorted_sweep():
    from mseetc.sharding import shard_tiles
    n = 4096 * 8
    feasible = np.arange(n) >= n // 2                    # sorted trip-time sweep: the infeasible half comes first
    parts = shard_tiles(n, 8)
    assert sorted(np.concatenate(parts).tolist()) == list(range(n))
    work = np.array([feasible[p].sum() for p in parts])
    assert work.max() == work.min() == n // 16
