"""Multi-GPU sharding of independent OCP instances (one process per GPU, `torch.distributed` for plumbing).

Instances are independent NLPs (reference ocp.py:310-409): every rank solves its share of ONE batch and there is NO
collective on the solve path.  Only the results and statistics are gathered afterwards on rank 0 (device tensors over NCCL /
NVLink on the GPU box followed by a single device-to-host copy; pickled host arrays over gloo in the CPU tests).

Two partitions:
  shard_ranges   contiguous [start, stop) per rank, balancing sum(N_i + 1) -- for batches in random order (Monte Carlo,
                 synthetic tracks with mixed interval counts)
  shard_tiles    32-instance tiles dealt round-robin to the ranks -- for SORTED batches such as a trip-time sweep, whose
                 infeasible instances (cheap: screened without iterating) all sit at one end
"""
import numpy as np


def shard_ranges(intervals_per_instance, world_size):
    """Contiguous [start, stop) per rank, balancing sum(N_i + 1).  Every instance is assigned exactly once."""
    w = np.asarray(intervals_per_instance, dtype=np.int64) + 1
    n = len(w)
    cum = np.concatenate([[0], np.cumsum(w)])
    total = cum[-1]
    cuts = [0]
    for r in range(1, world_size):
        target = total * r / world_size
        j = int(np.searchsorted(cum, target, side='left'))
        if j > 0 and abs(cum[j - 1] - target) <= abs(cum[min(j, n)] - target):
            j -= 1
        cuts.append(min(max(j, cuts[-1]), n))
    cuts.append(n)
    return [(cuts[r], cuts[r + 1]) for r in range(world_size)]


def shard_tiles(n_instances, world_size, tile=32):
    """Index arrays (ascending) per rank: tile j of `tile` consecutive instances goes to rank j % world_size."""
    idx = np.arange(n_instances)
    owner = (idx // tile) % world_size
    return [idx[owner == r] for r in range(world_size)]


def shard_indices(n_instances, world_size, partition='ranges', intervals_per_instance=None):
    if partition == 'tiles':
        return shard_tiles(n_instances, world_size)
    if intervals_per_instance is None:
        intervals_per_instance = np.ones(n_instances, dtype=np.int64)
    return [np.arange(a, b) for a, b in shard_ranges(intervals_per_instance, world_size)]


def _world(group):
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized():
        return 1, 0, None
    return dist.get_world_size(group), dist.get_rank(group), dist.get_backend(group)


def solve_sharded(solve_fn, n_instances, intervals_per_instance=None, group=None, partition='ranges'):
    """Run `solve_fn(indices) -> dict of numpy arrays (first axis = instance)` on this rank's share and gather the per-rank
    results on rank 0 as host objects (returns the dict in the caller's instance order there, None elsewhere)."""
    import torch.distributed as dist
    world, rank, _ = _world(group)
    parts = shard_indices(n_instances, world, partition, intervals_per_instance)
    local = solve_fn(parts[rank]) if len(parts[rank]) else {}
    if world == 1:
        return local
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(local, gathered, dst=0, group=group)
    if rank != 0:
        return None
    keys = list(next(p for p in gathered if p).keys())
    order = np.argsort(np.concatenate([parts[r] for r in range(world) if len(parts[r])]), kind='stable')
    return {k: np.concatenate([np.asarray(p[k]) for p in gathered if p])[order] for k in keys}


RESULT_KEYS = ('z', 'lam', 'obj', 'kkt', 'iters', 'status')


def solve_batch_sharded(solver, terminalTime, initialTime=0, terminalVelocity=1, initialVelocity=1, overrides=None,
                        partition='ranges', group=None, **kw):
    """`casadiSolver.solve_batch` of ONE batch spread over the ranks of the process group.

    Every rank calls this with the same (full-length) arguments and the solver of its own GPU; it solves its share
    (`partition`: 'ranges' or 'tiles', see the module docstring) and rank 0 returns the results of the whole batch in the
    caller's order (numpy arrays, as solve_batch does); the other ranks return None.  With the nccl backend the per-rank
    device results are gathered over NVLink and copied to the host once; nothing is exchanged while the ranks solve."""
    import torch
    import torch.distributed as dist
    world, rank, backend = _world(group)
    arrs = [np.atleast_1d(np.asarray(a, dtype=float)) for a in (terminalTime, initialTime, terminalVelocity, initialVelocity)]
    overrides = dict(overrides or {})
    n = max([len(a) for a in arrs] + [len(np.atleast_1d(v)) for v in overrides.values()])
    if world == 1:
        res = solver.solve_batch(terminalTime, initialTime, terminalVelocity, initialVelocity, overrides=overrides, **kw)
        res['instances_per_rank'] = [n]
        return res
    parts = shard_indices(n, world, partition, np.full(n, solver.numIntervals))
    idx = parts[rank]
    pick = lambda a: (a if np.ndim(a) == 0 or len(np.atleast_1d(a)) == 1 else np.asarray(a)[idx])
    if 'screen' in kw and not isinstance(kw['screen'], (bool, np.bool_)):
        kw = dict(kw, screen=np.broadcast_to(np.asarray(kw['screen'], dtype=bool), (n,))[idx])
    use_nccl = backend == 'nccl'
    local = solver.solve_batch(*[pick(a) for a in (terminalTime, initialTime, terminalVelocity, initialVelocity)],
                               overrides={k: pick(v) for k, v in overrides.items()}, to_host=not use_nccl, **kw)
    counts = [len(p) for p in parts]
    if not use_nccl:
        payload = {k: np.asarray(local[k]) for k in RESULT_KEYS if local.get(k) is not None}
        gathered = [None] * world if rank == 0 else None
        dist.gather_object(payload, gathered, dst=0, group=group)
        if rank != 0:
            return None
        order = np.argsort(np.concatenate(parts), kind='stable')
        res = {k: np.concatenate([g[k] for g in gathered])[order] for k in gathered[0]}
    else:
        cmax = max(counts)
        res = {}
        order = torch.from_numpy(np.argsort(np.concatenate(parts), kind='stable')) if rank == 0 else None
        for k in RESULT_KEYS:
            t = local.get(k)
            if t is None:
                continue
            pad = t
            if t.shape[0] < cmax:                      # dist.gather needs equal shapes
                pad = torch.zeros((cmax,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
                pad[:t.shape[0]] = t
            pad = pad.contiguous()
            bufs = [torch.empty_like(pad) for _ in range(world)] if rank == 0 else None
            dist.gather(pad, bufs, dst=0, group=group)
            if rank == 0:
                full = torch.cat([b[:c] for b, c in zip(bufs, counts)], dim=0)
                full = full.index_select(0, order.to(full.device))
                host = solver._pinned('sharded_' + k, full)
                host.copy_(full, non_blocking=True)
                res[k] = host
        if rank != 0:
            return None
        torch.cuda.current_stream().synchronize()
        res = {k: v.numpy() for k, v in res.items()}
    res['instances_per_rank'] = counts
    return res
