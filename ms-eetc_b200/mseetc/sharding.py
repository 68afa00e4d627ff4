"""Multi-GPU sharding of independent OCP instances (one process per GPU, `torch.distributed` for plumbing).

Instances are independent NLPs (reference ocp.py:310-409): every rank solves a contiguous range, chosen so that
the number of shooting intervals (work ~ sum of N_i) is balanced; there is NO collective on the solve path.
Only the results and statistics are gathered afterwards (host side, any backend: nccl on the GPU box, gloo in the
CPU tests).
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


def solve_sharded(solve_fn, n_instances, intervals_per_instance=None, group=None):
    """Run `solve_fn(start, stop) -> dict of numpy arrays (first axis = instance)` on this rank's range and gather
    the per-rank results on rank 0 (returns the concatenated dict there, None elsewhere)."""
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if intervals_per_instance is None:
        intervals_per_instance = np.ones(n_instances, dtype=np.int64)
    start, stop = shard_ranges(intervals_per_instance, world)[rank]
    local = solve_fn(start, stop) if stop > start else {}
    if world == 1:
        return local
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(local, gathered, dst=0, group=group)
    if rank != 0:
        return None
    keys = list(next(p for p in gathered if p).keys())
    return {k: np.concatenate([np.asarray(p[k]) for p in gathered if p]) for k in keys}
