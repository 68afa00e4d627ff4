"""Probe: duration of every launch of one sweep (2048 feasible instances, one stream, per-kernel events) against the number of
instances still running -- what the lock-step tail costs, with and without compaction of the running batch."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'ms-eetc_b200')]
import numpy as np
import __graft_entry__ as ge
ge.build()
from mseetc.ocp import casadiSolver
from mseetc.train import Train
from mseetc.track import Track
from mseetc import _cabi
opts = {'numIntervals': 300, 'maxIterations': 500, 'integrationMethod': 'RK', 'integrationOptions': {'order': 4, 'numSteps': 1, 'numApproxSteps': 1}}
train = Train(config={'id': 'NL_Intercity_VIRM6'})
T = np.linspace(1036.0, 1243.0, 2048)
for compaction in (False, True):
    s = casadiSolver(train, Track(config={'id': 'CH_StGallen_Wil'}), opts)
    h = s._ensure_handle(); h.set_compaction(compaction); _cabi.set_profiling(h, True)
    for rep in range(2):
        res = s.solve_batch(T, screen=False)
    tl = _cabi.last_timeline(h)
    names = _cabi.KERNEL_CLASSES
    iters = res['iters']
    rows = {}
    tick = {n: 0 for n in names}
    for cls, a, b in tl:
        n = names[int(cls)]
        rows.setdefault(n, []).append(1e3 * (b - a))
    running = [int((iters > t).sum()) for t in range(int(iters.max()) + 1)]
    print(json.dumps(dict(compaction=compaction, compactions=list(h.last_compactions()), total_ms=float(tl[-1][2] - tl[0][1]), running_per_iteration=running,
                          trial_eval_us=[round(x) for x in rows['cell_trial']], cell_step_us=[round(x) for x in rows['cell_step']],
                          sweep_us=[round(x) for x in rows['inst_step']])))
