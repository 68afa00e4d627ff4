"""Probe: duration of the Riccati-sweep kernel (k_step) vs number of instances / active lanes (all instances feasible)."""
import os, sys
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
solver = casadiSolver(Train(config={'id': 'NL_Intercity_VIRM6'}), Track(config={'id': 'CH_StGallen_Wil'}), opts)
h = solver._ensure_handle()
_cabi.set_profiling(h, True)
for n in (1, 8, 32, 33, 512, 4096, 16384):
    T = np.linspace(1100.0, 1400.0, n)
    for rep in range(2):
        res = solver.solve_batch(T, screen=False)
    p = _cabi.last_profile(h)
    print('n=%6d  ok=%d  iters max %d | ' % (n, int((res['status'] == 0).sum()), res['iters'].max()) +
          ' '.join('%s %.0fus' % (k, 1e3 * v['ms'] / max(1, v['launches'])) for k, v in p.items() if k != 'misc'), flush=True)
