"""Probe: IP iterations vs trip time relative to the minimum time (bench sweep, feasible half)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'ms-eetc_b200')]
import numpy as np
import __graft_entry__ as ge
ge.build()
import bench
from mseetc.ocp import casadiSolver
from mseetc.train import Train
from mseetc.track import Track
solver = casadiSolver(Train(config={'id': 'NL_Intercity_VIRM6'}), Track(config={'id': 'CH_StGallen_Wil'}), bench.OPTS)
T = bench.sweep_times(4096)
res = solver.solve_batch(T)
tmin = res['tmin'][0]
ok = res['status'] == 0
r = T / tmin
edges = [1.0, 1.002, 1.005, 1.01, 1.02, 1.04, 1.08, 1.12, 1.16, 1.2001]
for lo, hi in zip(edges[:-1], edges[1:]):
    m = ok & (r >= lo) & (r < hi)
    if m.any():
        print('T/Tmin in [%.3f, %.3f): n=%4d iters mean %.1f max %d' % (lo, hi, m.sum(), res['iters'][m].mean(), res['iters'][m].max()))
print('ticks-equivalent max iters', res['iters'][ok].max())
