"""Probe: IP iterations of the energy sweep and of the time-optimal presolve vs mu_init (starting profile = speed envelope)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'ms-eetc_b200')]
import numpy as np
import __graft_entry__ as ge
ge.build()
import bench
from mseetc.ocp import casadiSolver
from mseetc.train import Train
from mseetc.track import Track

train = Train(config={'id': 'NL_Intercity_VIRM6'})
T = np.linspace(1036.0, 1243.0, 512)
ref = None
for mu in (0.1, 0.03, 0.01, 3e-3, 1e-3, 1e-4):
    solver = casadiSolver(train, Track(config={'id': 'CH_StGallen_Wil'}), bench.OPTS)
    solver.muInit = mu
    solver.streams = 1
    res = solver.solve_batch(T, screen=False)
    ok = res['status'] == 0
    if ref is None:
        ref = res['cost'].copy()
    dev = np.max(np.abs(res['cost'][ok] - ref[ok]) / ref[ok]) if ok.any() else float('nan')
    ts = solver._time_sibling()
    ts._handle = None
    ts.muInit = mu
    dur, st = solver.minimum_time()
    tr = ts.solve_batch(1.5 * 1035.0, screen=False)
    print('mu_init %7.0e | energy: ok %d/%d iters mean %.1f max %d  cost dev %.1e | time-opt: status %s tmin %.6f iters %d' % (
        mu, ok.sum(), len(T), res['iters'][ok].mean(), res['iters'].max(), dev, st, dur[0], tr['iters'][0]), flush=True)
