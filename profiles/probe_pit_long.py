"""Probe: long horizon (BASELINE config 4) with sequential vs parallel-in-time sweeps."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'ms-eetc_b200')]
import numpy as np, torch
import __graft_entry__ as ge
ge.build()
from mseetc.ocp import casadiSolver
from mseetc.train import Train
from mseetc.synthetic import random_track
train = Train(config={'id': 'NL_Intercity_VIRM6'})
for N in (2000, 20000):
    rng = np.random.default_rng(7)
    track = random_track(rng, length=200e3, title='synthetic 200 km')
    o = {'numIntervals': N, 'maxIterations': 1000, 'integrationMethod': 'RK', 'integrationOptions': {'order': 4, 'numSteps': 1, 'numApproxSteps': 1}}
    ref = None
    for lanes in (1, 8, 32):
        ts = casadiSolver(train, track, dict(o, energyOptimal=False)); ts.sweepLanes = lanes
        lim = np.minimum(ts.points['Speed limit [m/s]'].values[:-1], ts._base['velocityMax'])
        ts.solve_batch(1.5 * float(np.sum(ts.steps / lim)), screen=False)
        torch.cuda.synchronize(); t = time.perf_counter(); tres = ts.solve_batch(1.5 * float(np.sum(ts.steps / lim)), screen=False); torch.cuda.synchronize(); dtt = time.perf_counter() - t
        tmin = float(tres['z'][0][-2])
        es = casadiSolver(train, track, o); es.sweepLanes = lanes
        es.solve_batch(1.10 * 7490.8, screen=False)
        torch.cuda.synchronize(); t = time.perf_counter(); res = es.solve_batch(1.10 * 7490.8, screen=False); torch.cuda.synchronize(); dte = time.perf_counter() - t
        if ref is None: ref = res
        print('N=%5d lanes=%2d  time-opt: st %d it %d %.3fs tmin %.4f fb %d | energy: st %d it %d %.3fs cost %.6f kkt %.1e fb %d dz %.1e' % (
            N, lanes, tres['status'][0], tres['iters'][0], dtt, tmin, ts._ensure_handle().last_sweep_fallbacks(), res['status'][0], res['iters'][0], dte,
            res['cost'][0], res['kkt'][0], es._ensure_handle().last_sweep_fallbacks(), np.abs(res['z'] - ref['z']).max()), flush=True)
