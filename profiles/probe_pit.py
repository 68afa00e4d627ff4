"""Probe: parallel-in-time sweeps (8 / 32 lanes per instance) vs sequential sweeps -- accuracy, fallbacks, kernel time."""
import os, sys, time
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
for n in (1, 512, 4096):
    T = np.linspace(1040.0, 1400.0, n) if n > 1 else np.array([1242.0])
    ref = None
    for lanes in (1, 8, 32):
        s = casadiSolver(train, Track(config={'id': 'CH_StGallen_Wil'}), opts)
        s.streams = 1; s.sweepLanes = lanes
        h = s._ensure_handle(); _cabi.set_profiling(h, True)
        for rep in range(2):
            t0 = time.perf_counter(); res = s.solve_batch(T, screen=False); dt = time.perf_counter() - t0
        p = _cabi.last_profile(h)['inst_step']
        if ref is None: ref = res
        print('n=%5d lanes=%2d ok=%d iters mean %.1f max %d  k_step %.0f us/launch  wall %.1f ms  fallbacks %d  |obj-ref| rel %.1e  dz %.1e' % (
            n, lanes, int((res['status'] == 0).sum()), res['iters'].mean(), res['iters'].max(), 1e3 * p['ms'] / max(1, p['launches']), 1e3 * dt,
            h.last_sweep_fallbacks(), np.max(np.abs(res['obj'] - ref['obj']) / np.abs(ref['obj'])), np.abs(res['z'] - ref['z']).max()), flush=True)
