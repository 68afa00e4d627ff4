"""Probe: parallel-in-time sweeps (8 / 16 / 32 chunk lanes per instance) vs sequential sweeps -- accuracy, fallbacks, kernel time.
Run on the GPU box: python profiles/probe_pit.py [n ...]"""
import os, sys, time, json
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
sizes = [int(a) for a in sys.argv[1:]] or [32, 512, 2048, 4096]
rows = []
for n in sizes:
    T = np.linspace(1040.0, 1240.0, n)
    ref = None
    for lanes in [int(x) for x in os.environ.get('PROBE_LANES', '1,8,16,32').split(',')]:
        s = casadiSolver(train, Track(config={'id': 'CH_StGallen_Wil'}), opts)
        s.streams = 1; s.sweepLanes = lanes
        h = s._ensure_handle(); _cabi.set_profiling(h, True)
        for rep in range(3):
            t0 = time.perf_counter(); res = s.solve_batch(T, screen=False); dt = time.perf_counter() - t0
        prof = _cabi.last_profile(h)
        p = prof['inst_step']
        tot = sum(v['ms'] for v in prof.values())
        if ref is None: ref = {k: np.array(v) for k, v in res.items() if isinstance(v, np.ndarray)}
        row = dict(n=n, lanes=lanes, converged=int((res['status'] == 0).sum()), iters_mean=float(res['iters'].mean()), iters_max=int(res['iters'].max()),
                   same_iters=bool(np.array_equal(res['iters'], ref['iters'])), k_step_us=1e3 * p['ms'] / max(1, p['launches']), launches=p['launches'],
                   device_ms=tot, wall_ms=1e3 * dt, fallbacks=int(h.last_sweep_fallbacks()), fallback_reasons=h.last_sweep_fallback_reasons(),
                   obj_rel=float(np.max(np.abs(res['obj'] - ref['obj']) / np.abs(ref['obj']))), dz=float(np.abs(res['z'] - ref['z']).max()))
        rows.append(row)
        print(json.dumps(row), flush=True)
