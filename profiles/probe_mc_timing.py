"""Probe: wall-time breakdown of a large parameter Monte Carlo through solve_batch (first and repeated call)."""
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
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
rng = np.random.default_rng(20260101)
train = Train(config={'id': 'NL_Intercity_VIRM6'})
solver = casadiSolver(train, Track(config={'id': '00_var_speed_limit_100'}), bench.OPTS)
ov = dict(mass=391000 * rng.uniform(0.85, 1.15, n), r0=train.r0 * rng.uniform(0.8, 1.2, n), r1=train.r1 * rng.uniform(0.8, 1.2, n),
          r2=train.r2 * rng.uniform(0.8, 1.2, n), etaTraction=rng.uniform(0.80, 0.92, n), etaRgBrake=rng.uniform(0.55, 0.85, n))
for rep in range(3):
    t = time.perf_counter()
    res = solver.solve_batch(1541.0, overrides=ov)
    w = time.perf_counter() - t
    print('call %d: wall %.3f s  ok %d  ' % (rep, w, int((res['status'] == 0).sum())) + ' '.join('%s %.1f' % (k, 1e3 * v) for k, v in res['timing'].items()), flush=True)
