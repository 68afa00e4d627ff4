"""Probe: iterations of the time-optimal presolve (casadiSolver.minimum_time) vs mu_init, on the bundled tracks and random tracks."""
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
from mseetc.synthetic import random_track

train = Train(config={'id': 'NL_Intercity_VIRM6'})
cases = [('swiss300', Track(config={'id': 'CH_StGallen_Wil'}), 300), ('swiss100', Track(config={'id': 'CH_StGallen_Wil'}), 100),
         ('flat200', Track(config={'id': '00_var_speed_limit_100'}), 200)]
rng = np.random.default_rng(5)
for i in range(6):
    cases.append(('rand%d' % i, random_track(rng), int(rng.integers(100, 400))))
for mu in (0.1, 1e-2, 1e-3, 1e-4, 1e-5, 1e-6):
    line = []
    for name, track, N in cases:
        o = dict(bench.OPTS); o['numIntervals'] = N
        try:
            solver = casadiSolver(train, track, o)
            ts = solver._time_sibling(); ts._handle = None; ts.muInit = mu
            ref = ts.solve_batch
            it = []
            def spy(*a, **k):
                r = ref(*a, **k); it.append(int(r['iters'][0])); return r
            ts.solve_batch = spy
            dur, st = solver.minimum_time()
            line.append('%s %d/%s/%.3f' % (name, st[0], '+'.join(map(str, it)), dur[0]))
        except Exception as e:
            line.append('%s ERR %s' % (name, str(e)[:40]))
    print('mu %.0e | ' % mu + ' | '.join(line), flush=True)
