"""Small solve for compute-sanitizer runs (memcheck / racecheck / initcheck)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'ms-eetc_b200')]
import numpy as np
import __graft_entry__ as ge
ge.build()
from mseetc.ocp import casadiSolver
from mseetc.train import Train
from mseetc.track import Track
opts = {'numIntervals': 60, 'maxIterations': 200, 'integrationMethod': 'RK', 'integrationOptions': {'order': 4, 'numSteps': 1, 'numApproxSteps': 1}}
s = casadiSolver(Train(config={'id': 'NL_Intercity_VIRM6'}), Track(config={'id': '00_var_speed_limit_100'}), opts)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 70
T = np.linspace(1300.0, 1700.0, n) if n > 1 else np.array([1600.0])
r = s.solve_batch(T, screen=(n > 1))
print('status', np.bincount(r['status'], minlength=6), 'iters max', r['iters'].max())
