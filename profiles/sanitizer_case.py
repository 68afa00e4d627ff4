import sys
sys.path[:0] = ['/root/repo', '/root/repo/ms-eetc_b200']
import numpy as np
import __graft_entry__ as ge
ge.build()
from mseetc.ocp import casadiSolver
from mseetc.train import Train
from mseetc.track import Track
opts = {'numIntervals': 60, 'maxIterations': 200, 'integrationMethod': 'RK', 'integrationOptions': {'order': 4, 'numSteps': 1, 'numApproxSteps': 1}}
s = casadiSolver(Train(config={'id': 'NL_Intercity_VIRM6'}), Track(config={'id': '00_var_speed_limit_100'}), opts)
T = np.linspace(1300.0, 1700.0, 70)
r = s.solve_batch(T)
print('status', np.bincount(r['status'], minlength=6), 'iters max', r['iters'].max())
