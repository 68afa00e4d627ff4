"""Oracle optimum of an integrateLosses = True problem with the spline loss map (reference ocp.py:231-241, simulations/table3.py
train: pn brake off, totalLossesFunction(train, 27000, 0.96)) -> tests/golden/intlosses_dynamic_flat_N60.json.

ORACLE output (the reference formulation: rows on E(sqrt(b_i), t_{i+1} - t_i, Fel_i, Fpb_i), loss energies by RK4 in the time
domain -- 8 steps, 32 where the speed crosses a kink of the loss map --, torch autograd derivatives), not a reference output: CasADi + IPOPT + CVODES cannot run in this image.  ~5 minutes."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'ms-eetc_b200')]
from common import FLAT_JSON, fig5_train, oracle_nlp, oracle_solve       # noqa: E402
from oracle.problem import load_track                                   # noqa: E402

N, T, AUX, ETAG, STEPS = 60, 1541.0, 27000.0, 0.96, 8
tr = fig5_train(); tr.losses = ('dynamic', AUX, ETAG, 1.0)
nlp = oracle_nlp(tr, load_track(FLAT_JSON), N, integrateLosses=True, oracleLossSteps=STEPS)
t0 = time.time()
r = oracle_solve(nlp, T)
u = nlp.unpack(r.x)
out = dict(N=N, T=T, auxiliaries=AUX, etaGear=ETAG, oracle_loss_steps=STEPS, status=r.status, iterations=int(r.iters), kkt=float(r.kkt),
           objective=float(r.f), cost_kwh=float(nlp.cost(r.f)), t=u['t'].tolist(), b=u['b'].tolist(), Fel=u['Fel'].tolist(), s=u['s'].tolist(),
           seconds=round(time.time() - t0))
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'intlosses_dynamic_flat_N60.json'), 'w'))
print(out['status'], out['iterations'], out['cost_kwh'], out['seconds'])
