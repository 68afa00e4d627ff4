"""Probe: the 4096-instance trip-time sweep of the bench with the other formulations / integrators (one GPU, device-resident results)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'ms-eetc_b200')]
import numpy as np
import torch
import __graft_entry__ as ge
ge.build()
import bench
from mseetc.ocp import casadiSolver
from mseetc.train import Train
from mseetc.track import Track

train = Train(config={'id': 'NL_Intercity_VIRM6'})
track = Track(config={'id': 'CH_StGallen_Wil'})
n = 4096
for name, extra in (('RK (bench)', {}), ('integrateLosses', {'integrateLosses': True}),
                    ('IRK radau 2', {'integrationMethod': 'IRK', 'integrationOptions': {'order': 2, 'numApproxSteps': 1}}),
                    ('CVODES-equivalent', {'integrationMethod': 'CVODES', 'integrationOptions': {}})):
    opts = dict(bench.OPTS); opts.update(extra)
    solver = casadiSolver(train, track, opts)
    tmin = float(np.atleast_1d(solver.minimum_time()[0])[0])
    T = tmin * (0.8 + 0.4 * np.arange(n) / (n - 1))
    best = None
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        res = solver.solve_batch(T, to_host=False)
        torch.cuda.synchronize(); w = time.perf_counter() - t0
        best = w if best is None else min(best, w)
    st = res['status'].cpu().numpy() if hasattr(res['status'], 'cpu') else res['status']
    feas = T >= tmin
    print(json.dumps(dict(variant=name, tmin=float(tmin), ms=1e3 * best, feasible=int(feas.sum()), converged_feasible=int(((st == 0) | (st == 6))[feas].sum()),
                          flagged_infeasible=int((st == 4)[~feas].sum()), feasible_solves_per_s=float(feas.sum() / best))), flush=True)
