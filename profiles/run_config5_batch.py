"""BASELINE configs[4] at full size through the batch preprocessing path: 16 384 random tracks (rng 11), mixed interval counts,
spline loss map, pn brake off, T = 1.15 Tmin_i -- construction, presolve and solve times (one JSON line)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'ms-eetc_b200')]
import numpy as np
import __graft_entry__ as ge
ge.build()
import torch
from mseetc.train import Train
from mseetc.efficiency import totalLossesFunction
from mseetc.trackbatch import TrackBatch, solve_tracks
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
train = Train(config={'id': 'NL_Intercity_VIRM6'}); train.forceMinPn = 0
train.powerLosses = totalLossesFunction(train, auxiliaries=27000, etaGear=0.96)
opts = {'maxIterations': 500, 'integrationMethod': 'RK', 'integrationOptions': {'order': 4, 'numSteps': 1, 'numApproxSteps': 1}}
for rep in range(2):
    rng = np.random.default_rng(11)
    t0 = time.perf_counter()
    batch = TrackBatch.random(rng, n)
    N = rng.choice([100, 200, 300, 400], n)
    order = np.argsort(N, kind='stable')                 # equal interval counts share tiles
    batch, N = batch.subset(order), N[order]
    t_gen = time.perf_counter() - t0
    res = solve_tracks(train, batch, N, opts, timeFactor=1.15)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
ok = (res['status'] == 0) | (res['status'] == 6)
feas = (res['tmin'] > 0) & (res['grid_error'] == 0)
print(json.dumps(dict(n=n, generation_s=t_gen, timing=res['timing'], wall_s=wall, grid_errors=int((res['grid_error'] != 0).sum()),
                      time_optimal_converged=int(feas.sum()), energy_converged=int((ok & feas).sum()), acceptable=int((res['status'] == 6).sum()),
                      failed_of_feasible=int((feas & ~ok).sum()), status_histogram={int(k): int((res['status'] == k).sum()) for k in np.unique(res['status'])},
                      restarted=int(len(res.get('restarted', []))), iters_mean=float(res['iters'][ok & feas].mean()), solves_per_s=n / wall)))
