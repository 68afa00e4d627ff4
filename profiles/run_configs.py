"""BASELINE.json configs 3, 4, 5 (SURVEY.md section 8d recipes) through the public API -- prints one JSON line per config.
Sizes can be scaled down with --scale (fraction of the nominal instance counts)."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'ms-eetc_b200')]
import numpy as np
import __graft_entry__ as ge
ge.build()
import torch
from mseetc.ocp import casadiSolver, solve_instances
from mseetc.train import Train
from mseetc.track import Track
from mseetc.efficiency import totalLossesFunction
from mseetc.synthetic import random_track

OPTS = {'numIntervals': 300, 'maxIterations': 500, 'integrationMethod': 'RK', 'integrationOptions': {'order': 4, 'numSteps': 1, 'numApproxSteps': 1}}


def timed(fn):
    torch.cuda.synchronize(); t = time.perf_counter(); r = fn(); torch.cuda.synchronize(); return r, time.perf_counter() - t


def config3(n):
    "parameter Monte Carlo: randomised mass, Davis coefficients, efficiencies; half the batch on the dynamic loss map"
    rng = np.random.default_rng(20260101)
    half = n // 2
    train = Train(config={'id': 'NL_Intercity_VIRM6'})
    track = Track(config={'id': '00_var_speed_limit_100'})
    solver = casadiSolver(train, track, OPTS)
    ov = dict(mass=391000 * rng.uniform(0.85, 1.15, half), r0=train.r0 * rng.uniform(0.8, 1.2, half), r1=train.r1 * rng.uniform(0.8, 1.2, half),
              r2=train.r2 * rng.uniform(0.8, 1.2, half), etaTraction=rng.uniform(0.80, 0.92, half), etaRgBrake=rng.uniform(0.55, 0.85, half))
    solver.solve_batch(1541.0, overrides={k: v[:64] for k, v in ov.items()})     # warm-up
    res, dt = timed(lambda: solver.solve_batch(1541.0, overrides=ov))
    out = dict(static=dict(n=half, wall_s=dt, solves_per_s=half / dt, converged=int((res['status'] == 0).sum()), infeasible=int((res['status'] == 4).sum()),
                           other=int(((res['status'] != 0) & (res['status'] != 4)).sum()), iters_mean=float(res['iters'][res['status'] == 0].mean()),
                           kkt_max=float(res['kkt'][res['status'] == 0].max())))
    dtrain = Train(config={'id': 'NL_Intercity_VIRM6'})
    dtrain.forceMinPn = 0
    dtrain.powerLosses = totalLossesFunction(dtrain, auxiliaries=27000, etaGear=0.96)
    dsolver = casadiSolver(dtrain, track, dict(OPTS, minimumVelocity=1))
    ov2 = dict(mass=391000 * rng.uniform(0.85, 1.15, half), r0=dtrain.r0 * rng.uniform(0.8, 1.2, half), r1=dtrain.r1 * rng.uniform(0.8, 1.2, half),
               r2=dtrain.r2 * rng.uniform(0.8, 1.2, half), auxiliaries=rng.uniform(20e3, 35e3, half), tableScale=rng.uniform(0.9, 1.1, half))
    dsolver.solve_batch(1541.0, overrides={k: v[:64] for k, v in ov2.items()})
    res, dt = timed(lambda: dsolver.solve_batch(1541.0, overrides=ov2))
    out['dynamic'] = dict(n=half, wall_s=dt, solves_per_s=half / dt, converged=int((res['status'] == 0).sum()), infeasible=int((res['status'] == 4).sum()),
                          other=int(((res['status'] != 0) & (res['status'] != 4)).sum()), iters_mean=float(res['iters'][res['status'] == 0].mean()),
                          kkt_max=float(res['kkt'][res['status'] == 0].max()))
    return out


def config5(n):
    "synthetic track batch: random profiles, mixed interval counts, dynamic loss map, pn brake off, T = 1.15 Tmin_i"
    rng = np.random.default_rng(11)
    train = Train(config={'id': 'NL_Intercity_VIRM6'})
    train.forceMinPn = 0
    train.powerLosses = totalLossesFunction(train, auxiliaries=27000, etaGear=0.96)
    t0 = time.perf_counter()
    solvers, tsolvers = [], []
    while len(solvers) < n:
        N = int(rng.choice([100, 200, 300, 400]))
        track = random_track(rng)
        try:
            o = dict(OPTS, numIntervals=N, minimumVelocity=1)
            solvers.append(casadiSolver(train, track, o))
            tsolvers.append(casadiSolver(train, track, dict(o, energyOptimal=False)))
        except ValueError:
            continue                       # more track sections than intervals (track.py:103-105): draw again
    build_s = time.perf_counter() - t0
    order = np.argsort([s.numIntervals for s in solvers], kind='stable')      # equal interval counts share warps
    solvers = [solvers[i] for i in order]; tsolvers = [tsolvers[i] for i in order]
    lim = [np.minimum(s.points['Speed limit [m/s]'].values[:-1], s._base['velocityMax']) for s in solvers]
    horizon = 1.5 * np.array([float(np.sum(s.steps / l)) for s, l in zip(solvers, lim)])
    tres, dt_t = timed(lambda: solve_instances(tsolvers, horizon, screen=False))
    nint = np.array([s.numIntervals for s in solvers])
    stp = 4
    tmin = tres['z'][np.arange(n), nint * stp]
    okT = tres['status'] == 0
    T = np.where(okT, 1.15 * tmin, horizon)
    res, dt = timed(lambda: solve_instances(solvers, T, screen=False))
    ok = (res['status'] == 0) & okT
    return dict(n=n, host_construction_s=build_s, time_optimal=dict(wall_s=dt_t, converged=int(okT.sum()), iters_mean=float(tres['iters'][okT].mean())),
                energy=dict(wall_s=dt, solves_per_s=n / dt, converged=int(ok.sum()), failed_of_feasible=int((okT & ~ok).sum()),
                            iters_mean=float(res['iters'][ok].mean()), kkt_max=float(res['kkt'][ok].max())),
                interval_counts={int(k): int((nint == k).sum()) for k in (100, 200, 300, 400)})


def config4(N):
    "long horizon: one instance on a synthetic 200 km track, T = 1.10 Tmin"
    rng = np.random.default_rng(7)
    train = Train(config={'id': 'NL_Intercity_VIRM6'})
    track = random_track(rng, length=200e3, title='synthetic 200 km')
    o = dict(OPTS, numIntervals=N, maxIterations=1000)
    ts = casadiSolver(train, track, dict(o, energyOptimal=False))
    lim = np.minimum(ts.points['Speed limit [m/s]'].values[:-1], ts._base['velocityMax'])
    tres, dt_t = timed(lambda: ts.solve_batch(1.5 * float(np.sum(ts.steps / lim)), screen=False))
    out = dict(N=N, time_optimal=dict(status=int(tres['status'][0]), iters=int(tres['iters'][0]), wall_s=dt_t, ticks=int(tres['ticks'])))
    if tres['status'][0] == 0:
        tmin = float(tres['z'][0][-2])
        es = casadiSolver(train, track, o)
        res, dt = timed(lambda: es.solve_batch(1.10 * tmin, screen=False))
        out['tmin_s'] = tmin
        out['energy'] = dict(status=int(res['status'][0]), iters=int(res['iters'][0]), wall_s=dt, ticks=int(res['ticks']), cost_kwh=float(res['cost'][0]),
                             kkt=float(res['kkt'][0]))
    return out


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--scale', type=float, default=1.0)
    ap.add_argument('--which', default='3,5,4')
    ap.add_argument('--n4', type=int, default=20000)
    a = ap.parse_args()
    for w in a.which.split(','):
        if w == '3':
            print(json.dumps({'config': 3, 'result': config3(int(65536 * a.scale))}), flush=True)
        if w == '5':
            print(json.dumps({'config': 5, 'result': config5(int(16384 * a.scale))}), flush=True)
        if w == '4':
            for N in sorted({2000, a.n4}):
                print(json.dumps({'config': 4, 'result': config4(N)}), flush=True)
