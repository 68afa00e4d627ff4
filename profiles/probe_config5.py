import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'ms-eetc_b200'), os.path.join(ROOT, 'profiles')]
import numpy as np, torch
import run_configs as rc
from mseetc.ocp import casadiSolver, solve_instances
from mseetc.train import Train
from mseetc.efficiency import totalLossesFunction
from mseetc.synthetic import random_track
rng = np.random.default_rng(11)
train = Train(config={'id': 'NL_Intercity_VIRM6'}); train.forceMinPn = 0
train.powerLosses = totalLossesFunction(train, auxiliaries=27000, etaGear=0.96)
solvers, tsolvers = [], []
while len(solvers) < 1024:
    N = int(rng.choice([100, 200, 300, 400])); track = random_track(rng)
    try:
        o = dict(rc.OPTS, numIntervals=N, minimumVelocity=1)
        solvers.append(casadiSolver(train, track, o)); tsolvers.append(casadiSolver(train, track, dict(o, energyOptimal=False)))
    except ValueError:
        continue
lim = [np.minimum(s.points['Speed limit [m/s]'].values[:-1], s._base['velocityMax']) for s in solvers]
horizon = 1.5 * np.array([float(np.sum(s.steps / l)) for s, l in zip(solvers, lim)])
for rep in range(2):
    t = time.perf_counter(); tres = solve_instances(tsolvers, horizon, screen=False); torch.cuda.synchronize(); dt = time.perf_counter() - t
    print('time-opt wall %.3f ticks %d launches %d status hist %s iters max %d' % (dt, tres['ticks'], tres['launches'], np.bincount(tres['status'], minlength=6).tolist(), tres['iters'].max()))
nint = np.array([s.numIntervals for s in solvers]); tmin = tres['z'][np.arange(len(solvers)), nint * 4]
T = np.where(tres['status'] == 0, 1.15 * tmin, horizon)
for rep in range(2):
    t = time.perf_counter(); res = solve_instances(solvers, T, screen=False); torch.cuda.synchronize(); dt = time.perf_counter() - t
    print('energy   wall %.3f ticks %d launches %d status hist %s iters max %d' % (dt, res['ticks'], res['launches'], np.bincount(res['status'], minlength=6).tolist(), res['iters'].max()))
bad = np.where((res['status'] != 0) & (tres['status'] == 0))[0]
print('failed energy: iters', res['iters'][bad][:20].tolist(), 'kkt', ['%.1e' % x for x in res['kkt'][bad][:20]])
tbad = np.where(tres['status'] != 0)[0]
print('failed time-opt idx', tbad.tolist(), 'N', nint[tbad].tolist(), 'iters', tres['iters'][tbad].tolist(), 'kkt', ['%.1e' % x for x in tres['kkt'][tbad]])
print('failed energy idx', bad.tolist(), 'status', res['status'][bad].tolist(), 'N', nint[bad].tolist())
