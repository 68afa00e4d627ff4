"""Probe: where a step of the bench workload goes.  (a) the time-optimal presolve alone, (b) the sweep alone with the minimum
time known up front (screening at the first KKT evaluation), (c) the sweep with all instances feasible-looking (no screening)."""
import os, sys, time
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
from mseetc import _cabi

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
streams = int(sys.argv[2]) if len(sys.argv) > 2 else 1
train = Train(config={'id': 'NL_Intercity_VIRM6'})
solver = casadiSolver(train, Track(config={'id': 'CH_StGallen_Wil'}), bench.OPTS)
T = bench.sweep_times(n)
dev = torch.device('cuda', 0)
N = bench.N_INT


def show(tag, handles, wall):
    tot = {}
    for hh in handles:
        for k, v in _cabi.last_profile(hh).items():
            t = tot.setdefault(k, [0.0, 0])
            t[0] += v['ms']; t[1] += v['launches']
    print('%-30s wall %.2f ms | sum %.2f | ' % (tag, wall, sum(v[0] for v in tot.values())) +
          ' '.join('%s %.2f/%d' % (k, v[0], v[1]) for k, v in tot.items()), flush=True)


ts = solver._time_sibling()
ht = ts._ensure_handle()
_cabi.set_profiling(ht, True)
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    dur, st = solver.minimum_time()
    torch.cuda.synchronize(); w = 1e3 * (time.perf_counter() - t0)
show('presolve alone (public API)', [ht], w)
tmin = float(dur[0])
print('tmin', tmin, 'status', st)

zero = np.zeros(n)
P, M = solver._planes(n, T, zero, zero + 1.0, zero + 1.0, {}, (1 - train.etaTraction) / train.etaTraction, 1 - train.etaRgBrake)
perm, parts = _cabi.StreamPool.interleave(n, streams) if streams > 1 else (np.arange(n), None)
P = np.ascontiguousarray(P[:, perm])
ds, c0, bmax = solver._tables(solver._base['rho'], solver._base['g'], solver._base['velocityMax'])
up = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(device=dev, dtype=dt)
args = (up(P, torch.float64), up(np.full(n, N, np.int32), torch.int32), up(np.zeros(n, np.int32), torch.int32),
        up(np.array([0, N], np.int32), torch.int32), up(ds, torch.float64), up(c0, torch.float64), up(bmax, torch.float64))
solver.streams = streams
pool = _cabi.StreamPool(solver._make_handle, max(1, streams), dev)
tm_known = torch.full((n,), tmin, dtype=torch.float64, device=dev)
for tag, tm in (('sweep, tmin known', tm_known), ('sweep, no screening', None)):
    for prof in (True, False):
        for hh in pool.handles:
            _cabi.set_profiling(hh, prof)
        for rep in range(3):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            out = pool.solve(*args, tmin=tm, parts=parts)
            torch.cuda.synchronize(); w = 1e3 * (time.perf_counter() - t0)
        if prof:
            show(tag + ' (events on)', pool.handles, w)
        else:
            print('%-30s wall %.2f ms   ticks %s  ok %d' % (tag + ' (events off)', w, out.get('ticks'), int((out['status'] == 0).sum().item())), flush=True)
