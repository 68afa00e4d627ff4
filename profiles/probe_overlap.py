"""Probe: device timeline of a k-stream solve (tmin known up front): which kernels of the streams overlap."""
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
streams = int(sys.argv[2]) if len(sys.argv) > 2 else 2
train = Train(config={'id': 'NL_Intercity_VIRM6'})
solver = casadiSolver(train, Track(config={'id': 'CH_StGallen_Wil'}), bench.OPTS)
T = bench.sweep_times(n)
dev = torch.device('cuda', 0)
N = bench.N_INT
zero = np.zeros(n)
P, M = solver._planes(n, T, zero, zero + 1.0, zero + 1.0, {}, (1 - train.etaTraction) / train.etaTraction, 1 - train.etaRgBrake)
perm, parts = _cabi.StreamPool.interleave(n, streams) if streams > 1 else (np.arange(n), None)
P = np.ascontiguousarray(P[:, perm])
ds, c0, bmax = solver._tables(solver._base['rho'], solver._base['g'], solver._base['velocityMax'])
up = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(device=dev, dtype=dt)
args = (up(P, torch.float64), up(np.full(n, N, np.int32), torch.int32), up(np.zeros(n, np.int32), torch.int32),
        up(np.array([0, N], np.int32), torch.int32), up(ds, torch.float64), up(c0, torch.float64), up(bmax, torch.float64))
pool = _cabi.StreamPool(solver._make_handle, max(1, streams), dev)
tm = torch.full((n,), 1035.5535722974237, dtype=torch.float64, device=dev)
for hh in pool.handles:
    _cabi.set_profiling(hh, True)
for rep in range(3):
    out = pool.solve(*args, tmin=tm, parts=parts)
torch.cuda.synchronize()
names = _cabi.KERNEL_CLASSES
rows = []
for i, hh in enumerate(pool.handles):
    for c, a, b in _cabi.last_timeline(hh, pool.handles[0]):
        rows.append((a, b, i, names[int(c)]))
rows.sort()
lo, hi = (float(sys.argv[3]), float(sys.argv[4])) if len(sys.argv) > 4 else (10.0, 12.5)
print('launches %d, last end %.2f ms; window %.1f..%.1f ms' % (len(rows), max(r[1] for r in rows), lo, hi))
for a, b, i, nm in rows:
    if lo <= a <= hi:
        print('%9.3f %9.3f  %6.0f us  s%d %s%s' % (a, b, 1e3 * (b - a), i, '    ' * i, nm))
