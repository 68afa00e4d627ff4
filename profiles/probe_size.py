"""Probe: per-launch kernel times of one handle on one stream vs number of (all feasible, all active) instances."""
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

train = Train(config={'id': 'NL_Intercity_VIRM6'})
solver = casadiSolver(train, Track(config={'id': 'CH_StGallen_Wil'}), bench.OPTS)
dev = torch.device('cuda', 0)
N = bench.N_INT
if os.environ.get('MSEETC_LANES'):
    solver.sweepLanes = int(os.environ['MSEETC_LANES'])      # 1 sequential, 8 / 16 / 32 parallel in time (default: auto)
h = solver._make_handle()
_cabi.set_profiling(h, True)
ds, c0, bmax = solver._tables(solver._base['rho'], solver._base['g'], solver._base['velocityMax'])
up = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(device=dev, dtype=dt)
sizes = [int(a) for a in sys.argv[1:]] or [32, 256, 512, 1024, 2048, 4096, 8192]
reps = 1 if len(sys.argv) > 1 else 3      # one solve per size when a size list is given (ncu captures)
for n in sizes:
    T = np.linspace(1040.0, 1243.0, n)
    zero = np.zeros(n)
    P, M = solver._planes(n, T, zero, zero + 1.0, zero + 1.0, {}, (1 - train.etaTraction) / train.etaTraction, 1 - train.etaRgBrake)
    args = (up(P, torch.float64), up(np.full(n, N, np.int32), torch.int32), up(np.zeros(n, np.int32), torch.int32),
            up(np.array([0, N], np.int32), torch.int32), up(ds, torch.float64), up(c0, torch.float64), up(bmax, torch.float64))
    for rep in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = h.solve_device(*args)
        torch.cuda.synchronize(); w = 1e3 * (time.perf_counter() - t0)
    p = _cabi.last_profile(h)
    print('n=%5d wall %6.2f ms ticks %d ok %d | ' % (n, w, out['ticks'], int((out['status'] == 0).sum().item())) +
          ' '.join('%s %.0f' % (k, 1e3 * v['ms'] / max(1, v['launches'])) for k, v in p.items()), flush=True)
