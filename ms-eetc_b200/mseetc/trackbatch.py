"""Many tracks at once: the preprocessing of reference track.py (mergeDataFrames :377-383, computeDiscretizationPoints :91-107)
for thousands of tracks without one pandas frame per track, feeding the batched device solver directly.

Additive API (no reference analogue; BASELINE configs[4]: 16 384 random tracks with mixed interval counts):
  TrackBatch             step functions of n tracks in CSR form (speed limit [m/s], gradient [permil], curvature [1/m])
  TrackBatch.discretize  grids and forward-filled values of all tracks (native, libmseetc_b200.so: mseetc_discretize_tracks)
  solve_tracks           one device call for all tracks: per-interval SoA tables, optional minimum-time presolve (T = factor * Tmin)
"""
import ctypes
import time as _time

import numpy as np

from mseetc import _cabi


def _csr(rows):
    off = np.zeros(len(rows) + 1, dtype=np.int32)
    off[1:] = np.cumsum([len(r[0]) for r in rows])
    return off, np.concatenate([np.asarray(r[0], dtype=float) for r in rows]), np.concatenate([np.asarray(r[1], dtype=float) for r in rows])


class TrackBatch:
    def __init__(self, length, limits, gradients, curvatures):
        "length [n]; limits / gradients / curvatures: (offsets [n+1], positions, values) with ascending positions starting at 0"
        self.length = np.ascontiguousarray(length, dtype=float)
        self.tables = [tuple(np.ascontiguousarray(a, dtype=(np.int32 if i == 0 else float)) for i, a in enumerate(t))
                       for t in (limits, gradients, curvatures)]
        self.n = len(self.length)

    @classmethod
    def from_tracks(cls, tracks):
        "From Track objects (reference track.py:114-169): their three step-function frames."
        rows = lambda frames: _csr([(f.index.values, f.iloc[:, 0].values) for f in frames])
        return cls([t.length for t in tracks], rows([t.speedLimits for t in tracks]), rows([t.gradients for t in tracks]),
                   rows([t.curvatures for t in tracks]))

    @classmethod
    def random(cls, rng, n, length=None):
        """n synthetic tracks with the statistics of SURVEY.md 8(d) configs 4/5 (same distributions as synthetic.random_track, own
        draw order): length U(5, 50) km, gradient sections every 500-3000 m with N(0, 6 permil) clipped to +-25, speed limits from
        {80, 100, 120, 140} km/h every 5-20 km, curve radius infinite with probability 0.7 else +-U(300, 3000) m."""
        L = rng.uniform(5e3, 50e3, n) if length is None else np.full(n, float(length))

        def sections(lo, hi):
            m = int(np.ceil(L.max() / lo)) + 1
            starts = np.concatenate([np.zeros((n, 1)), np.cumsum(rng.uniform(lo, hi, (n, m)), axis=1)], axis=1)[:, :m]
            return starts, starts < L[:, None]
        gs, gk = sections(500.0, 3000.0)
        gv = np.clip(rng.normal(0.0, 6.0, gs.shape), -25.0, 25.0)
        radius = rng.uniform(300.0, 3000.0, gs.shape) * np.where(rng.uniform(size=gs.shape) < 0.5, 1.0, -1.0)
        cv = np.where(rng.uniform(size=gs.shape) < 0.7, 0.0, 1.0 / radius)
        ls, lk = sections(5e3, 20e3)
        lv = rng.choice([80.0, 100.0, 120.0, 140.0], ls.shape) / 3.6
        pack = lambda starts, keep, vals: (np.concatenate([[0], np.cumsum(keep.sum(axis=1))]).astype(np.int32), starts[keep], vals[keep])
        return cls(L, pack(ls, lk, lv), pack(gs, gk, gv), pack(gs, gk, cv))

    def subset(self, idx):
        idx = np.asarray(idx)

        def take(tab):
            off, pos, val = tab
            sel = np.concatenate([np.arange(off[i], off[i + 1]) for i in idx]) if len(idx) else np.zeros(0, dtype=int)
            return np.concatenate([[0], np.cumsum(off[idx + 1] - off[idx])]).astype(np.int32), pos[sel], val[sel]
        return TrackBatch(self.length[idx], *[take(t) for t in self.tables])

    def discretize(self, numIntervals):
        """Grid and forward-filled values of every track: dict(off [n+1] interval offsets, pos / limit / grad / curv node arrays
        indexed off[t] + t + k, error [n]: 1 where the grid does not come out with numIntervals + 1 points, track.py:103-105)."""
        n = self.n
        nint = np.ascontiguousarray(np.broadcast_to(np.asarray(numIntervals, dtype=np.int32), (n,)))
        off = np.concatenate([[0], np.cumsum(nint)]).astype(np.int32)
        total = int(off[-1]) + n
        out = [np.full(total, np.nan) for _ in range(4)]
        err = np.zeros(n, dtype=np.int32)
        P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        L = _cabi.lib()
        L.mseetc_discretize_tracks.argtypes = [ctypes.c_int32] + [ctypes.c_void_p] * 17
        (lo, lp, lv), (go, gp, gv), (co, cp, cv) = self.tables
        _cabi._check(L.mseetc_discretize_tracks(n, P(self.length), P(nint), P(lo), P(lp), P(lv), P(go), P(gp), P(gv), P(co), P(cp), P(cv),
                                                P(off), P(out[0]), P(out[1]), P(out[2]), P(out[3]), P(err)), 'mseetc_discretize_tracks')
        return dict(n_int=nint, off=off, pos=out[0], limit=out[1], grad=out[2], curv=out[3], error=err)


def _curve_res(kappa, g):
    k = np.abs(kappa)
    return np.where(k <= 1 / 300, g * 0.5 * k / (1 - 30 * k), g * 0.65 * k / (1 - 55 * k))


def solve_tracks(train, batch, numIntervals, optsDict=None, terminalTime=None, timeFactor=None, initialTime=0.0, terminalVelocity=1.0,
                 initialVelocity=1.0, device=None, restart=True):
    """Solve one instance per track of `batch` (TrackBatch) in one device call.

    numIntervals: int or array [n]; optsDict: the options of casadiSolver except numIntervals (reference ocp.py:12-74);
    terminalTime: array [n], or None with timeFactor: T_i = timeFactor * Tmin_i from a time-optimal batch solved first
    (ocp.py:146-150).  Tracks whose grid cannot be built (track.py:103-105) get status -1.  Returns the dictionary of
    mseetc.ocp.solve_instances plus 'tmin', 'grid_error' and the construction / solve times."""
    import torch
    from mseetc.ocp import casadiSolver, classify_losses
    from mseetc.track import Track
    if not torch.cuda.is_available():
        raise RuntimeError("mseetc_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    t_begin = _time.perf_counter()
    n = batch.n
    grid = batch.discretize(numIntervals)
    good = np.flatnonzero(grid['error'] == 0)
    if len(good) < n:                                   # solve the tracks whose grid exists, report the others
        sub = solve_tracks(train, batch.subset(good), np.broadcast_to(np.asarray(numIntervals), (n,))[good], optsDict,
                           None if terminalTime is None else np.broadcast_to(np.asarray(terminalTime, dtype=float), (n,))[good], timeFactor,
                           initialTime, terminalVelocity, initialVelocity, device, restart)
        res = {}
        for k, v in sub.items():
            if isinstance(v, np.ndarray) and v.shape[:1] == (len(good),):
                full = np.zeros((n,) + v.shape[1:], dtype=v.dtype)
                full[good] = v
                res[k] = full
            else:
                res[k] = v
        res['status'][grid['error'] != 0] = -1
        res['grid_error'] = grid['error']
        return res
    nint, off = grid['n_int'], grid['off']
    Nmax = int(nint.max())
    # a representative solver object supplies the problem structure, the option checks and the scalar planes
    rep_track = Track.fromData(batch.length[0], [(0.0, 100.0)])        # structure only: its grid and limits are not used below
    opts = dict(optsDict or {})
    opts['numIntervals'] = int(nint[0])
    rep = casadiSolver(train, rep_track, opts)
    dev = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
    # ---- per-interval SoA tables of all tracks (reference train.py:252-254, ocp.py:266-272), vectorised over the flat node arrays
    b = rep._base
    node_track = np.repeat(np.arange(n), nint + 1)
    first = (off[:-1] + np.arange(n)).astype(np.int64)
    last = first + nint
    is_last = np.zeros(len(node_track), dtype=bool)
    is_last[last] = True
    pos, lim = grid['pos'], grid['limit']
    ds = (np.roll(pos, -1) - pos)[~is_last]
    c0 = (b['g'] * (grid['grad'] / 1e3) / b['rho'] + _curve_res(grid['curv'], b['g']) / b['rho'])[~is_last]      # same rounding as casadiSolver._tables_build
    bmax = np.minimum(np.minimum(lim, b['velocityMax']), np.roll(lim, 1)) ** 2
    bmax[first] = 1.0
    bmax[last] = 1.0
    running = np.add.reduceat(ds / np.minimum(lim, b['velocityMax'])[~is_last], off[:-1].astype(np.int64))        # free-running time at the limits
    bc = lambda a: np.broadcast_to(np.asarray(a, dtype=float), (n,))
    t0, vN, v0 = bc(initialTime), bc(terminalVelocity), bc(initialVelocity)
    lossT, lossR = (classify_losses(train)[1:] if (rep.energyOptimal and rep._lossKind == 'static') else (0.0, 0.0))

    def planes(solver, T):
        P, M = solver._planes(n, T, t0, v0, vN, {}, lossT if solver.energyOptimal else 0.0, lossR if solver.energyOptimal else 0.0)
        vmin = float(solver.velocityMin)
        P[_cabi.PARAM_INDEX['B_START']] = np.minimum(np.maximum(v0, vmin), lim[first]) ** 2          # ocp.py:343-344 with each track's limits
        P[_cabi.PARAM_INDEX['B_END']] = np.minimum(np.maximum(vN, vmin), lim[last]) ** 2
        if not solver.energyOptimal:
            P[_cabi.PARAM_INDEX['OBJ_SCALE']] = batch.length / b['velocityMax']                      # ocp.py:280-282
        return np.ascontiguousarray(P), M
    up = lambda a, dt: torch.from_numpy(np.array(a)).to(device=dev, dtype=dt)
    tabs = (up(nint, torch.int32), up(np.arange(n, dtype=np.int32), torch.int32), up(off, torch.int32), up(ds, torch.float64),
            up(c0, torch.float64), up(bmax, torch.float64))
    t_build = _time.perf_counter() - t_begin

    def handle(solver, guess):
        numSteps, numApprox, tableau = solver._integrator()
        h = _cabi.Handle(Nmax, solver.withPnBrake, solver.withPower, solver.energyOptimal, {'none': 0, 'static': 1, 'dynamic': 2}[solver._lossKind],
                         numSteps, numApprox, int(solver.opts.maxIterations), mu_init=float(solver.muInit),
                         initial_guess={'reference': 0, 'profile': 1}[guess], stall_iterations=int(solver.stallIterations))
        if solver._lossKind == 'dynamic' and solver.energyOptimal:
            dp = solver.train.powerLosses.device_params
            h.set_loss_map(dp['knots_load'], dp['knots_speed'], dp['coef'])
        h.set_sweep_lanes(0 if solver.sweepLanes == 'auto' else int(solver.sweepLanes))
        if tableau is not None:
            h.set_integrator(tableau['A'], tableau['w'], tableau['maxIter'])
        if solver.opts.integrateLosses and solver.energyOptimal:
            h.set_integrate_losses(True)
        return h

    def run(solver, T, tmin_dev=None):
        P, M = planes(solver, T)
        out = handle(solver, solver.initialGuess).solve_device(up(P, torch.float64), *tabs, tmin=tmin_dev)
        r = {k: (v.cpu().numpy() if hasattr(v, 'cpu') else v) for k, v in out.items() if v is not None}
        if restart and solver.initialGuess == 'profile':
            broke = np.flatnonzero((r['status'] == 2) | (r['status'] == 3) | (r['status'] == 5))
            if len(broke):           # once more from the reference's starting point (see casadiSolver.solve_batch)
                sel = torch.from_numpy(broke).to(dev)
                starts = np.concatenate([np.arange(off[i], off[i + 1]) for i in broke])
                nstarts = np.concatenate([np.arange(first[i], last[i] + 1) for i in broke])
                nb = nint[broke]
                sub_tabs = (up(nb, torch.int32), up(np.arange(len(broke), dtype=np.int32), torch.int32),
                            up(np.concatenate([[0], np.cumsum(nb)]), torch.int32), up(ds[starts], torch.float64), up(c0[starts], torch.float64),
                            up(bmax[nstarts], torch.float64))
                h2 = handle(solver, 'reference')
                o2 = h2.solve_device(up(P[:, broke], torch.float64), *sub_tabs)
                r2 = {k: (v.cpu().numpy() if hasattr(v, 'cpu') else v) for k, v in o2.items() if v is not None}
                better = (r2['status'] == 0) | (r2['status'] == 6)
                for key in ('z', 'obj', 'kkt', 'iters', 'status'):
                    r[key][broke[better]] = r2[key][better]
                r['restarted'] = broke
        return r, P, M
    tmin = None
    stp = 4 + int(rep.withPnBrake)
    t_pre = 0.0
    if terminalTime is None:
        if timeFactor is None:
            raise ValueError("solve_tracks needs terminalTime or timeFactor")
        ts = _time.perf_counter()
        tr, _, _ = run(rep._time_sibling(), t0 + 1.5 * running)
        tmin = np.where(tr['status'] == 0, tr['z'][np.arange(n), nint * stp] - t0, 0.0)
        T = t0 + timeFactor * np.where(tmin > 0, tmin, 1.5 * running)
        t_pre = _time.perf_counter() - ts
    else:
        T = bc(terminalTime)
    ts = _time.perf_counter()
    res, P, M = run(rep, T)
    t_solve = _time.perf_counter() - ts
    ok = (res['status'] == 0) | (res['status'] == 6)
    res['z'] = res['z'] * ok[:, None]
    scale = P[_cabi.PARAM_INDEX['OBJ_SCALE']]
    res['cost'] = ((1e-6 / 3.6) * M if rep.energyOptimal else 1.0) * res['obj'] * scale
    res.update(tmin=tmin, terminalTime=T, n_intervals=nint, grid_error=grid['error'],
               timing=dict(construction=t_build, presolve=t_pre, solve=t_solve, wall=_time.perf_counter() - t_begin))
    return res
