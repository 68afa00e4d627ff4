"""Drop-in replacement of ``mseetc.ocp``: same ``OptionsCasadiSolver`` / ``casadiSolver`` construction and
``solve()`` call surface (reference ocp.py:12-74, :77-307, :310-409), but the NLP is never built symbolically
and never handed to IPOPT: construction turns train + track + options into per-interval SoA tables and scalar
parameter planes, and ``solve`` / ``solve_batch`` run the batched interior-point / Riccati kernels of
``libmseetc_b200.so`` on the GPU through the C ABI.

Additive API (no reference analogue): ``casadiSolver.solve_batch`` solves many instances that share this
solver's track and options (trip-time sweeps, parameter Monte Carlo) in one device call.
"""
import time as _time

import numpy as np
import pandas as pd

from mseetc.train import *  # noqa: F401,F403  (the reference re-exports the train module the same way)
from mseetc.train import OptionsRK, OptionsIRK, OptionsCVODES
from mseetc.track import computeDiscretizationPoints
from mseetc.utils import Options, postProcessDataFrame
from mseetc import _cabi

_ACC_INF = 10   # stand-in for an absent force / acceleration limit (reference ocp.py:104)

# Starting point of the interior-point iteration: 'profile' = dynamically consistent speed-envelope profile built on the
# device (same optimum, about half the iterations); 'reference' = the constant guess of the reference (ocp.py:325-339),
# which reproduces IPOPT-like iteration counts.  Per solver: set `solver.initialGuess` before the first solve.
DEFAULT_INITIAL_GUESS = 'profile'
# Trial evaluations (iterations + line-search back-tracks) without a 10 % gain of the best KKT error after which an instance is abandoned with the status it would reach
# anyway (Maximum_Iterations_Exceeded); otherwise one instance cycling around a kink of a non-smooth loss map holds the whole
# lock-step batch until maxIterations.  0 disables the watchdog (`solver.stallIterations = 0`).
DEFAULT_STALL_ITERATIONS = 80
# Sub-batches solved concurrently on separate CUDA streams (one host thread each); see _cabi.StreamPool.  With the sequential
# sweeps (a 150 us dependent chain per iteration whatever the batch size) a second stream filled the gaps (+10 %); with the
# parallel-in-time sweeps one stream is as fast (23.7 vs 23.6 ms per 4096-instance sweep) and every kernel runs alone.
DEFAULT_STREAMS = 1
MIN_INSTANCES_PER_STREAM = 512


class OptionsCasadiSolver(Options):

    def __init__(self, paramsDict):
        self.numIntervals = 100          # shooting intervals (piecewise constant controls)
        self.maxIterations = 1e3         # interior-point iteration limit
        self.energyOptimal = True        # False: minimum time
        self.minimumVelocity = 1         # lower bound on speed [m/s]
        self.integrationMethod = 'RK'    # 'RK', 'IRK' or 'CVODES'
        self.integrationOptions = {}     # options of the chosen method
        self.integrateLosses = False     # False: mid-point rule for the losses of an interval
        super().__init__(paramsDict)

    def overwriteDefaults(self, paramsDict):
        super().overwriteDefaults(paramsDict)
        nested = paramsDict.get('integrationOptions', {})
        kinds = {'RK': OptionsRK, 'IRK': OptionsIRK, 'CVODES': OptionsCVODES}
        if self.integrationMethod in kinds:
            self.integrationOptions = kinds[self.integrationMethod](nested)

    def checkValues(self):
        self.checkPositiveInteger(self.numIntervals, 'Number of intervals', allowZero=False)
        self.checkPositiveInteger(self.maxIterations, 'Maximum number of iterations', allowZero=False)
        if not isinstance(self.energyOptimal, bool):
            raise ValueError("'energyOptimal' flag must be a boolean!")
        if type(self.minimumVelocity) not in {int, float} or self.minimumVelocity <= 0:
            raise ValueError("Minimum velocity should be a strictly positive number!")
        if self.integrationMethod not in {'RK', 'IRK', 'CVODES'}:
            raise ValueError("Unknown integration method!")
        if not isinstance(self.integrateLosses, bool):
            raise ValueError("'integrateLosses' flag must be a boolean!")


def classify_losses(train):
    """Recognise the loss model of a train: ('none'|'static', cT, cR) with
    PLtr/v = cT*f and PLrgb/v = -cR*f (reference train.py:204 + utils.py:197-220).

    Arbitrary Python callables cannot run on the device; the shipped families are recognised by probing."""
    if not hasattr(train, 'powerLosses'):
        if hasattr(train, 'etaTraction') and hasattr(train, 'etaRgBrake'):
            return 'static', (1 - train.etaTraction) / train.etaTraction, 1 - train.etaRgBrake
        raise ValueError("Power losses function of train must by either explicitly or implicitly defined!")
    fun = train.powerLosses
    kind = getattr(fun, 'mseetc_kind', None)
    if kind == 'dynamic':
        return 'dynamic', 0.0, 0.0
    if kind is not None:
        raise NotImplementedError("loss model '{}' is not available in this build of the device library".format(kind))
    fmax = train.forceMax if train.forceMax is not None else _ACC_INF * train.mass * train.rho
    F, V = np.meshgrid(np.array([0.05, 0.3, 0.9]) * fmax, np.array([2.0, 15.0, 33.0]))
    one = np.ones_like(F)
    pos = np.asarray(fun(F, V), dtype=float) * one / (F * V)       # = (1-etaT)/etaT for the static family
    neg = np.asarray(fun(-F, V), dtype=float) * one / (F * V)      # = (1-etaR)
    zero = np.asarray(fun(0.0 * F, V), dtype=float) * one
    cT, cR = float(pos.flat[0]), float(neg.flat[0])
    ok = np.allclose(pos, cT, rtol=1e-12, atol=1e-15) and np.allclose(neg, cR, rtol=1e-12, atol=1e-15) and np.all(zero == 0)
    if not ok:
        raise NotImplementedError("train.powerLosses is not one of the loss families available on the device "
                                  "(none / constant efficiencies)")
    return ('none' if cT == 0 and cR == 0 else 'static'), cT, cR


class _Presolve:
    "Time-optimal solve of the distinct problems of a batch on a side stream, in a second host thread."

    def __init__(self, solver, dev, tmin_dev, args, inverse, mask=None):
        import threading
        self.solver, self.dev, self.tmin_dev, self.args, self.inverse = solver, dev, tmin_dev, args, inverse
        self.mask = mask          # None, or boolean [n]: False = instance the caller asserted feasible (never screened, plane value -1)
        self.tmin, self.error = None, None
        self.thread = threading.Thread(target=self._run, daemon=True)

    def start(self):
        self.thread.start()

    def _run(self):
        import torch
        self.t_start = _time.perf_counter()
        try:
            torch.cuda.set_device(self.dev)
            side = torch.cuda.Stream(device=self.dev, priority=-5)      # tiny, latency-critical kernels: ahead of the batch
            solver, dev = self.solver, self.dev
            sib = solver._time_sibling()
            t0, vN, v0, sub, device = self.args
            m = len(t0)
            f64 = torch.float64
            with torch.cuda.stream(side):
                # lean path: everything up to the moment the certificate is in device memory avoids host round trips and
                # allocations (persistent staging / result buffers of the sibling, in-place device arithmetic)
                lim = np.minimum(solver.points['Speed limit [m/s]'].values[:-1], solver._base['velocityMax'])
                horizon = t0 + 1.5 * float(np.sum(solver.steps / lim))          # same bound as minimum_time
                P, _ = sib._planes(m, horizon, t0, v0, vN, dict(sub), 0.0, 0.0)
                ds, c0, bmax, trk_of, trk_off = sib._track_tables(m, sub)
                i32 = torch.int32
                args = (sib._upload('P', P, f64, dev), sib._upload('nint', np.full(m, sib.numIntervals, np.int32), i32, dev),
                        sib._upload('trk_of', trk_of, i32, dev), sib._upload('trk_off', trk_off, i32, dev),
                        sib._upload('ds', ds, f64, dev), sib._upload('c0', c0, f64, dev), sib._upload('bmax', bmax, f64, dev))
                inv = sib._upload('inverse', np.asarray(self.inverse, dtype=np.int64), torch.int64, dev)
                t0_dev = sib._upload('t0', np.ascontiguousarray(t0, dtype=float), f64, dev)
                out = sib._ensure_handle().solve_device(*args, want_z=True, out=sib._device_out('solve', m, dev, False))
                pre = sib._dev.setdefault('pre', {})
                if pre.get('m') != m or pre['dur'].device != dev:
                    pre.update(m=m, dur=torch.empty(m, dtype=f64, device=dev), ok=torch.empty(m, dtype=torch.bool, device=dev))
                torch.sub(out['z'][:, -2], t0_dev, out=pre['dur'])
                torch.eq(out['status'], 0, out=pre['ok'])
                pre['dur'].mul_(pre['ok'])                   # no certificate where the time-optimal solve did not converge
                if self.mask is None:
                    torch.index_select(pre['dur'], 0, inv, out=self.tmin_dev)
                else:
                    keep = sib._upload('mask', np.asarray(self.mask, dtype=np.bool_), torch.bool, dev)
                    tmp = torch.index_select(pre['dur'], 0, inv)
                    tmp.masked_fill_(~keep, -1.0)
                    self.tmin_dev.copy_(tmp)
                side.synchronize()
                t_pub = _time.perf_counter()
                dur = pre['dur'].cpu().numpy()
                if np.any(dur <= 0.0):
                    # rare: retry the instances that did not converge with longer horizons (host path), publish again
                    dur2, st2 = solver.minimum_time(t0, vN, v0, overrides=sub, device=device)
                    dur = np.where(st2 == 0, dur2, 0.0)
                    full = np.ascontiguousarray(dur[self.inverse])
                    if self.mask is not None:
                        full[~np.asarray(self.mask)] = -1.0
                    self.tmin_dev.copy_(torch.from_numpy(full).to(dev))
                    side.synchronize()
            self.tmin = np.ascontiguousarray(dur[self.inverse])
            if self.mask is not None:
                self.tmin[~np.asarray(self.mask)] = 0.0        # not computed for instances the caller asserted feasible
            self.detail = dict(publish_at=t_pub - self.t_start)
        except Exception as exc:      # surfaced by join()
            self.error = exc
        self.t_end = _time.perf_counter()

    def join(self):
        self.thread.join()
        if self.error is not None:
            raise self.error
        return self.tmin


def _curve_res(kappa, g):
    k = np.abs(kappa)
    return np.where(k <= 1 / 300, g * 0.5 * k / (1 - 30 * k), g * 0.65 * k / (1 - 55 * k))


class casadiSolver():
    "Solver object with the reference's name and interface; the numerics run on the GPU."

    def __init__(self, train, track, optsDict={}):
        track.checkFields()
        train.checkFields()
        opts = OptionsCasadiSolver(optsDict)

        self.train = train
        self.opts = opts
        self.numIntervals = int(opts.numIntervals)
        self.velocityMin = opts.minimumVelocity
        self.energyOptimal = opts.energyOptimal
        self.totalMass = train.mass * train.rho
        self.withRgBrake = train.forceMin != 0
        self.withPnBrake = train.forceMinPn != 0
        self.withPower = train.powerMax is not None or train.powerMin is not None
        self.trackLength = track.length
        if opts.energyOptimal:
            self.scalingFactorObjective = 3.6 / (1e-6 * self.totalMass)           # objective in kWh
        else:
            self.scalingFactorObjective = track.length / train.velocityMax        # fastest conceivable trip

        self.points = computeDiscretizationPoints(track, self.numIntervals)
        self.steps = np.diff(self.points.index)
        self._lossKind, _, _ = classify_losses(train) if opts.energyOptimal else ('none', 0.0, 0.0)
        # snapshot of everything the kernels need; train attributes are read NOW (callers mutate them afterwards)
        self._base = self._train_scalars(train)
        self._handle = None
        self._dev = {}
        self.initialGuess = DEFAULT_INITIAL_GUESS
        self.stallIterations = DEFAULT_STALL_ITERATIONS
        self.muInit = 0.1             # IPOPT's mu_init (the reference leaves the default)
        self.streams = DEFAULT_STREAMS
        self._pool = None
        self.sweepLanes = 'auto'      # 1 sequential sweeps | 8 | 16 | 32 chunk lanes per instance (parallel in time) | 'auto'

    # ------------------------------------------------------------------ packing
    @staticmethod
    def _train_scalars(train):
        keys = ('mass', 'rho', 'g', 'velocityMax', 'forceMax', 'forceMin', 'forceMinPn', 'powerMax', 'powerMin', 'accMax',
                'accMin', 'r0', 'r1', 'r2')
        d = {k: getattr(train, k) for k in keys}
        return d

    def _planes(self, n, T, t0, v0, vN, overrides, lossT, lossR):
        "Parameter planes [PARAM_COUNT, n] in specific units (reference ocp.py:96-116,343-355)."
        b = dict(self._base)
        get = lambda k: np.broadcast_to(np.asarray(overrides[k], dtype=float), (n,)) if k in overrides else b[k]
        mass, rho = get('mass'), get('rho')
        M = mass * rho
        vmax = get('velocityMax')
        opt = lambda k, dflt: (get(k) / M) if (k in overrides or b[k] is not None) else dflt
        fMax = opt('forceMax', _ACC_INF)
        fMin = opt('forceMin', -_ACC_INF) if self.withRgBrake else 0.0
        fMinPn = opt('forceMinPn', -_ACC_INF) if self.withPnBrake else -1.0
        hasPmax = 'powerMax' in overrides or b['powerMax'] is not None
        hasPmin = 'powerMin' in overrides or b['powerMin'] is not None
        if self.withPower:
            up = get('powerMax') / M if hasPmax else fMax * vmax
            lo = 0.0 if not self.withRgBrake else (get('powerMin') / M if hasPmin else fMin * vmax)
            pUp, pLo = np.abs(up), -np.abs(lo)
        else:
            pUp, pLo = 1.0, -1.0
        accMax = np.minimum(_ACC_INF, get('accMax')) if ('accMax' in overrides or b['accMax'] is not None) else _ACC_INF
        accMin = np.maximum(-_ACC_INF, -np.abs(get('accMin'))) if ('accMin' in overrides or b['accMin'] is not None) else -_ACC_INF
        scale = 3.6 / (1e-6 * M) if self.energyOptimal else self.trackLength / vmax
        lim0 = self.points['Speed limit [m/s]'].values[0]
        limN = self.points['Speed limit [m/s]'].values[-1]
        v0c = np.minimum(np.maximum(v0, self.velocityMin), lim0)
        vNc = np.minimum(np.maximum(vN, self.velocityMin), limN)
        P = np.empty((len(_cabi.PARAMS), n))
        rows = dict(SR0=get('r0') / M, SR1=get('r1') / M, SR2=get('r2') / M, FEL_LO=fMin, FEL_UP=fMax, FPB_LO=fMinPn,
                    POW_LO=pLo, POW_UP=pUp, ACC_LO=accMin, ACC_UP=accMax, LOSS_TR=lossT, LOSS_RG=lossR,
                    BMIN=float(self.velocityMin) ** 2, OBJ_SCALE=scale, T_END=T, T_START=t0, B_START=v0c ** 2, B_END=vNc ** 2, MASS=M)
        dyn = dict(DYN_AUX=0.0, DYN_ETAG=1.0, DYN_FMAX=1.0, DYN_PMAX=1.0, DYN_SCALE=1.0)
        if self._lossKind == 'dynamic':
            dp = self.train.powerLosses.device_params
            pick = lambda key, dflt: np.broadcast_to(np.asarray(overrides[key], dtype=float), (n,)) if key in overrides else dflt
            dyn = dict(DYN_AUX=pick('auxiliaries', dp['auxiliaries']), DYN_ETAG=pick('etaGear', dp['etaGear']), DYN_FMAX=dp['forceMax'],
                       DYN_PMAX=dp['powerMax'], DYN_SCALE=pick('tableScale', dp['tableScale']))
        rows.update(dyn)
        for name, val in rows.items():
            P[_cabi.PARAM_INDEX[name]] = val
        return P, M

    def _tables(self, rho, g, vmax):
        "Per-interval tables ds, c0 and node table bmax for one (rho, g, vmax); the last result is kept (the grid is fixed at construction)."
        key = (float(rho), float(g), float(vmax))
        cached = self._dev.get('tables')
        if cached is not None and cached[0] == key:
            return cached[1]
        out = self._tables_build(rho, g, vmax)
        self._dev['tables'] = (key, out)
        return out

    def _tables_build(self, rho, g, vmax):
        pts = self.points
        N = self.numIntervals
        grad = pts['Gradient [permil]'].values[:N] / 1e3
        curv = pts['Curvature [1/m]'].values[:N]
        lim = pts['Speed limit [m/s]'].values
        c0 = g * grad / rho + _curve_res(curv, g) / rho                          # reference train.py:252-254
        bmax = np.ones(N + 1)
        bmax[1:N] = np.minimum(np.minimum(lim[1:N], vmax), lim[0:N - 1]) ** 2    # reference ocp.py:266-269
        return np.asarray(self.steps, dtype=float), c0, bmax

    def _integrator(self):
        "(numSteps, numApproxSteps, collocation tableau or None) of this solver's integrationMethod / integrationOptions"
        from mseetc.train import integratorSetup
        return integratorSetup(self.opts.integrationMethod, self.opts.integrationOptions)

    def _integrator_key(self):
        ns, na, tab = self._integrator()
        return (self.opts.integrationMethod, ns, na) + (() if tab is None else (tab['A'].tobytes(), tab['maxIter']))

    def _make_handle(self):
        numSteps, numApprox, tableau = self._integrator()
        h = _cabi.Handle(self.numIntervals, self.withPnBrake, self.withPower, self.energyOptimal,
                         {'none': 0, 'static': 1, 'dynamic': 2}[self._lossKind], numSteps, numApprox,
                         int(self.opts.maxIterations), mu_init=float(self.muInit),
                         initial_guess={'reference': 0, 'profile': 1}[self.initialGuess], stall_iterations=int(self.stallIterations))
        if self._lossKind == 'dynamic' and self.energyOptimal:
            dp = self.train.powerLosses.device_params
            h.set_loss_map(dp['knots_load'], dp['knots_speed'], dp['coef'])
        h.set_sweep_lanes(0 if self.sweepLanes == 'auto' else int(self.sweepLanes))      # 0: the library picks per call
        if tableau is not None:            # 'IRK' / 'CVODES' (reference train.py:303-322): collocation steps instead of explicit RK4
            h.set_integrator(tableau['A'], tableau['w'], tableau['maxIter'])
        if self.opts.integrateLosses and self.energyOptimal:      # reference ocp.py:118-120,231-241
            h.set_integrate_losses(True)
        return h

    def _ensure_pool(self, dev):
        if self._pool is None or len(self._pool.handles) != int(self.streams) or self._pool.device != dev:
            self._pool = _cabi.StreamPool(self._make_handle, int(self.streams), dev)
        return self._pool

    def _pinned(self, key, like):
        """Page-locked staging buffers are kept per solver in a ring of three per output (allocating 50 MB of pinned memory per call
        costs more than the solve): a returned array stays valid until the second-next solve_batch call on this solver (a call
        may use two slots: its own and the one of the nested re-solve of wrongly screened instances)."""
        import torch
        ring = self._dev.setdefault('pinned', {})
        slot = ring.setdefault(key, {'bufs': [None, None, None], 'next': 0})
        i = slot['next']
        slot['next'] = (i + 1) % 3
        buf = slot['bufs'][i]
        if buf is None or buf.shape != like.shape or buf.dtype != like.dtype:
            buf = torch.empty(like.shape, dtype=like.dtype, pin_memory=True)
            slot['bufs'][i] = buf
        return buf

    def _device_out(self, tag, n, dev, want_lam):
        "Device result buffers are kept per solver and batch size as well (no cudaMalloc of 50 MB blocks inside a solve)."
        import torch
        h = self._ensure_handle()
        N, stp = h.problem.n_intervals_max, 3 + h.nu
        key = (tag, n, str(dev), bool(want_lam))
        cache = self._dev.setdefault('out', {})
        if key not in cache:
            for k in [k for k in cache if k[0] == tag]:
                del cache[k]
            f64, i32 = dict(dtype=torch.float64, device=dev), dict(dtype=torch.int32, device=dev)
            cache[key] = dict(z=torch.zeros((n, N * stp + 2), **f64), lam=torch.zeros((n, N * h.rows), **f64) if want_lam else None,
                              obj=torch.empty(n, **f64), kkt=torch.empty(n, **f64), iters=torch.empty(n, **i32), status=torch.empty(n, **i32))
        return dict(cache[key])

    def _upload(self, name, array, dtype, dev):
        """Host array -> persistent page-locked staging buffer -> persistent device tensor, asynchronously on the current stream
        (a pageable cudaMemcpy can stall behind unrelated work of the context, and allocates on every call)."""
        import torch
        a = np.ascontiguousarray(array)
        cache = self._dev.setdefault('in', {})
        slot = cache.get(name)
        if slot is None or slot[0].shape != a.shape or slot[1].device != dev or slot[0].dtype != dtype:
            slot = (torch.empty(a.shape, dtype=dtype, pin_memory=True), torch.empty(a.shape, dtype=dtype, device=dev))
            cache[name] = slot
        slot[0].copy_(torch.from_numpy(a))
        slot[1].copy_(slot[0], non_blocking=True)
        return slot[1]

    def _track_tables(self, n, overrides, perm=None):
        "Track tables of a batch: shared unless rho / velocityMax vary per instance (then one table per instance, in the caller's order)."
        N = self.numIntervals
        if any(k in overrides for k in ('rho', 'velocityMax')):
            rho = np.broadcast_to(np.asarray(overrides.get('rho', self._base['rho']), dtype=float), (n,))
            vmx = np.broadcast_to(np.asarray(overrides.get('velocityMax', self._base['velocityMax']), dtype=float), (n,))
            tabs = [self._tables(rho[i], self._base['g'], vmx[i]) for i in range(n)]
            ds = np.concatenate([t[0] for t in tabs]); c0 = np.concatenate([t[1] for t in tabs]); bmax = np.concatenate([t[2] for t in tabs])
            trk_of = np.arange(n, dtype=np.int32) if perm is None else perm.astype(np.int32)
            trk_off = (np.arange(n + 1) * N).astype(np.int32)
        else:
            ds, c0, bmax = self._tables(self._base['rho'], self._base['g'], self._base['velocityMax'])
            trk_of = np.zeros(n, dtype=np.int32)
            trk_off = np.array([0, N], dtype=np.int32)
        return ds, c0, bmax, trk_of, trk_off

    def _ensure_handle(self):
        if self._handle is None:
            self._handle = self._make_handle()
        return self._handle

    # ------------------------------------------------------------------ batched solve (additive API)
    def _time_sibling(self):
        "The same problem in minimum-time mode (used to certify infeasible trip times)."
        if getattr(self, '_sibling', None) is None:
            import copy
            sib = copy.copy(self)
            sib.energyOptimal = False
            sib._lossKind = 'none'
            sib._handle = None
            sib._sibling = None
            sib._restart = None
            sib._pool = None
            sib._dev = {}             # own staging / result buffers: the sibling runs concurrently on a second host thread
            sib.scalingFactorObjective = self.trackLength / self._base['velocityMax']
            if self.initialGuess == 'profile':
                # the speed-envelope starting profile is close to the time-optimal run: a small initial barrier parameter keeps
                # the iteration near it (27 instead of 41 iterations on CH_StGallen_Wil, profiles/probe_muinit_time.py)
                sib.muInit = 1e-4
            self._sibling = sib
        return self._sibling

    def _restart_sibling(self):
        "The same problem started from the reference's own initial guess (second attempt for instances whose iteration broke down)."
        if getattr(self, '_restart', None) is None:
            import copy
            sib = copy.copy(self)
            sib.initialGuess = 'reference'
            sib.muInit = 0.1
            sib._handle, sib._sibling, sib._restart, sib._pool, sib._dev = None, None, None, None, {}
            self._restart = sib
        return self._restart

    def minimum_time(self, initialTime=0, terminalVelocity=1, initialVelocity=1, overrides=None, device=None):
        """Minimum trip duration t_N - t_0 of each instance (time-optimal mode of the same problem,
        reference ocp.py:146-150).  Returns (durations, status)."""
        sib = self._time_sibling()
        t0 = np.atleast_1d(np.asarray(initialTime, dtype=float))
        # upper bound on t_N for the time-optimal solve: a multiple of the free-running time at the speed limits
        # (the initial guess spreads the nodes linearly up to that bound, ocp.py:325-339, so it should not be loose)
        lim = np.minimum(self.points['Speed limit [m/s]'].values[:-1], self._base['velocityMax'])
        running = float(np.sum(self.steps / lim))
        res = None
        for factor in (1.5, 2.5, 4.0):
            cur = sib.solve_batch(t0 + factor * running, initialTime, terminalVelocity, initialVelocity, overrides=overrides,
                                  screen=False, device=device)
            if res is None:
                # own copies: `cur` are views of a two-deep ring of pinned buffers that the next passes reuse
                res = {k: (np.array(v) if isinstance(v, np.ndarray) else v) for k, v in cur.items()}
            else:
                redo = res['status'] != 0
                for key in ('z', 'status', 'obj', 'kkt', 'iters'):
                    res[key][redo] = cur[key][redo]
            if np.all(res['status'] == 0):
                break
        return res['z'][:, -2] - np.broadcast_to(t0, res['z'][:, -2].shape), res['status']

    def solve_batch(self, terminalTime, initialTime=0, terminalVelocity=1, initialVelocity=1, overrides=None,
                    want_multipliers=False, device=None, screen=True, to_host=True, restart=True, tables=False):
        """Solve n instances that share this solver's track, options and problem structure.

        terminalTime / initialTime / terminalVelocity / initialVelocity: scalars or arrays of length n.
        overrides: optional dict of per-instance train attributes (arrays of length n), any of
        mass, rho, r0, r1, r2, forceMax, forceMin, forceMinPn, powerMax, powerMin, accMax, accMin, velocityMax,
        etaTraction, etaRgBrake; with the dynamic loss map also auxiliaries, etaGear, tableScale.
        Returns a dict of numpy arrays: z [n, nz] (reference variable order), cost, kkt, iters, status,
        plus timing; nothing is post-processed.  The arrays are views of page-locked staging buffers that are reused by the
        second-next call on this solver -- copy them if they must live longer.

        screen=True (energy-optimal mode only): terminalTime is an upper bound on t_N, so an instance is infeasible
        exactly when it is below the minimum trip time.  The minimum time of every distinct
        (train, boundary speeds) combination in the batch is computed first by a time-optimal solve and instances
        below it are reported as 'Infeasible_Problem_Detected' without iterating.  `screen` may also be a boolean array of
        length n: False marks instances the caller knows to be feasible (e.g. a Monte Carlo at a timetable value with
        slack) -- they take no part in the minimum-time presolve and are never screened.

        restart=True: an instance whose iteration breaks down from the speed-envelope starting profile (line search or step
        computation fails -- where IPOPT would enter its restoration phase, which this solver does not have) is solved once
        more from the reference's own starting point (ocp.py:325-339, mu_init 0.1), from which the iteration takes the
        reference's path.

        tables=True adds the post-processed trajectory tables of all instances, computed on the device (what
        postProcessDataFrame adds to the table that solve() returns, reference ocp.py:407 / utils.py:223-336, with the defaults
        CVODES=True, integrateLosses=False): res['table'] is an array [n, numIntervals+1, 23] with the columns
        res['table_columns'] (index column 'Time [s]' first); casadiSolver.table(res, i) makes the DataFrame of instance i.
        Rows of failed instances are NaN.

        to_host=False leaves z / lam / obj / kkt / iters / status as torch tensors on the device (used by
        mseetc.sharding.solve_batch_sharded, which gathers them over NVLink before one device-to-host copy)."""
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("mseetc_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        overrides = dict(overrides or {})
        overrides_in = dict(overrides)
        arrs = [np.atleast_1d(np.asarray(a, dtype=float)) for a in (terminalTime, initialTime, terminalVelocity, initialVelocity)]
        n = max([len(a) for a in arrs] + [len(np.atleast_1d(v)) for v in overrides.values()])
        screen_mask = None
        if not isinstance(screen, (bool, np.bool_)):
            screen_mask = np.broadcast_to(np.asarray(screen, dtype=bool), (n,))
            screen = bool(screen_mask.any())
            if screen_mask.all():
                screen_mask = None
        T, t0, vN, v0 = [np.broadcast_to(a, (n,)) for a in arrs]
        etaT = np.asarray(overrides.pop('etaTraction', getattr(self.train, 'etaTraction', 1.0)), dtype=float)
        etaR = np.asarray(overrides.pop('etaRgBrake', getattr(self.train, 'etaRgBrake', 1.0)), dtype=float)
        if not self.energyOptimal or self._lossKind == 'dynamic':
            lossT, lossR = 0.0, 0.0
        elif hasattr(self.train, 'powerLosses'):
            _, lossT, lossR = classify_losses(self.train)
        else:
            lossT, lossR = (1 - etaT) / etaT, 1 - etaR
        t_begin = _time.perf_counter()
        P, M = self._planes(n, T, t0, v0, vN, overrides, lossT, lossR)
        # several streams: the instances are dealt to the sub-batches tile by tile (balanced work, see StreamPool.interleave);
        # everything below works in that order and the results are put back in the caller's order on the device
        pooled = int(self.streams) > 1 and n >= 2 * MIN_INSTANCES_PER_STREAM
        perm, parts = _cabi.StreamPool.interleave(n, int(self.streams)) if pooled else (None, None)
        if perm is not None:
            P = np.ascontiguousarray(P[:, perm])
        dev = torch.device(device if device is not None else 'cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        tmin = None
        presolve = None
        if screen and self.energyOptimal:
            # minimum trip time of every distinct problem, computed CONCURRENTLY (second host thread + second stream) with
            # the energy-optimal batch; the kernels pick the values up as soon as they are in device memory
            # a pure trip-time sweep (no overrides, scalar boundary data) is one problem: no need to compare the columns
            same_problem = not overrides_in and all(len(a) == 1 for a in arrs[1:])
            key = None if same_problem else np.delete(P, [_cabi.PARAM_INDEX['T_END'], _cabi.PARAM_INDEX['LOSS_TR'], _cabi.PARAM_INDEX['LOSS_RG'],
                                                         _cabi.PARAM_INDEX['OBJ_SCALE'], _cabi.PARAM_INDEX['DYN_AUX'], _cabi.PARAM_INDEX['DYN_ETAG'],
                                                         _cabi.PARAM_INDEX['DYN_SCALE']], axis=0)
            mask_dealt = None
            if screen_mask is not None:
                mask_dealt = screen_mask[perm] if perm is not None else screen_mask
            if n == 1 or same_problem or np.all(key == key[:, :1]):
                first, inverse = np.array([0]), np.zeros(n, dtype=np.intp)
            elif mask_dealt is None:
                _, first, inverse = np.unique(key, axis=1, return_index=True, return_inverse=True)
                inverse = np.asarray(inverse).reshape(-1)
            else:
                # only the instances that are to be screened take part in the presolve
                sel = np.flatnonzero(mask_dealt)
                _, f_sel, inv_sel = np.unique(key[:, sel], axis=1, return_index=True, return_inverse=True)
                first = sel[f_sel]
                inverse = np.zeros(n, dtype=np.intp)
                inverse[sel] = np.asarray(inv_sel).reshape(-1)
            first = perm[first] if perm is not None else first          # indices into the caller's arrays
            sub = {k: np.broadcast_to(np.asarray(v, dtype=float), (n,))[first] for k, v in overrides.items()}
            tmin_dev = torch.zeros(n, dtype=torch.float64, device=dev)      # 0 = "not known yet", -1 = "never screen this instance"
            if mask_dealt is not None:
                tmin_dev.masked_fill_(torch.from_numpy(np.ascontiguousarray(~mask_dealt)).to(dev), -1.0)
            presolve = _Presolve(self, dev, tmin_dev, (t0[first], vN[first], v0[first], sub, device), inverse, mask_dealt)
        N = self.numIntervals
        ds, c0, bmax, trk_of, trk_off = self._track_tables(n, overrides, perm)
        t_pack = _time.perf_counter() - t_begin
        f64, i32 = torch.float64, torch.int32
        args = (self._upload('P', P, f64, dev), self._upload('nint', np.full(n, N, np.int32), i32, dev), self._upload('trk_of', trk_of, i32, dev),
                self._upload('trk_off', trk_off, i32, dev), self._upload('ds', ds, f64, dev), self._upload('c0', c0, f64, dev),
                self._upload('bmax', bmax, f64, dev))
        tm = presolve.tmin_dev if presolve is not None else None
        t_up = _time.perf_counter() - t_begin
        buf = self._device_out('solve', n, dev, want_multipliers)
        back = self._upload('back', _cabi.StreamPool.interleave_inverse(n, int(self.streams)), torch.int64, dev) if pooled else None
        ordered = self._device_out('ordered', n, dev, want_multipliers) if pooled else None
        # ---- concurrent part: the presolve thread and the stream threads of the pool spend their time inside the library (GIL
        # released), but each needs the GIL for a moment to get there; with the interpreter's default 5 ms switch interval those
        # hand-overs can cost more than the solve, so it is shortened for the duration of the call
        import sys
        interval = sys.getswitchinterval()
        sys.setswitchinterval(1e-4)
        try:
            if presolve is not None and not pooled:
                presolve.start()
            if pooled:
                out = self._ensure_pool(dev).solve(*args, tmin=tm, want_lam=want_multipliers, parts=parts, out=buf,
                                                   on_started=presolve.start if presolve is not None else None)
                for k, v in ordered.items():
                    if v is not None:
                        torch.index_select(out[k], 0, back, out=v)
                out = dict(out, **{k: v for k, v in ordered.items() if v is not None})
            else:
                out = self._ensure_handle().solve_device(*args, want_z=True, want_lam=want_multipliers, tmin=tm, out=buf)
            t_solve = _time.perf_counter() - t_begin
            if presolve is not None:
                tmin = presolve.join()
                if perm is not None:
                    tmin = tmin[_cabi.StreamPool.interleave_inverse(n, int(self.streams))]
        finally:
            sys.setswitchinterval(interval)
        t_join = _time.perf_counter() - t_begin
        # the reference returns no trajectory for a failed solve (ocp.py:364-370): blank those rows on the device
        out['z'].mul_(((out['status'] == 0) | (out['status'] == 6)).to(out['z'].dtype).unsqueeze(1))
        res = {}
        if not to_host:
            # results stay on the device; the certificate is applied there as well
            if tmin is not None:
                tmin_t = torch.from_numpy(np.ascontiguousarray(tmin)).to(dev)
                short_t = (tmin_t > 0) & (torch.from_numpy(np.ascontiguousarray(T - t0)).to(dev) < tmin_t * (1 - _cabi.TMIN_MARGIN))
                out['status'][(out['status'] != 0) & short_t] = 4
            res = {k: v for k, v in out.items() if v is not None}
            torch.cuda.current_stream(dev).synchronize()
            res['tmin'] = tmin
            res['h2d_bytes'] = int(P.nbytes + 4 * n * 2 + trk_off.nbytes + ds.nbytes + c0.nbytes + bmax.nbytes + (tmin.nbytes if tmin is not None else 0))
            res['wall'] = _time.perf_counter() - t_begin
            res['totalMass'] = M
            res['scale'] = P[_cabi.PARAM_INDEX['OBJ_SCALE']]
            return res
        for k, v in out.items():                          # device -> pinned host buffers -> numpy
            if v is None:
                continue
            if hasattr(v, 'cpu'):
                host = self._pinned(k, v)
                host.copy_(v, non_blocking=True)
                res[k] = host
            else:
                res[k] = v
        torch.cuda.current_stream(dev).synchronize()
        res = {k: (v.numpy() if hasattr(v, 'numpy') else v) for k, v in res.items()}
        if tmin is not None:
            # an instance below its minimum trip time is infeasible whatever the iteration did before the certificate arrived
            short = (tmin > 0) & ((T - t0) < tmin * (1 - _cabi.TMIN_MARGIN))
            res['status'][(res['status'] != 0) & (res['status'] != 6) & short] = 4
            # the device also screens with a speed-envelope bound before the first iteration (inst_screen, core.cuh); every such
            # flag is checked against the exact certificate and an instance flagged wrongly is solved again without screening
            wrong = np.flatnonzero((res['status'] == 4) & ~short)
            if len(wrong):
                pick = lambda a: np.broadcast_to(np.asarray(a, dtype=float), (n,))[wrong]
                redo = self.solve_batch(T[wrong], t0[wrong], vN[wrong], v0[wrong], overrides={k: pick(v) for k, v in overrides_in.items()},
                                        want_multipliers=want_multipliers, device=device, screen=False)
                for key in ('z', 'lam', 'obj', 'kkt', 'iters', 'status'):
                    if key in res and isinstance(res[key], np.ndarray):
                        res[key][wrong] = redo[key]

        if restart and self.initialGuess == 'profile':
            broke = np.flatnonzero((res['status'] == 2) | (res['status'] == 3) | (res['status'] == 5))
            if len(broke):
                pick = lambda a: np.broadcast_to(np.asarray(a, dtype=float), (n,))[broke]
                redo = self._restart_sibling().solve_batch(T[broke], t0[broke], vN[broke], v0[broke], overrides={k: pick(v) for k, v in overrides_in.items()},
                                                           want_multipliers=want_multipliers, device=device, screen=False, restart=False)
                better = (redo['status'] == 0) | (redo['status'] == 6)
                for key in ('z', 'lam', 'obj', 'kkt', 'iters', 'status'):
                    if key in res and isinstance(res[key], np.ndarray):
                        res[key][broke[better]] = redo[key][better]
                res['restarted'] = broke
        if tables:
            res['table'] = self._tables_on_device(res, P, ds, c0, trk_of, trk_off, overrides, perm, n, dev)
            res['table_columns'] = _cabi.TABLE_COLUMNS
        res['h2d_bytes'] = int(P.nbytes + 4 * n * 2 + trk_off.nbytes + ds.nbytes + c0.nbytes + bmax.nbytes + (tmin.nbytes if tmin is not None else 0))
        res['d2h_bytes'] = int(sum(v.nbytes for v in res.values() if isinstance(v, np.ndarray)))
        res['tmin'] = tmin
        res['timing'] = dict(pack=t_pack, upload=t_up - t_pack, solve=t_solve - t_up, presolve_join=t_join - t_solve,
                             d2h=_time.perf_counter() - t_begin - t_join)
        self._last_timing = dict(res['timing'])
        if presolve is not None:
            res['timing']['presolve_thread'] = presolve.t_end - presolve.t_start
            for k, v in getattr(presolve, 'detail', {}).items():
                res['timing']['presolve_' + k] = v
        if isinstance(out.get('spans_ms'), list):
            res['spans_ms'] = out['spans_ms']
        res['wall'] = _time.perf_counter() - t_begin
        scale = P[_cabi.PARAM_INDEX['OBJ_SCALE']]
        # reference ocp.py:361: cost in kWh (energy) or s (time)
        res['cost'] = ((1e-6 / 3.6) * M if self.energyOptimal else 1.0) * res['obj'] * scale
        res['totalMass'] = M
        return res

    def _tables_on_device(self, res, P, ds, c0, trk_of, trk_off, overrides, perm, n, dev):
        "Post-processed tables of a solved batch (device kernels of csrc/table.cuh); P, trk_of are in the dealt order `perm`."
        import torch
        f64, i32 = torch.float64, torch.int32
        up = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(device=dev, dtype=dt)
        N = self.numIntervals
        back = np.argsort(perm) if perm is not None else None
        Pc = P[:, back] if back is not None else P                    # caller's order, like the results
        tk = trk_of[back] if back is not None else trk_of
        ntracks = len(trk_off) - 1
        pts = self.points
        one = np.stack([pts.index.values.astype(float), pts['Speed limit [m/s]'].values.astype(float),
                        pts['Gradient [permil]'].values.astype(float), pts['Curvature [1/m]'].values.astype(float)])
        nodes = np.tile(one, (1, ntracks))
        mass = np.array(np.broadcast_to(np.asarray(overrides.get('mass', self._base['mass']), dtype=float), (n,)))
        tab = _cabi.postprocess_device(self._ensure_handle(), up(res['z'], f64), up(Pc, f64), up(np.full(n, N, np.int32), i32), up(tk, i32),
                                       up(trk_off, i32), up(ds, f64), up(c0, f64), up(nodes, f64), up(mass, f64), up(res['status'], i32))
        host = self._pinned('table', tab)
        host.copy_(tab, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        return host.numpy()

    @staticmethod
    def table(res, i):
        "DataFrame of instance i of a solve_batch(..., tables=True) result: the table solve() returns (reference ocp.py:401-407)."
        cols = res['table_columns']
        df = pd.DataFrame(res['table'][i][:, 1:], columns=list(cols[1:]), index=pd.Index(res['table'][i][:, 0], name=cols[0]))
        return df

    # ------------------------------------------------------------------ reference call surface
    def solve(self, terminalTime, initialTime=0, terminalVelocity=1, initialVelocity=1):
        if not isinstance(initialTime, (int, float)) or initialTime < 0:
            raise ValueError("Initial time must be a positive number, not {}!".format(initialTime))
        if not isinstance(terminalTime, (int, float)) or terminalTime <= 0:
            raise ValueError("Terminal time must be a strictly positive number, not {}!".format(terminalTime))
        res = self.solve_batch(terminalTime, initialTime, terminalVelocity, initialVelocity, screen=False)
        status = int(res['status'][0])
        if status not in _cabi.SUCCESS_CODES and self.energyOptimal:
            # classify the failure: below the minimum trip time the problem is infeasible (what IPOPT's restoration
            # phase would report); the time-optimal solve is only paid for on failure
            dur, st = self.minimum_time(initialTime, terminalVelocity, initialVelocity)
            if st[0] == 0 and (terminalTime - initialTime) < dur[0] * (1 - _cabi.TMIN_MARGIN):
                status = 4
        stats = {'Solver status': _cabi.STATUS_STRINGS.get(status, 'Internal_Error'), 'IP iterations': int(res['iters'][0]),
                 'CPU time [s]': res['wall'], 'Cost': float(res['cost'][0])}
        if status not in _cabi.SUCCESS_CODES:
            print("Solver failed with status '{}'".format(stats['Solver status']))
            return None, stats
        print("Solver converged in {:4d} iterations.".format(stats['IP iterations']))
        df = self.table_from_z(res['z'][0])
        df = postProcessDataFrame(df, self.points, self.train)
        return df, stats

    def table_from_z(self, z):
        "De-interleave one solution vector into the reference's raw table (reference ocp.py:376-405)."
        N = self.numIntervals
        stp = 4 + int(self.withPnBrake)
        body = np.asarray(z[:N * stp]).reshape(N, stp)
        tail = z[N * stp:N * stp + 2]
        nanrow = lambda col: np.append(col, np.nan)
        o = 1 + int(self.withPnBrake)
        t = np.append(body[:, o + 1], tail[0])
        b = np.append(body[:, o + 2], tail[1])
        df = pd.DataFrame({'Time [s]': t, 'Position [m]': self.points.index.values}).set_index('Time [s]')
        df['Velocity [m/s]'] = np.sqrt(b)
        df['Force (el) [N]'] = nanrow(body[:, 0]) * self.totalMass
        df['Force (pnb) [N]'] = nanrow(body[:, 1]) * self.totalMass if self.withPnBrake else np.zeros(N + 1)
        df['Slacks'] = nanrow(body[:, o]) * self.totalMass
        return df


def solve_instances(solvers, terminalTime, initialTime=0, terminalVelocity=1, initialVelocity=1, screen=True, device=None, restart=True):
    """Additive API: one device call for instances that live on DIFFERENT tracks and/or interval counts
    (BASELINE config 5: random tracks, mixed numIntervals).  `solvers` is a list of casadiSolver objects with the same
    problem structure (brakes, power rows, objective, loss family, integrator options); instance i is
    solvers[i].solve(terminalTime[i], ...).  Returns the same dictionary as casadiSolver.solve_batch; z rows are padded
    to the largest interval count."""
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("mseetc_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    n = len(solvers)
    ref = solvers[0]
    sig = lambda s: (s.withPnBrake, s.withRgBrake, s.withPower, s.energyOptimal, s._lossKind, bool(s.opts.integrateLosses)) + s._integrator_key()
    if any(sig(s) != sig(ref) for s in solvers):
        raise ValueError("solve_instances needs solvers with identical problem structure")
    bc = [np.broadcast_to(np.atleast_1d(np.asarray(a, dtype=float)), (n,)) for a in (terminalTime, initialTime, terminalVelocity, initialVelocity)]
    T, t0, vN, v0 = bc
    dev = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
    t_begin = _time.perf_counter()
    Nmax = max(s.numIntervals for s in solvers)
    planes, Ms, tabs = [], [], []
    for i, s in enumerate(solvers):
        if s.energyOptimal and s._lossKind == 'static':
            _, lossT, lossR = classify_losses(s.train)
        else:
            lossT, lossR = 0.0, 0.0
        P, M = s._planes(1, T[i:i + 1], t0[i:i + 1], v0[i:i + 1], vN[i:i + 1], {}, lossT, lossR)
        planes.append(P); Ms.append(float(np.atleast_1d(M)[0]))
        tabs.append(s._tables(s._base['rho'], s._base['g'], s._base['velocityMax']))
    P = np.ascontiguousarray(np.concatenate(planes, axis=1))
    nint = np.array([s.numIntervals for s in solvers], dtype=np.int32)
    trk_off = np.concatenate([[0], np.cumsum(nint)]).astype(np.int32)
    up = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(device=dev, dtype=dt)
    numSteps, numApprox, tableau = ref._integrator()
    guess = ref.initialGuess if restart is not None else 'reference'       # restart=None: this IS the second attempt

    def mk(energy, loss):
        hd = _cabi.Handle(Nmax, ref.withPnBrake, ref.withPower, energy, loss, numSteps, numApprox,
                          int(ref.opts.maxIterations), initial_guess={'reference': 0, 'profile': 1}[guess],
                          stall_iterations=int(ref.stallIterations))
        hd.set_sweep_lanes(0 if ref.sweepLanes == 'auto' else int(ref.sweepLanes))
        if tableau is not None:
            hd.set_integrator(tableau['A'], tableau['w'], tableau['maxIter'])
        if ref.opts.integrateLosses and energy:
            hd.set_integrate_losses(True)
        return hd
    dev_tabs = dict(nint=up(nint, torch.int32), trk_of=up(np.arange(n, dtype=np.int32), torch.int32), trk_off=up(trk_off, torch.int32),
                    ds=up(np.concatenate([t[0] for t in tabs]), torch.float64), c0=up(np.concatenate([t[1] for t in tabs]), torch.float64),
                    bmax=up(np.concatenate([t[2] for t in tabs]), torch.float64))
    tmin = None
    if screen and ref.energyOptimal:
        # minimum trip time of every instance (tracks differ, so there is nothing to share): time-optimal batch first
        ht = mk(False, 0)
        Pt = P.copy()
        lim = [np.minimum(s.points['Speed limit [m/s]'].values[:-1], s._base['velocityMax']) for s in solvers]
        Pt[_cabi.PARAM_INDEX['T_END']] = t0 + 1.5 * np.array([float(np.sum(s.steps / l)) for s, l in zip(solvers, lim)])
        Pt[_cabi.PARAM_INDEX['OBJ_SCALE']] = np.array([s.trackLength / s._base['velocityMax'] for s in solvers])
        tr = ht.solve_device(up(Pt, torch.float64), dev_tabs['nint'], dev_tabs['trk_of'], dev_tabs['trk_off'], dev_tabs['ds'], dev_tabs['c0'],
                             dev_tabs['bmax'])
        stp = 4 + int(ref.withPnBrake)
        idx = torch.from_numpy(nint.astype(np.int64) * stp).to(dev)
        tN = tr['z'].gather(1, idx.unsqueeze(1)).squeeze(1)
        tmin_dev = torch.where(tr['status'] == 0, tN - up(t0, torch.float64), torch.zeros_like(tN))
        tmin = tmin_dev.cpu().numpy()
    else:
        tmin_dev = None
    h = mk(ref.energyOptimal, {'none': 0, 'static': 1, 'dynamic': 2}[ref._lossKind])
    if ref._lossKind == 'dynamic' and ref.energyOptimal:
        dp = ref.train.powerLosses.device_params
        h.set_loss_map(dp['knots_load'], dp['knots_speed'], dp['coef'])
    out = h.solve_device(up(P, torch.float64), dev_tabs['nint'], dev_tabs['trk_of'], dev_tabs['trk_off'], dev_tabs['ds'], dev_tabs['c0'],
                         dev_tabs['bmax'], tmin=tmin_dev)
    out['z'].mul_(((out['status'] == 0) | (out['status'] == 6)).to(out['z'].dtype).unsqueeze(1))
    res = {k: (v.cpu().numpy() if hasattr(v, 'cpu') else v) for k, v in out.items() if v is not None}
    if tmin is not None:
        short = (tmin > 0) & ((T - t0) < tmin * (1 - _cabi.TMIN_MARGIN))
        res['status'][(res['status'] != 0) & short] = 4
        wrong = np.flatnonzero((res['status'] == 4) & ~short)      # early envelope flag without an exact certificate: solve again
        if len(wrong):
            pick = lambda a: np.broadcast_to(np.asarray(a, dtype=float), (n,))[wrong]
            redo = solve_instances([solvers[i] for i in wrong], pick(terminalTime), pick(initialTime), pick(terminalVelocity),
                                   pick(initialVelocity), screen=False, device=device)
            w = redo['z'].shape[1]
            for key in ('obj', 'kkt', 'iters', 'status'):
                res[key][wrong] = redo[key]
            res['z'][wrong] = 0.0
            res['z'][wrong, :w] = redo['z']
    if restart and guess == 'profile':
        # instances whose iteration broke down from the speed-envelope profile: once more from the reference's starting point
        broke = np.flatnonzero((res['status'] == 2) | (res['status'] == 3) | (res['status'] == 5))
        if len(broke):
            pick = lambda a: np.broadcast_to(np.asarray(a, dtype=float), (n,))[broke]
            redo = solve_instances([solvers[i] for i in broke], pick(terminalTime), pick(initialTime), pick(terminalVelocity),
                                   pick(initialVelocity), screen=False, device=device, restart=None)
            w = redo['z'].shape[1]
            better = (redo['status'] == 0) | (redo['status'] == 6)
            for key in ('obj', 'kkt', 'iters', 'status'):
                res[key][broke[better]] = redo[key][better]
            res['z'][broke[better]] = 0.0
            res['z'][broke[better], :w] = redo['z'][better]
            res['restarted'] = broke
    res['wall'] = _time.perf_counter() - t_begin
    scale = P[_cabi.PARAM_INDEX['OBJ_SCALE']]
    M = np.array(Ms)
    res['cost'] = ((1e-6 / 3.6) * M if ref.energyOptimal else 1.0) * res['obj'] * scale
    res['tmin'] = tmin
    res['n_intervals'] = nint
    return res


if __name__ == '__main__':
    from mseetc.train import Train
    from mseetc.track import Track

    train = Train(config={'id': 'NL_Intercity_VIRM6', 'max deceleration': None, 'max acceleration': {'unit': 'm/s^2', 'value': 0.45}})
    track = Track(config={'id': '00_var_speed_limit_100'})
    solver = casadiSolver(train, track, {'numIntervals': 200, 'integrationMethod': 'RK', 'integrationOptions': {'numApproxSteps': 1}})
    df, stats = solver.solve(1541)
    if df is not None:
        print("Objective value = {:.2f} {}".format(stats['Cost'], 'kWh' if solver.opts.energyOptimal else 's'))
    else:
        print("Solver failed!")
