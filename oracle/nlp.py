"""Oracle: the reference NLP (variables, rows, bounds, objective) — test infrastructure only.

Exact layout of ``mseetc/ocp.py:136-284``:
  z = [Fel_0,(Fpb_0),s_0,t_0,b_0, ..., Fel_{N-1},(Fpb_{N-1}),s_{N-1},t_{N-1},b_{N-1}, t_N,b_N]
  g rows per interval i (ocp.py:183-241):
     [Fel_i*sqrt(b_i), Fel_i*sqrt(b_{i+1})]      (only if a power limit exists, :184-192)
     a(b_i,Fel_i,Fpb_i)                           (:199-201)
     t_{i+1}-Phi_t, b_{i+1}-Phi_b                 (:204-213)
     s_i - PLtr(Fel_i,vMid)/vMid, s_i - PLrgb(..)/vMid   (energy-optimal, :216-229)
Stage functions and all their first/second derivatives are generated with sympy from the
formulas of ``train.py:251-259`` (ODE), ``train.py:298-301`` + CasADi ``simpleRK`` (classic RK4,
numSteps equal steps), ``train.py:324-344`` (ERK4+ time approximation) — deliberately a
different derivative technique than the hand-written forward-mode jets of the CUDA path.
"""
import functools

import numpy as np
import scipy.sparse as sp
import sympy as sy

from .problem import curvature_resistance

ACC_INF = 10.0  # ocp.py:104


@functools.lru_cache(maxsize=None)
def _stage_functions(withPn, numSteps, numApprox, lossKind):
    b0, Fel, Fpb, b1 = sy.symbols('b0 Fel Fpb b1', real=True)
    ds, c0, sr0, sr1, sr2, cT, cR = sy.symbols('ds c0 sr0 sr1 sr2 cT cR', real=True)
    F = Fel + (Fpb if withPn else 0)

    def acc(b):  # train.py:251-254
        return F - (sr0 + sr1 * sy.sqrt(b) + sr2 * b) - c0

    def rk4(x, f, h, n):  # CasADi simpleRK: n equal steps of classic RK4 over horizon h
        dt = h / n
        for _ in range(n):
            k1 = f(x)
            k2 = f([xi + dt / 2 * ki for xi, ki in zip(x, k1)])
            k3 = f([xi + dt / 2 * ki for xi, ki in zip(x, k2)])
            k4 = f([xi + dt * ki for xi, ki in zip(x, k3)])
            x = [xi + dt / 6 * (a + 2 * b + 2 * c + d) for xi, a, b, c, d in zip(x, k1, k2, k3, k4)]
        return x

    if numApprox > 0:  # train.py:324-344 (b only by RK4, time by the trapezoid-in-1/v rule)
        fb = lambda x: [2 * ds * acc(x[0])]
        pts = [sy.Integer(0)] + [sy.Rational(i, numApprox) for i in range(1, numApprox + 1)]
        bf = [b0] + [rk4([b0], fb, p, numSteps)[0] for p in pts[1:]]
        tau = sum(2 * ds * (pts[i + 1] - pts[i]) / (sy.sqrt(bf[i]) + sy.sqrt(bf[i + 1])) for i in range(numApprox))
        phib = bf[-1]
    else:  # train.py:298-299: RK4 on (t, b)
        ftb = lambda x: [ds / sy.sqrt(x[1]), 2 * ds * acc(x[1])]
        tt, phib = rk4([sy.Integer(0), b0], ftb, sy.Integer(1), numSteps)
        tau = tt

    vMid = (sy.sqrt(b0) + sy.sqrt(b1)) / 2  # ocp.py:221
    rows = {
        'p0': Fel * sy.sqrt(b0),   # ocp.py:189
        'p1': Fel * sy.sqrt(b1),
        'acc': acc(b0),            # ocp.py:199
        'ct': -tau,                # + t1 - t0 added linearly (ocp.py:207-210)
        'cb': b1 - phib,
    }
    if lossKind == 'static':       # train.py:204 + utils.py:197-220 (both pieces linear in f)
        rows['ltr'] = -(Fel * vMid * cT) / vMid   # + s
        rows['lrg'] = -(-cR * Fel * vMid) / vMid
    elif lossKind == 'none':
        rows['ltr'] = sy.Integer(0) * Fel
        rows['lrg'] = sy.Integer(0) * Fel
    v = (b0, Fel, Fpb, b1)
    exprs = []
    for name in rows:
        e = rows[name]
        exprs.append(e)
        gr = [sy.diff(e, vi) for vi in v]
        exprs += gr
        for i in range(4):
            for j in range(i, 4):
                exprs.append(sy.diff(gr[i], v[j]))
    fn = sy.lambdify((b0, Fel, Fpb, b1, ds, c0, sr0, sr1, sr2, cT, cR), exprs, modules='numpy', cse=True)
    return list(rows.keys()), fn


HESS_PAIRS = [(i, j) for i in range(4) for j in range(i, 4)]  # over (b0, Fel, Fpb, b1)


class ReferenceNLP:
    """min f(z) s.t. lbg<=g(z)<=ubg, lbz<=z<=ubz exactly as assembled in ocp.py:166-284."""

    def __init__(self, train, pos, grad_permil, limit, curv, track_length, opts=None, loss_rows=None, interval_fn=None, energy_fn=None):
        o = dict(numIntervals=len(pos) - 1, energyOptimal=True, minimumVelocity=1.0, numSteps=1, numApproxSteps=0)
        o.update(opts or {})
        self.opts = o
        N = self.N = len(pos) - 1
        self.pos = np.asarray(pos, float)
        self.ds = np.diff(self.pos)
        self.limit = np.asarray(limit, float)
        self.train = train
        M = self.M = train.mass * train.rho                      # ocp.py:96-97
        self.withRg = train.forceMin != 0                        # ocp.py:101-102
        self.withPn = train.forceMinPn != 0
        self.energy = bool(o['energyOptimal'])
        self.vmin = o['minimumVelocity']
        self.forceMax = train.forceMax / M if train.forceMax is not None else ACC_INF   # ocp.py:106-108
        self.forceMin = train.forceMin / M if train.forceMin is not None else -ACC_INF
        self.forceMinPn = train.forceMinPn / M if train.forceMinPn is not None else -ACC_INF
        pMax = train.powerMax / M if train.powerMax is not None else None            # ocp.py:110-111
        pMin = train.powerMin / M if train.powerMin is not None else None
        self.accMax = min(ACC_INF, train.accMax if train.accMax is not None else ACC_INF)   # ocp.py:113-114
        self.accMin = max(-ACC_INF, -abs(train.accMin) if train.accMin is not None else -ACC_INF)
        self.withPower = pMax is not None or pMin is not None                         # ocp.py:184
        if self.withPower:
            up = pMax if pMax is not None else self.forceMax * train.velocityMax      # ocp.py:186-187
            lo = 0 if not self.withRg else pMin if pMin is not None else self.forceMin * train.velocityMax
            self.pUp, self.pLo = abs(up), -abs(lo)
        self.sr = (train.r0 / M, train.r1 / M, train.r2 / M)                          # train.py:181-183
        self.c0 = train.g * (np.asarray(grad_permil, float)[:N] / 1e3) / train.rho \
            + curvature_resistance(np.asarray(curv, float)[:N], train.g) / train.rho  # train.py:254
        losses = train.losses if train.losses is not None else ('none',)
        self.lossKind = losses[0]
        self.cT = self.cR = 0.0
        if self.lossKind == 'static':
            self.cT = (1 - losses[1]) / losses[1]
            self.cR = (1 - losses[2])
        # integrateLosses = True (ocp.py:231-241): energy_fn = oracle.intlosses.energy_fn(...); rows s_i - E(b_i, Fel_i, Fpb_i, t_{i+1} - t_i)
        self.energy_fn = energy_fn
        self.interval_fn = interval_fn  # None: explicit RK4 (sympy, below); else oracle.irk.interval_rows(...) for the 'IRK' integrator
        self.loss_rows = loss_rows  # callable (Fel, b0, b1) -> ((val,grad3,hess6) x2) for non-symbolic maps
        if self.energy:
            self.scale = 3.6 / (1e-6 * M)                                             # ocp.py:278
        else:
            self.scale = track_length / train.velocityMax                            # ocp.py:282
        # (with an interval_fn the symbolic shooting rows are replaced in _stage: build the cheapest variant, sympy's expressions of
        # nested RK4 steps grow beyond use after two steps)
        rk = (int(o['numSteps']), int(o['numApproxSteps'])) if interval_fn is None else (1, 1)
        self.names, self.fn = _stage_functions(self.withPn, rk[0], rk[1], self.lossKind if self.lossKind in ('static', 'none') else 'none')
        # ---- index maps
        nu = self.nu = 1 + int(self.withPn)
        st = self.stride = 3 + nu
        self.nz = N * st + 2
        k = np.arange(N)
        self.iFel = k * st
        self.iFpb = k * st + 1 if self.withPn else None
        self.iS = k * st + nu
        self.iT = np.concatenate([k * st + nu + 1, [N * st]])
        self.iB = np.concatenate([k * st + nu + 2, [N * st + 1]])
        rows_per = (2 if self.withPower else 0) + 1 + 2 + (2 if self.energy else 0)
        self.rows_per = rows_per
        self.ng = N * rows_per
        r = k * rows_per
        off = 0
        self.rP0 = self.rP1 = None
        if self.withPower:
            self.rP0, self.rP1 = r, r + 1
            off = 2
        self.rAcc = r + off
        self.rCt = r + off + 1
        self.rCb = r + off + 2
        self.rLtr = self.rLrg = None
        if self.energy:
            self.rLtr, self.rLrg = r + off + 3, r + off + 4

    # ------------------------------------------------------------------ bounds / start
    def bounds(self, T, t0=0.0, v0=1.0, vN=1.0):
        N = self.N
        v0 = min(max(v0, self.vmin), self.limit[0])      # ocp.py:343-344
        vN = min(max(vN, self.vmin), self.limit[-1])
        lbz = np.empty(self.nz)
        ubz = np.empty(self.nz)
        lbz[self.iFel] = self.forceMin if self.withRg else 0.0      # ocp.py:175-176
        ubz[self.iFel] = self.forceMax
        if self.withPn:
            lbz[self.iFpb] = self.forceMinPn
            ubz[self.iFpb] = 0.0
        lbz[self.iS] = 0.0
        ubz[self.iS] = np.inf
        lim = np.minimum(self.limit[1:N], self.train.velocityMax)   # ocp.py:266-269
        lim = np.minimum(lim, self.limit[0:N - 1])
        lbz[self.iT] = t0
        ubz[self.iT] = T
        lbz[self.iB[1:N]] = self.vmin ** 2
        ubz[self.iB[1:N]] = lim ** 2
        ubz[self.iT[0]] = t0
        lbz[self.iB[0]] = ubz[self.iB[0]] = v0 ** 2
        lbz[self.iB[N]] = ubz[self.iB[N]] = vN ** 2
        lbg = np.zeros(self.ng)
        ubg = np.zeros(self.ng)
        if self.withPower:
            for r in (self.rP0, self.rP1):
                lbg[r] = self.pLo
                ubg[r] = self.pUp
        lbg[self.rAcc] = self.accMin
        ubg[self.rAcc] = self.accMax
        if self.energy:
            ubg[self.rLtr] = np.inf
            ubg[self.rLrg] = np.inf
        return lbz, ubz, lbg, ubg

    def x0(self, T, t0=0.0):
        """ocp.py:325-339."""
        z = np.empty(self.nz)
        z[self.iFel] = 0.5
        if self.withPn:
            z[self.iFpb] = -0.1
        z[self.iS] = 1.0
        z[self.iT] = t0 + np.arange(self.N + 1) * ((T - t0) / self.N)
        z[self.iB] = (60 / 3.6) ** 2
        return z

    # ------------------------------------------------------------------ evaluation
    def _stage(self, z):
        b = z[self.iB]
        Fel = z[self.iFel]
        Fpb = z[self.iFpb] if self.withPn else np.zeros(self.N)
        out = self.fn(b[:-1], Fel, Fpb, b[1:], self.ds, self.c0, *self.sr, self.cT, self.cR)
        res = {}
        per = 1 + 4 + 10
        for r, name in enumerate(self.names):
            blk = [np.broadcast_to(np.asarray(x, float), (self.N,)) for x in out[r * per:(r + 1) * per]]
            res[name] = (blk[0], blk[1:5], blk[5:])
        if self.energy_fn is not None and self.energy:      # the rows are assembled separately (they involve t_i, t_{i+1})
            zero = np.zeros(self.N)
            res['ltr'] = res['lrg'] = (zero, [zero] * 4, [zero] * 10)
        if self.interval_fn is not None:     # train.py:303-310: the shooting rows from the collocation integrator
            res.update(self.interval_fn(b[:-1], Fel, Fpb, b[1:], self.ds, self.c0, self.sr, self.withPn))
        if self.loss_rows is not None and self.energy:
            # numeric loss rows supplied as functions of (Fel, b0, b1): map into the (b0,Fel,Fpb,b1) slots
            for name, (val, g3, h6) in zip(('ltr', 'lrg'), self.loss_rows(Fel, b[:-1], b[1:])):
                zero = np.zeros(self.N)
                gF, g0, g1 = g3
                hFF, hF0, hF1, h00, h01, h11 = h6
                grad = [g0, gF, zero, g1]
                hess = [h00, hF0, zero, h01, hFF, zero, hF1, zero, zero, h11]
                res[name] = (val, grad, hess)
        return res

    def f(self, z):
        Fel = z[self.iFel]
        if self.energy:   # ocp.py:223,243-245
            if self.energy_fn is not None:   # ocp.py:235
                obj = np.sum(self.ds * Fel + z[self.iS]) + 1e-3 * np.sum(np.diff(Fel) ** 2)
            else:
                obj = np.sum(self.ds * (Fel + z[self.iS])) + 1e-3 * np.sum(np.diff(Fel) ** 2)
        else:             # ocp.py:146-150
            obj = z[self.iT[-1]] + 1e-4 * (Fel @ Fel + (z[self.iFpb] @ z[self.iFpb] if self.withPn else 0.0))
        return obj / self.scale

    def grad_f(self, z):
        gr = np.zeros(self.nz)
        Fel = z[self.iFel]
        if self.energy:
            gr[self.iFel] = self.ds
            gr[self.iS] = self.ds if self.energy_fn is None else 1.0
            d = np.diff(Fel)
            gr[self.iFel[1:]] += 2e-3 * d
            gr[self.iFel[:-1]] -= 2e-3 * d
        else:
            gr[self.iT[-1]] = 1.0
            gr[self.iFel] = 2e-4 * Fel
            if self.withPn:
                gr[self.iFpb] = 2e-4 * z[self.iFpb]
        return gr / self.scale

    def g(self, z):
        s = self._stage(z)
        g = np.empty(self.ng)
        if self.withPower:
            g[self.rP0] = s['p0'][0]
            g[self.rP1] = s['p1'][0]
        g[self.rAcc] = s['acc'][0]
        t = z[self.iT]
        g[self.rCt] = t[1:] - t[:-1] + s['ct'][0]
        g[self.rCb] = s['cb'][0]
        if self.energy:
            g[self.rLtr] = z[self.iS] + s['ltr'][0]
            g[self.rLrg] = z[self.iS] + s['lrg'][0]
            if self.energy_fn is not None:
                (etr, _, _), (erg, _, _) = self._energies(z, derivs=False)
                g[self.rLtr] -= etr
                g[self.rLrg] -= erg
        return g

    def _energies(self, z, derivs=True):
        "loss energies of all intervals at z (memoised per point: g, jac and hess of one iterate share one evaluation)"
        key = (z.tobytes(), derivs)
        memo = self.__dict__.setdefault('_energy_memo', {})
        if key not in memo:
            if derivs is False and (z.tobytes(), True) in memo:
                return memo[(z.tobytes(), True)]
            if len(memo) > 8:
                memo.clear()
            b, t = z[self.iB], z[self.iT]
            Fpb = z[self.iFpb] if self.withPn else np.zeros(self.N)
            memo[key] = self.energy_fn(b[:-1], z[self.iFel], Fpb, t[1:] - t[:-1], self.c0, b[1:], derivs)
        return memo[key]

    def _energy_cols(self):
        "columns and signs of the arguments (b0, Fel, Fpb, dt = t1 - t0) of the energy rows: dt contributes through t1 (+) and t0 (-)"
        return [(self.iB[:-1], 0, 1.0), (self.iFel, 1, 1.0)] + ([(self.iFpb, 2, 1.0)] if self.withPn else []) + \
               [(self.iT[1:], 3, 1.0), (self.iT[:-1], 3, -1.0)]

    def _rowspec(self):
        spec = []
        if self.withPower:
            spec += [('p0', self.rP0), ('p1', self.rP1)]
        spec += [('acc', self.rAcc), ('ct', self.rCt), ('cb', self.rCb)]
        if self.energy:
            spec += [('ltr', self.rLtr), ('lrg', self.rLrg)]
        return spec

    def _varcols(self):
        cols = [self.iB[:-1], self.iFel, self.iFpb if self.withPn else None, self.iB[1:]]
        return cols

    def jac(self, z):
        s = self._stage(z)
        cols = self._varcols()
        R, C, V = [], [], []
        for name, rows in self._rowspec():
            _, gr, _ = s[name]
            for vi in range(4):
                if cols[vi] is None:
                    continue
                R.append(rows)
                C.append(cols[vi])
                V.append(gr[vi])
        ones = np.ones(self.N)
        R += [self.rCt, self.rCt]
        C += [self.iT[1:], self.iT[:-1]]
        V += [ones, -ones]
        if self.energy:
            R += [self.rLtr, self.rLrg]
            C += [self.iS, self.iS]
            V += [ones, ones]
        if self.energy and self.energy_fn is not None:
            for rows, (_, gr, _) in zip((self.rLtr, self.rLrg), self._energies(z)):
                for cols_, a, sg in self._energy_cols():
                    R.append(rows); C.append(cols_); V.append(-sg * gr[a])
        J = sp.coo_matrix((np.concatenate(V), (np.concatenate(R), np.concatenate(C))), shape=(self.ng, self.nz))
        return J.tocsr()

    def hess(self, z, sigma, lam):
        """Full symmetric Hessian of sigma*f + lam'g."""
        s = self._stage(z)
        cols = self._varcols()
        R, C, V = [], [], []
        for name, rows in self._rowspec():
            _, _, hs = s[name]
            lr = lam[rows]
            for (i, j), h in zip(HESS_PAIRS, hs):
                if cols[i] is None or cols[j] is None:
                    continue
                R.append(cols[i]); C.append(cols[j]); V.append(lr * h)
                if i != j:
                    R.append(cols[j]); C.append(cols[i]); V.append(lr * h)
        if self.energy and self.energy_fn is not None:
            pair = {(i, j): q for q, (i, j) in enumerate(HESS_PAIRS)}
            for rows, (_, _, hs) in zip((self.rLtr, self.rLrg), self._energies(z)):
                lr = lam[rows]
                ec = self._energy_cols()
                for c1, a1, s1 in ec:
                    for c2, a2, s2 in ec:
                        h = hs[pair[(min(a1, a2), max(a1, a2))]]
                        R.append(c1); C.append(c2); V.append(-lr * s1 * s2 * h)
        w = sigma / self.scale
        if self.energy:
            a, b = self.iFel[1:], self.iFel[:-1]
            e = np.full(self.N - 1, 2e-3 * w)
            R += [a, b, a, b]; C += [a, b, b, a]; V += [e, e, -e, -e]
        else:
            R.append(self.iFel); C.append(self.iFel); V.append(np.full(self.N, 2e-4 * w))
            if self.withPn:
                R.append(self.iFpb); C.append(self.iFpb); V.append(np.full(self.N, 2e-4 * w))
        H = sp.coo_matrix((np.concatenate(V), (np.concatenate(R), np.concatenate(C))), shape=(self.nz, self.nz))
        return H.tocsr()

    # ------------------------------------------------------------------ result slicing (ocp.py:361,376-405)
    def cost(self, fval):
        return ((1e-6 / 3.6) * self.M if self.energy else 1.0) * fval * self.scale

    def unpack(self, z):
        return dict(t=z[self.iT], b=z[self.iB], Fel=z[self.iFel],
                    Fpb=z[self.iFpb] if self.withPn else np.zeros(self.N), s=z[self.iS])
