"""Oracle: the reference's dynamic loss map (efficiency.py) and the two loss-epigraph rows built from it —
test infrastructure only.

Restates ``mseetc/efficiency.py:7-12`` (forceToLoad), ``:23-51`` (createSpline: CasADi ``interpolant('bspline')`` =
cubic not-a-knot tensor B-spline, 0 outside the grid; built here with scipy's RectBivariateSpline, a different
construction than the product's), ``:54-98`` (table = min(A,B)*4 motors, speeds from the Hz grid), ``:101-141``
(gear + motor + auxiliaries + transformer, zero where the motor map is zero) and ``mseetc/utils.py:197-220``
(splitLosses: tangent extension at f = +-1e-10).  Everything is evaluated in torch float64 and differentiated with
autograd (first, second and the mixed third derivatives needed through the tangent slope) — deliberately not the
hand-derived chain rules of the CUDA path.
"""
import json
import os

import numpy as np
import torch
from scipy.interpolate import RectBivariateSpline

_DATA = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'ms-eetc_b200', 'mseetc', 'motor_losses.json')
R_TRAFO, V_CAT = 10.0, 15000.0
TOL = 1e-10           # utils.py:200


def _hz_to_ms(f):
    return (((np.asarray(f, float) - 20.0) / (170.0 - 20.0)) * (160.0 - 20.0) + 20.0) / 3.6      # efficiency.py:56-62


class DynamicLossMap:
    def __init__(self, forceMax, auxiliaries=27000.0, etaGear=1.0, tableScale=1.0):
        raw = json.load(open(_DATA))
        table = np.minimum(np.array(raw['config_a_losses_w']), np.array(raw['config_b_losses_w'])) * 4 * tableScale   # [load, freq]
        loads = np.array(raw['loads_percent'], float)
        loads[-1] += 1e-4                                                       # efficiency.py:27-28
        speeds = _hz_to_ms(raw['frequencies_hz'])
        spl = RectBivariateSpline(loads, speeds, table, kx=3, ky=3, s=0)
        tx, ty = spl.get_knots()
        self.tx, self.ty = torch.tensor(tx, dtype=torch.float64), torch.tensor(ty, dtype=torch.float64)
        self.coef = torch.tensor(spl.get_coeffs().reshape(len(tx) - 4, len(ty) - 4), dtype=torch.float64)
        self.box = (loads[0], loads[-1], speeds[0], speeds[-1])
        self.forceMax = float(forceMax)
        self.powerMax = float(forceMax) * float(_hz_to_ms(55.0))               # efficiency.py:64-65
        self.aux, self.etaGear = float(auxiliaries), float(etaGear)

    # ---- cubic B-spline basis (de Boor), 4 non-zero functions at x, differentiable in x
    @staticmethod
    def _basis(t, x):
        n = len(t) - 4
        i = torch.clamp(torch.searchsorted(t, x.detach(), right=True) - 1, 3, n - 1)
        N = [torch.ones_like(x)]
        for d in range(1, 4):
            new = [torch.zeros_like(x) for _ in range(d + 1)]
            for r in range(d):
                tl, tr = t[i + r + 1 - d], t[i + r + 1]
                tmp = N[r] / (tr - tl)
                new[r] = new[r] + (tr - x) * tmp
                new[r + 1] = new[r + 1] + (x - tl) * tmp
            N = new
        return i, N

    def spline(self, load, v):
        inside = (load >= self.box[0]) & (load <= self.box[1]) & (v >= self.box[2]) & (v <= self.box[3])
        lc = torch.clamp(load, self.box[0], self.box[1]) + 0 * load
        vc = torch.clamp(v, self.box[2], self.box[3]) + 0 * v
        # keep the derivative inside the box (clamp has zero slope only outside, where the value is 0 anyway)
        i, Nx = self._basis(self.tx, lc)
        j, Ny = self._basis(self.ty, vc)
        val = torch.zeros_like(load)
        for a in range(4):
            for b in range(4):
                val = val + self.coef[i - 3 + a, j - 3 + b] * Nx[a] * Ny[b]
        return torch.where(inside, val, torch.zeros_like(val))

    def motor(self, f, v, clamp_load=False):
        vMin, vMax = self.box[2], self.box[3]
        vc = torch.clamp(v, vMin, vMax)                                         # efficiency.py:40 (zero slope outside)
        absf = torch.where(f >= 0, f, -f)                                       # efficiency.py:42 (slope +1 at f = 0, as CasADi)
        tp = self.powerMax / self.forceMax
        load = torch.where(vc <= tp, 100 * absf / self.forceMax, 100 * absf * vc / self.powerMax)   # efficiency.py:10-12
        if clamp_load:      # see oracle/intlosses.py: stage points of a time integration that overshoot the power hyperbola by rounding
            load = torch.where((load > self.box[1]) & (load <= 1.001 * self.box[1]), torch.full_like(load, self.box[1]), load)
        return self.spline(load, vc)

    def total(self, f, v, clamp_load=False):
        "efficiency.py:103-139, absolute units: force [N], losses [W]"
        tr = f >= 0
        pw_t, pw_b = f * v, -f * v
        gear = torch.where(tr, ((1 - self.etaGear) / self.etaGear) * pw_t, (1 - self.etaGear) * pw_b)
        mot = self.motor(f, v, clamp_load)
        pm_t = pw_t + gear + mot + self.aux
        pm_b = pw_b - gear - mot - self.aux
        arg = torch.where(tr, V_CAT ** 2 - 4 * R_TRAFO * pm_t, V_CAT ** 2 + 4 * R_TRAFO * pm_b)
        trafo = (V_CAT - torch.sqrt(torch.clamp(arg, min=1.0))) ** 2 / (4 * R_TRAFO)
        tot = gear + mot + self.aux + trafo
        return torch.where(mot > 0, tot, torch.zeros_like(tot))

    def specific(self, fs, v, M, clamp_load=False):
        return self.total(fs * M, v, clamp_load) / M                            # train.py:216

    def split(self, fs, v, M, clamp_load=False):
        "utils.py:197-220: (funTr, funRgb) of specific force fs and speed v"
        def slope(sign):
            ft = torch.full_like(v, sign * TOL, requires_grad=True)
            (g,) = torch.autograd.grad(self.specific(ft, v, M).sum(), ft, create_graph=True)
            return g
        beta = self.specific(torch.zeros_like(v), v, M)
        full = self.specific(fs, v, M, clamp_load)
        fun_tr = torch.where(fs >= 0, full, slope(+1.0) * fs + beta)
        fun_rg = torch.where(fs < 0, full, slope(-1.0) * fs + beta)
        return fun_tr, fun_rg

    def rows(self, M):
        """Hook for ReferenceNLP(loss_rows=...): the nonlinear part  -PL{tr,rgb}(Fel, vMid)/vMid  of the two rows
        (ocp.py:221-229) with gradient (Fel,b0,b1) and Hessian (FF,F0,F1,00,01,11)."""
        def fn(Fel, b0, b1):
            x = [torch.tensor(np.asarray(a, float), dtype=torch.float64, requires_grad=True) for a in (Fel, b0, b1)]
            vmid = (torch.sqrt(x[1]) + torch.sqrt(x[2])) / 2
            ftr, frg = self.split(x[0], vmid, M)
            out = []
            for r in (-ftr / vmid, -frg / vmid):
                g = torch.autograd.grad(r.sum(), x, create_graph=True, allow_unused=True)
                g = [gi if gi is not None else torch.zeros_like(x[0]) for gi in g]
                H = []
                for gi in g:
                    if gi.requires_grad:
                        h = torch.autograd.grad(gi.sum(), x, retain_graph=True, allow_unused=True)
                        H.append([hi if hi is not None else torch.zeros_like(x[0]) for hi in h])
                    else:
                        H.append([torch.zeros_like(x[0])] * 3)
                n = lambda t: t.detach().numpy()
                out.append((n(r), [n(g[0]), n(g[1]), n(g[2])],
                            [n(H[0][0]), n(H[0][1]), n(H[0][2]), n(H[1][1]), n(H[1][2]), n(H[2][2])]))
            return out
        return fn
