"""Oracle: the loss energies of integrateLosses = True (mseetc/ocp.py:231-241, mseetc/train.py:367-413) -- test infrastructure only.

``TrainIntegrator.calcLosses(v0, dt, Fel, Fpb, grad, curv)`` integrates  x = (v, eTr, eBr),  x(0) = (v0, 0, 0)  over the time dt with
    dv/dt = a(v^2, Fel + Fpb),   d eTr/dt = lossesTr(Fel*M, v)/M,   d eBr/dt = lossesRgb(Fel*M, v)/M          (train.py:375-381)
and returns (eTr, eBr)(dt).  The reference integrates with CVODES (relTol 1e-6, absTol 1e-8; initLosses is called without a solver
argument at ocp.py:120); here classic RK4 with ``steps`` equal steps in torch float64 (8 by default: integration error far below
CVODES's tolerance on smooth right-hand sides; 32 steps where the speed crosses a kink of the spline loss map), derivatives w.r.t. (b0, Fel, Fpb, dt) by autograd -- the rows keep the
reference's dependence on t_k and t_{k+1} through dt = t_{k+1} - t_k (the CUDA path substitutes the shooting function for dt).
"""
import numpy as np
import torch


def energy_fn(sr, c0_of, power_fns, withPn, steps=8, kinks=(), steps_kink=32):
    """Returns fn(b0, Fel, Fpb, dt, c0) -> [(E, grad4, hess10)] for E in (eTr, eBr); grad over (b0, Fel, Fpb, dt), hess pairs
    00 01 02 03 11 12 13 22 23 33.  power_fns(fs, v) -> (PLtr, PLrgb) specific power losses [W/kg] as torch expressions."""
    def fn(b0, Fel, Fpb, dt, c0, b1=None, derivs=True):
        # intervals in which the speed crosses a kink of the loss map (`kinks`: speeds) are integrated with steps_kink steps: the
        # integrand is only continuous there (an adaptive integrator like the reference's CVODES refines there by itself)
        if len(kinks) and b1 is not None:
            va = np.minimum(np.sqrt(b0), np.sqrt(b1)) * 0.97; vb = np.maximum(np.sqrt(b0), np.sqrt(b1)) * 1.03
            hard = np.zeros(len(va), bool)
            for kv in kinks:
                hard |= (va < kv) & (kv < vb)
            if hard.any():
                out = run(b0, Fel, Fpb, dt, c0, steps, derivs)
                sub = run(b0[hard], Fel[hard], Fpb[hard], dt[hard], np.asarray(c0)[hard], steps_kink, derivs)
                for (E, g, H), (Es, gs, Hs) in zip(out, sub):
                    E[hard] = Es
                    for a, b_ in zip(g, gs):
                        a[hard] = b_
                    for a, b_ in zip(H, Hs):
                        a[hard] = b_
                return out
        return run(b0, Fel, Fpb, dt, c0, steps, derivs)

    def run(b0, Fel, Fpb, dt, c0, steps, derivs=True):
        x = [torch.tensor(np.asarray(a, float), dtype=torch.float64, requires_grad=True) for a in (b0, Fel, Fpb, dt)]
        tc0 = torch.tensor(np.asarray(c0, float), dtype=torch.float64)
        F = x[1] + (x[2] if withPn else 0.0 * x[2])

        def rhs(v):
            ptr, prg = power_fns(x[1], v)
            return x[3] * (F - (sr[0] + sr[1] * v + sr[2] * v * v) - tc0), x[3] * ptr, x[3] * prg
        v = torch.sqrt(x[0]); etr = torch.zeros_like(v); erg = torch.zeros_like(v)
        h = 1.0 / steps
        for _ in range(steps):
            k1 = rhs(v); k2 = rhs(v + 0.5 * h * k1[0]); k3 = rhs(v + 0.5 * h * k2[0]); k4 = rhs(v + h * k3[0])
            v = v + h / 6 * (k1[0] + 2 * k2[0] + 2 * k3[0] + k4[0])
            etr = etr + h / 6 * (k1[1] + 2 * k2[1] + 2 * k3[1] + k4[1])
            erg = erg + h / 6 * (k1[2] + 2 * k2[2] + 2 * k3[2] + k4[2])
        out = []
        n = lambda t: t.detach().numpy().copy()
        zero = torch.zeros_like(x[0])
        if not derivs:      # values only (line search)
            return [(n(etr), [], []), (n(erg), [], [])]
        for E in (etr, erg):
            g = torch.autograd.grad(E.sum(), x, create_graph=True, allow_unused=True)
            g = [gi if gi is not None else zero for gi in g]
            H = {}
            for i in range(4):
                if g[i].requires_grad:
                    hi = torch.autograd.grad(g[i].sum(), x, retain_graph=True, allow_unused=True)
                    hi = [hh if hh is not None else zero for hh in hi]
                else:
                    hi = [zero] * 4
                for j in range(i, 4):
                    H[(i, j)] = n(hi[j])
            out.append((n(E), [n(gi) for gi in g], [H[(i, j)] for i in range(4) for j in range(i, 4)]))
        return out
    return fn


def static_power_fns(cT, cR):
    "constant efficiencies: PLtr = cT f v, PLrgb = -cR f v  (train.py:204 + utils.py:197-220, both pieces linear in f)"
    return lambda fs, v: (cT * fs * v, -cR * fs * v)


def dynamic_power_fns(lossmap, M):
    """efficiency.totalLossesFunction split by utils.splitLosses (oracle.lossmap.DynamicLossMap.split).  The motor map is zero outside
    its grid (efficiency.py:40-51,137), whose upper load edge (100 % + 1e-4) is the power hyperbola F v = Pmax -- exactly where the
    power rows of the NLP are active at the nodes.  The end point of a time-domain integration over an interval reproduces the node
    speed only to the integration error, so the last stage of a fixed-step method can land 1e-6 beyond the edge and see zero losses
    with the weight of a whole stage (an adaptive integrator like the reference's CVODES loses a vanishing sliver there).  The load
    is therefore clamped to the edge inside the integration -- the value CVODES effectively integrates."""
    return lambda fs, v: lossmap.split(fs, v, M, clamp_load=True)
