"""Oracle: the implicit Runge-Kutta shooting interval of the reference (TrainIntegrator 'IRK', mseetc/train.py:303-310) --
test infrastructure only.

``ca.simpleIRK(ode, numSteps, order, collMethod, 'fast_newton', {'max_iter': maxIter})`` (CasADi integration_tools) builds, per
step of length dt = h/numSteps, the collocation equations on ``order`` points tau_1..tau_d (Radau or Gauss-Legendre, on (0,1]):

    dt * f(x_j) - sum_{r=0..d} C[r][j] * x_r = 0,   j = 1..d,        x_0 = state at the start of the step,
    x_end = sum_{r=0..d} D[r] * x_r,

with C[r][j] = l_r'(tau_j), D[r] = l_r(1) (Lagrange basis on tau_0 = 0, tau_1..tau_d), solves them with Newton's method from the
guess x_j = x_0 and differentiates the implicitly defined solution.  Restated here literally in that form (the CUDA path uses the
equivalent Runge-Kutta coefficients A, w and Newton passes in jet arithmetic); derivatives by reverse-mode autograd through the
unrolled Newton iterations in torch float64 -- a different technique on purpose.
"""
import numpy as np
import torch


def collocation_points(d, scheme):
    "CasADi collocation_points(d, scheme): roots on (0, 1] of the shifted Legendre ('legendre') / right-Radau ('radau') polynomial."
    from numpy.polynomial import legendre as L
    if scheme == 'legendre':
        x = L.legroots([0] * d + [1])
    elif scheme == 'radau':
        c = np.zeros(d + 1); c[d - 1] = 1.0; c[d] = -1.0        # P_{d-1} - P_d vanishes at x = 1
        x = L.legroots(c)
    else:
        raise ValueError(scheme)
    return np.sort((np.real(x) + 1.0) / 2.0)


def interpolators(tau_root):
    "CasADi collocation_interpolators: C[r][j] = l_r'(tau_j), D[r] = l_r(1)."
    n = len(tau_root)
    C = np.zeros((n, n)); D = np.zeros(n)
    for r in range(n):
        p = np.poly1d([1.0])
        for q in range(n):
            if q != r:
                p *= np.poly1d([1.0, -tau_root[q]]) / (tau_root[r] - tau_root[q])
        D[r] = p(1.0)
        dp = np.polyder(p)
        for j in range(n):
            C[r, j] = dp(tau_root[j])
    return C, D


def _steps(b0, F, h, numSteps, ds, c0, sr, C, D, iters, with_time):
    d = C.shape[0] - 1
    dt = h / numSteps
    acc = lambda b: F.unsqueeze(-1) - (sr[0] + sr[1] * torch.sqrt(b) + sr[2] * b) - c0.unsqueeze(-1)
    Ct = torch.tensor(C)
    b = b0
    t = torch.zeros_like(b0)
    for _ in range(numSteps):
        x = b.unsqueeze(-1).expand(-1, d).clone()
        for _ in range(iters):
            f = 2 * ds.unsqueeze(-1) * acc(x)                                              # db/dsigma at the points
            fp = 2 * ds.unsqueeze(-1) * (-(0.5 * sr[1] / torch.sqrt(x) + sr[2]))
            G = dt * f - (Ct[0, 1:] * b.unsqueeze(-1) + x @ Ct[1:, 1:])                    # G_j, j = 1..d
            J = torch.diag_embed(dt * fp) - Ct[1:, 1:].T                                   # dG_j / dx_r
            x = x - torch.linalg.solve(J, G.unsqueeze(-1)).squeeze(-1)
        if with_time:   # the time rows of the same collocation system are linear in the t_r once the b_r are known
            rhs = dt * ds.unsqueeze(-1) / torch.sqrt(x) - Ct[0, 1:] * t.unsqueeze(-1)
            ts = torch.linalg.solve(Ct[1:, 1:].T.expand(len(b0), d, d), rhs.unsqueeze(-1)).squeeze(-1)
            t = D[0] * t + ts @ torch.tensor(D[1:])
        b = D[0] * b + x @ torch.tensor(D[1:])
    return t, b


def shoot(b0, F, ds, c0, sr, numSteps, numApprox, order, scheme, iters=12):
    """tau = t1 - t0 and b1 of one shooting interval for arrays of intervals, with gradients w.r.t. (b0, F) and the three second
    derivatives (bb, bF, FF): returns (tau, [g_b, g_F], [h_bb, h_bF, h_FF]), (phib, ...)."""
    C, D = interpolators(np.concatenate([[0.0], collocation_points(order, scheme)]))
    tb0 = torch.tensor(np.asarray(b0, float), requires_grad=True)
    tF = torch.tensor(np.asarray(F, float), requires_grad=True)
    tds = torch.tensor(np.broadcast_to(np.asarray(ds, float), tb0.shape).copy())
    tc0 = torch.tensor(np.broadcast_to(np.asarray(c0, float), tb0.shape).copy())
    if numApprox > 0:      # train.py:324-344
        pts = [i / numApprox for i in range(numApprox + 1)]
        bf = [tb0] + [_steps(tb0, tF, p, numSteps, tds, tc0, sr, C, D, iters, False)[1] for p in pts[1:]]
        tau = sum(2 * tds * (pts[i + 1] - pts[i]) / (torch.sqrt(bf[i]) + torch.sqrt(bf[i + 1])) for i in range(numApprox))
        phib = bf[-1]
    else:
        tau, phib = _steps(tb0, tF, 1.0, numSteps, tds, tc0, sr, C, D, iters, True)
    out = []
    for y in (tau, phib):
        gb, gF = torch.autograd.grad(y.sum(), (tb0, tF), create_graph=True)
        hbb, hbF = torch.autograd.grad(gb.sum(), (tb0, tF), retain_graph=True)
        hFF, = torch.autograd.grad(gF.sum(), (tF,), retain_graph=True)
        out.append((y.detach().numpy(), [gb.detach().numpy(), gF.detach().numpy()], [hbb.numpy(), hbF.numpy(), hFF.numpy()]))
    return out


def interval_rows(order, scheme, numSteps, numApprox):
    "Hook for oracle.nlp.ReferenceNLP(interval_fn=...): the 'ct' and 'cb' rows with this integrator."
    def fn(b0, Fel, Fpb, b1, ds, c0, sr, withPn):
        (tau, gt, ht), (phi, gp, hp) = shoot(b0, Fel + Fpb, ds, c0, sr, numSteps, numApprox, order, scheme)
        zero = np.zeros_like(tau)
        res = {}
        for name, val, g, h, lin in (('ct', -tau, gt, ht, 0.0), ('cb', b1 - phi, gp, hp, 1.0)):
            gb, gF = -g[0], -g[1]
            hbb, hbF, hFF = -h[0], -h[1], -h[2]
            gP, hbP, hFP, hPP = (gF, hbF, hFF, hFF) if withPn else (zero, zero, zero, zero)
            # HESS_PAIRS over (b0, Fel, Fpb, b1): 00 01 02 03 11 12 13 22 23 33
            res[name] = (val, [gb, gF, gP, zero + lin], [hbb, hbF, hbP, zero, hFF, hFP, zero, hPP, zero, zero])
        return res
    return fn
