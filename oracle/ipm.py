"""Oracle: generic primal-dual interior-point filter line-search solver — test infrastructure only.

Restates the published IPOPT algorithm (Waechter & Biegler, Math. Prog. 106, 2006) with IPOPT 3.14
default parameters, because the reference passes only ``max_iter`` (``mseetc/ocp.py:290``).  The KKT
systems are solved on the *full sparse* primal-dual matrix with ``scipy.sparse.linalg.splu`` in the
reference's own variable/row ordering — deliberately unrelated to the Riccati recursion of the CUDA
path, so the two can check each other.

Known deviations from IPOPT proper (documented, not hidden):
  * no restoration phase: when the line search fails the solver returns ``Restoration_Failed``;
  * inertia is not available from LU: a curvature test along the step triggers the
    delta_w regularisation ladder instead;
  * "acceptable level" termination with IPOPT's defaults (1e-6 for 15 iterations), reported as success like CasADi does.
"""
import time

import numpy as np
import scipy.sparse as sp  # noqa: E402

np.seterr(invalid='ignore', divide='ignore')  # NaN trial points are rejected by the line search
import scipy.sparse.linalg as spla

EPS = np.finfo(float).eps


class Result(dict):
    __getattr__ = dict.get


def solve(nlp, x0, lbz, ubz, lbg, ubg, max_iter=500, tol=1e-8, mu_init=0.1, scaling=True, verbose=False,
          soc=True, ls_mult=True, relax=1e-8):
    t_start = time.perf_counter()
    n_all, m_all = len(x0), len(lbg)
    fixed = lbz == ubz                                   # fixed_variable_treatment = make_parameter
    free = np.where(~fixed)[0]
    xfull = np.array(x0, float)
    xfull[fixed] = lbz[fixed]
    n = len(free)
    eqr = np.where(lbg == ubg)[0]
    inr = np.where(lbg != ubg)[0]
    mE, mI = len(eqr), len(inr)

    # ---- gradient based scaling at the starting point (nlp_scaling_max_gradient = 100)
    J0 = nlp.jac(xfull)[:, free]
    if scaling:
        gmax = np.abs(nlp.grad_f(xfull)[free]).max()
        sf = 100.0 / gmax if gmax > 100 else 1.0
        rowmax = np.maximum(abs(J0).max(axis=1).toarray().ravel(), 1e-300)
        sg = np.where(rowmax > 100, 100.0 / rowmax, 1.0)
        sf, sg = max(sf, 1e-8), np.maximum(sg, 1e-8)
    else:
        sf, sg = 1.0, np.ones(m_all)
    Sg = sp.diags(sg)

    # ---- bounds (bound_relax_factor = 1e-8)
    xL, xU = lbz[free].copy(), ubz[free].copy()
    dL, dU = (lbg * sg)[inr], (ubg * sg)[inr]
    cE = (lbg * sg)[eqr]
    xL = np.where(np.isfinite(xL), xL - relax * np.maximum(1, np.abs(xL)), xL)
    xU = np.where(np.isfinite(xU), xU + relax * np.maximum(1, np.abs(xU)), xU)
    dL = np.where(np.isfinite(dL), dL - relax * np.maximum(1, np.abs(dL)), dL)
    dU = np.where(np.isfinite(dU), dU + relax * np.maximum(1, np.abs(dU)), dU)
    hxL, hxU, hdL, hdU = np.isfinite(xL), np.isfinite(xU), np.isfinite(dL), np.isfinite(dU)

    def push(v, L, U, hL, hU, k1=1e-2, k2=1e-2):       # bound_push / bound_frac
        v = v.copy()
        both = hL & hU
        pL = np.where(both, np.minimum(k1 * np.maximum(1, np.abs(L)), k2 * (U - L)), k1 * np.maximum(1, np.abs(L)))
        pU = np.where(both, np.minimum(k1 * np.maximum(1, np.abs(U)), k2 * (U - L)), k1 * np.maximum(1, np.abs(U)))
        v = np.where(hL, np.maximum(v, L + pL), v)
        v = np.where(hU, np.minimum(v, U - pU), v)
        return v

    def evalg(x):
        xfull[free] = x
        g = nlp.g(xfull) * sg
        return g[eqr] - cE, g[inr]

    def evalf(x):
        xfull[free] = x
        return nlp.f(xfull) * sf

    x = push(xfull[free], xL, xU, hxL, hxU)
    c, d = evalg(x)
    w = push(d, dL, dU, hdL, hdU)
    zL = np.where(hxL, 1.0, 0.0); zU = np.where(hxU, 1.0, 0.0)
    vL = np.where(hdL, 1.0, 0.0); vU = np.where(hdU, 1.0, 0.0)
    mu = mu_init
    kd = 1e-5                                             # kappa_d damping of one-sided bounds
    oneL_x, oneU_x = hxL & ~hxU, hxU & ~hxL
    oneL_w, oneU_w = hdL & ~hdU, hdU & ~hdL

    def derivs(x):
        xfull[free] = x
        gf = nlp.grad_f(xfull)[free] * sf
        J = (Sg @ nlp.jac(xfull))[:, free].tocsr()
        return gf, J[eqr], J[inr]

    def hessL(x, yc, yd):
        xfull[free] = x
        lam = np.zeros(m_all)
        lam[eqr] = yc; lam[inr] = yd
        return nlp.hess(xfull, sf, lam * sg)[free][:, free]

    gf, Jc, Jd = derivs(x)
    # least-square multiplier estimate (constr_mult_init_max = 1000)
    yc = np.zeros(mE); yd = np.zeros(mI)
    try:
        if not ls_mult:
            raise ValueError
        A = sp.bmat([[sp.identity(n + mI), sp.bmat([[Jc.T, Jd.T], [None, -sp.identity(mI)]])],
                     [sp.bmat([[Jc, None], [Jd, -sp.identity(mI)]]), None]], format='csc')
        rhs = -np.concatenate([gf - zL + zU, -vL + vU, np.zeros(mE + mI)])
        sol = spla.splu(A).solve(rhs)
        y = sol[n + mI:]
        if np.abs(y).max() <= 1000:
            yc, yd = y[:mE], y[mE:]
    except Exception:
        pass

    def slacks(x, w):
        return np.where(hxL, x - xL, 1.0), np.where(hxU, xU - x, 1.0), np.where(hdL, w - dL, 1.0), np.where(hdU, dU - w, 1.0)

    def barrier(x, w, mu, fval):
        sxL, sxU, swL, swU = slacks(x, w)
        if min(sxL.min(), sxU.min(), swL.min(initial=1), swU.min(initial=1)) <= 0:
            return np.inf
        phi = fval - mu * (np.log(sxL[hxL]).sum() + np.log(sxU[hxU]).sum() + np.log(swL[hdL]).sum() + np.log(swU[hdU]).sum())
        phi += kd * mu * (sxL[oneL_x].sum() + sxU[oneU_x].sum() + swL[oneL_w].sum() + swU[oneU_w].sum())
        return phi

    def errors(gf, Jc, Jd, c, d, x, w, yc, yd, zL, zU, vL, vU, mu):
        sxL, sxU, swL, swU = slacks(x, w)
        rx = gf + Jc.T @ yc + Jd.T @ yd - zL + zU
        rw = -yd - vL + vU
        dinf = max(np.abs(rx).max(), np.abs(rw).max(initial=0))
        pinf = max(np.abs(c).max(initial=0), np.abs(d - w).max(initial=0))
        comp = np.concatenate([(sxL * zL)[hxL], (sxU * zU)[hxU], (swL * vL)[hdL], (swU * vU)[hdU]])
        cinf = np.abs(comp - mu).max(initial=0)
        nzm = hxL.sum() + hxU.sum() + hdL.sum() + hdU.sum()
        z1 = zL.sum() + zU.sum() + vL.sum() + vU.sum()
        y1 = np.abs(yc).sum() + np.abs(yd).sum()
        sd = max(100.0, (y1 + z1) / max(1, mE + mI + nzm)) / 100.0
        sc = max(100.0, z1 / max(1, nzm)) / 100.0
        return max(dinf / sd, pinf, cinf / sc), dinf, pinf, cinf

    fval = evalf(x)
    theta = np.abs(c).sum() + np.abs(d - w).sum()
    theta_max, theta_min = 1e4 * max(1, theta), 1e-4 * max(1, theta)
    filt = []
    delta_last = 0.0
    status = 'Maximum_Iterations_Exceeded'
    it = 0
    n_reg = 0
    n_acc = 0
    log = []
    tau = max(0.99, 1 - mu)
    while True:
        E0, dinf, pinf, cinf0 = errors(gf, Jc, Jd, c, d, x, w, yc, yd, zL, zU, vL, vU, 0.0)
        if verbose:
            print('%4d f=%.10e th=%.2e dinf=%.2e mu=%.1e E0=%.2e' % (it, fval / sf, theta, dinf, mu, E0))
        log.append((it, fval / sf, theta, dinf, mu))
        if E0 <= tol:
            status = 'Solve_Succeeded'
            break
        # IPOPT's second termination test (defaults: acceptable_tol 1e-6, acceptable_iter 15, acceptable_constr_viol_tol 1e-2,
        # acceptable_compl_inf_tol 1e-2, acceptable_dual_inf_tol 1e10): CasADi reports it as success
        acc_level = E0 <= 1e-6 and dinf <= 1e10 and pinf <= 1e-2 and cinf0 <= 1e-2
        n_acc = n_acc + 1 if acc_level else 0
        if n_acc >= 15:
            status = 'Solved_To_Acceptable_Level'
            break
        if it >= max_iter:
            break
        while True:
            Emu = errors(gf, Jc, Jd, c, d, x, w, yc, yd, zL, zU, vL, vU, mu)[0]
            if Emu <= 10 * mu and mu > tol / 10 * (1 + 1e-12):
                mu = max(tol / 10, min(0.2 * mu, mu ** 1.5))
                tau = max(0.99, 1 - mu)
                filt = []
            else:
                break
        # ---- primal-dual system
        sxL, sxU, swL, swU = slacks(x, w)
        SigX = np.where(hxL, zL / sxL, 0) + np.where(hxU, zU / sxU, 0)
        SigW = np.where(hdL, vL / swL, 0) + np.where(hdU, vU / swU, 0)
        gphi_x = gf - np.where(hxL, mu / sxL, 0) + np.where(hxU, mu / sxU, 0) + kd * mu * (oneL_x * 1.0 - oneU_x * 1.0)
        gphi_w = -np.where(hdL, mu / swL, 0) + np.where(hdU, mu / swU, 0) + kd * mu * (oneL_w * 1.0 - oneU_w * 1.0)
        rx = gphi_x + Jc.T @ yc + Jd.T @ yd
        rw = gphi_w - yd
        W = hessL(x, yc, yd)
        rhs = -np.concatenate([rx, rw, c, d - w])
        delta = 0.0
        ntry = 0
        while True:
            K = sp.bmat([[W + sp.diags(SigX + delta), None, Jc.T, Jd.T],
                         [None, sp.diags(SigW + delta), None, -sp.identity(mI)],
                         [Jc, None, None, None],
                         [Jd, -sp.identity(mI), None, None]], format='csc')
            ok = True
            try:
                lu = spla.splu(K)
                sol = lu.solve(rhs)
                res = rhs - K @ sol                      # one step of iterative refinement
                sol += lu.solve(res)
            except RuntimeError:
                ok = False
            if ok and np.all(np.isfinite(sol)):
                dx, dw = sol[:n], sol[n:n + mI]
                curv = dx @ (W @ dx) + (SigX + delta) @ (dx * dx) + (SigW + delta) @ (dw * dw)
                if curv > 1e-14 * (dx @ dx + dw @ dw):
                    break
            ntry += 1
            n_reg += 1
            if delta == 0.0:
                delta = 1e-4 if delta_last == 0 else max(1e-20, delta_last / 3)
            else:
                delta *= 100 if delta_last == 0 else 8
            if delta > 1e40:
                status = 'Error_In_Step_Computation'
                break
        if status == 'Error_In_Step_Computation':
            break
        if delta > 0:
            delta_last = delta
        dyc, dyd = sol[n + mI:n + mI + mE], sol[n + mI + mE:]
        dzL = np.where(hxL, mu / sxL - zL - zL / sxL * dx, 0)
        dzU = np.where(hxU, mu / sxU - zU + zU / sxU * dx, 0)
        dvL = np.where(hdL, mu / swL - vL - vL / swL * dw, 0)
        dvU = np.where(hdU, mu / swU - vU + vU / swU * dw, 0)

        def ftb(v, dv, mask, sign):
            sel = mask & (sign * dv < 0)
            return np.min(-tau * v[sel] / (sign * dv[sel]), initial=1.0)

        a_max = min(ftb(sxL, dx, hxL, 1), ftb(sxU, dx, hxU, -1), ftb(swL, dw, hdL, 1), ftb(swU, dw, hdU, -1), 1.0)
        a_z = min(ftb(zL, dzL, hxL, 1), ftb(zU, dzU, hxU, 1), ftb(vL, dvL, hdL, 1), ftb(vU, dvU, hdU, 1), 1.0)
        gphid = gphi_x @ dx + gphi_w @ dw
        phi = barrier(x, w, mu, fval)
        # ---- filter line search
        if gphid < 0:
            a_min = 0.05 * min(1e-5, 1e-8 * theta / (-gphid) if theta > 0 else np.inf,
                               theta ** 1.1 / (-gphid) ** 2.3 if theta <= theta_min else np.inf)
        else:
            a_min = 0.05 * 1e-5
        alpha = a_max
        accepted = False
        nls = 0
        soc_used = False
        while alpha >= a_min or nls == 0:
            xt, wt = x + alpha * dx, w + alpha * dw
            ct, dt = evalg(xt)
            ft = evalf(xt)
            tht = np.abs(ct).sum() + np.abs(dt - wt).sum()
            pht = barrier(xt, wt, mu, ft)

            def acceptable(tht, pht):
                if not (np.isfinite(pht) and np.isfinite(tht)) or tht >= theta_max:
                    return False, False
                for (tf, pf) in filt:
                    if tht >= tf and pht >= pf:
                        return False, False
                switching = gphid < 0 and theta <= theta_min and alpha * (-gphid) ** 2.3 > theta ** 1.1
                if switching:
                    return pht - phi - 10 * EPS * abs(phi) <= 1e-8 * alpha * gphid, True
                return (tht <= (1 - 1e-5) * theta or pht - phi - 10 * EPS * abs(phi) <= -1e-8 * theta), False

            ok, armijo = acceptable(tht, pht)
            if ok:
                accepted = True
                break
            # second-order correction on the first trial (max_soc = 4, kappa_soc = 0.99)
            if soc and nls == 0 and tht >= theta:
                c_soc, dmw_soc = alpha * c + ct, alpha * (d - w) + (dt - wt)
                th_old = tht
                for _ in range(4):
                    rhs2 = -np.concatenate([rx, rw, c_soc, dmw_soc])
                    s2 = lu.solve(rhs2)
                    dx2, dw2 = s2[:n], s2[n:n + mI]
                    a2 = min(ftb(sxL, dx2, hxL, 1), ftb(sxU, dx2, hxU, -1), ftb(swL, dw2, hdL, 1), ftb(swU, dw2, hdU, -1), 1.0)
                    xt2, wt2 = x + a2 * dx2, w + a2 * dw2
                    ct2, dt2 = evalg(xt2)
                    ft2 = evalf(xt2)
                    th2 = np.abs(ct2).sum() + np.abs(dt2 - wt2).sum()
                    ph2 = barrier(xt2, wt2, mu, ft2)
                    ok2, arm2 = acceptable(th2, ph2)
                    if ok2:
                        accepted, soc_used = True, True
                        xt, wt, ct, dt, ft, tht, pht, armijo = xt2, wt2, ct2, dt2, ft2, th2, ph2, arm2
                        sol = s2
                        dyc, dyd = sol[n + mI:n + mI + mE], sol[n + mI + mE:]
                        alpha_y = a2
                        break
                    if th2 > 0.99 * th_old:
                        break
                    th_old = th2
                    c_soc, dmw_soc = a2 * c_soc + ct2, a2 * dmw_soc + (dt2 - wt2)
                if accepted:
                    break
            alpha *= 0.5
            nls += 1
        if not accepted:
            # no restoration phase here; like IPOPT, a failure at an acceptable point ends as "acceptable"
            status = 'Solved_To_Acceptable_Level' if acc_level else 'Restoration_Failed'
            break
        if not armijo:
            filt.append(((1 - 1e-5) * theta, phi - 1e-8 * theta))
        ay = alpha_y if soc_used else alpha
        x, w = xt, wt
        yc = yc + ay * dyc
        yd = yd + ay * dyd
        zL = zL + a_z * dzL; zU = zU + a_z * dzU; vL = vL + a_z * dvL; vU = vU + a_z * dvU
        sxL, sxU, swL, swU = slacks(x, w)
        ks = 1e10                                             # kappa_sigma safeguard, eq. (16)
        zL = np.where(hxL, np.clip(zL, mu / (ks * sxL), ks * mu / sxL), 0)
        zU = np.where(hxU, np.clip(zU, mu / (ks * sxU), ks * mu / sxU), 0)
        vL = np.where(hdL, np.clip(vL, mu / (ks * swL), ks * mu / swL), 0)
        vU = np.where(hdU, np.clip(vU, mu / (ks * swU), ks * mu / swU), 0)
        c, d, fval, theta = ct, dt, ft, tht
        gf, Jc, Jd = derivs(x)
        it += 1

    xfull[free] = x
    lam = np.zeros(m_all)
    lam[eqr] = yc * sg[eqr] / sf
    lam[inr] = yd * sg[inr] / sf
    zLf = np.zeros(n_all); zUf = np.zeros(n_all)
    zLf[free] = zL / sf; zUf[free] = zU / sf
    return Result(x=xfull.copy(), f=fval / sf, status=status, success=status in ('Solve_Succeeded', 'Solved_To_Acceptable_Level'), iters=it,
                  kkt=E0, dinf=dinf, pinf=pinf, mu=mu, lam=lam, zL=zLf, zU=zUf, n_reg=n_reg,
                  time=time.perf_counter() - t_start, log=log, slack=w / sg[inr])
