// Batched primal-dual interior-point iteration for the multiple-shooting train OCP.
//
// Replaces the NLP solve of the reference (mseetc/ocp.py:290 ca.nlpsol('ipopt'), :359 self.solver(...)):
// same variables, rows and bounds as ocp.py:166-272, same objective (ocp.py:146-154,223,243-245,276-284),
// IPOPT's published filter line-search algorithm with its default constants -- but the KKT system is solved
// by a Riccati recursion over the shooting intervals (state (t, b, Fel_{k-1}), controls (Fel, Fpb, s))
// instead of a sparse LDL^T, and all instances of a batch advance in lock step:
//
//   cell_trial   thread = (interval k, instance)   trial point  x+alpha*dx, constraint violation, barrier
//   inst_decide  thread = instance                 filter acceptance test (sequential, deterministic sums)
//   cell_eval    thread = (interval k, instance)   RK4 step + sensitivities, Lagrangian Hessian, condensed
//                                                  stage QP, KKT-error partials
//   inst_step    thread = instance                 convergence / barrier update, Riccati backward + forward,
//                                                  fraction-to-boundary, new step
//
// Every function here is host/device so that tests/hostsim can run the identical arithmetic on the CPU
// (test harness only -- the product has no CPU path).
#pragma once
#include "layout.cuh"
#include "model.cuh"

namespace mseetc {

#define MS_KAPPA_D 1e-5
#define MS_KAPPA_SIGMA 1e10
#define MS_VEL0SQ ((60.0 / 3.6) * (60.0 / 3.6))   // ocp.py:325

MS_HD void finish(const Ctx& c, int s, int status) {
    c.I(SI_STATUS, s) = status;
    c.I(SI_PHASE, s) = PH_DONE;
#if defined(__CUDA_ARCH__)
    atomicAdd(c.done, 1);
#else
    *c.done += 1;
#endif
}

MS_HD void count_cells(const Ctx& c, int which, int n) {
#if defined(__CUDA_ARCH__)
    atomicAdd(c.cnt + which, (unsigned long long)n);
#else
    c.cnt[which] += (unsigned long long)n;
#endif
}

MS_HD double relaxL(double L) { return L - 1e-8 * fmax(1.0, fabs(L)); }   // IPOPT bound_relax_factor
MS_HD double relaxU(double U) { return U + 1e-8 * fmax(1.0, fabs(U)); }

MS_HD double push2(double v, double L, double U) {   // IPOPT bound_push = bound_frac = 1e-2
    double pL = fmin(1e-2 * fmax(1.0, fabs(L)), 1e-2 * (U - L));
    double pU = fmin(1e-2 * fmax(1.0, fabs(U)), 1e-2 * (U - L));
    return fmin(fmax(v, L + pL), U - pU);
}
MS_HD double push1(double v, double L) { return fmax(v, L + 1e-2 * fmax(1.0, fabs(L))); }

struct Bnd {
    double felL, felU, fpbL, fpbU, slL, tL, tU, bL, bU, pL, pU, aL, aU, lL;
};

MS_HD Bnd load_bounds(const Ctx& c, int k, int s) {
    Bnd b;
    b.felL = relaxL(c.P(P_FEL_L, s));
    b.felU = relaxU(c.P(P_FEL_U, s));
    b.fpbL = relaxL(c.P(P_FPB_L, s));
    b.fpbU = relaxU(0.0);
    b.slL = relaxL(0.0);
    b.tL = relaxL(c.P(P_T0, s));
    b.tU = relaxU(c.P(P_T, s));
    b.bL = relaxL(c.P(P_BMIN, s));
    b.bU = relaxU(c.W(WS_TRK + TRK_BMAX, k, s));
    b.pL = relaxL(c.P(P_P_LO, s));
    b.pU = relaxU(c.P(P_P_UP, s));
    b.aL = relaxL(c.P(P_A_LO, s));
    b.aU = relaxU(c.P(P_A_UP, s));
    b.lL = relaxL(0.0);
    return b;
}

MS_HD IntervalCoef load_coef(const Ctx& c, int k, int s) {
    IntervalCoef q;
    q.ds = c.W(WS_TRK + TRK_DS, k, s);
    q.c0 = c.W(WS_TRK + TRK_C0, k, s);
    q.sr0 = c.P(P_SR0, s);
    q.sr1 = c.P(P_SR1, s);
    q.sr2 = c.P(P_SR2, s);
    return q;
}

// initial guess of the reference (ocp.py:325-339) pushed into the relaxed bounds
MS_HD double init_b(const Ctx& c, int j, int s, int N) {
    if (j == 0) return c.P(P_B0, s);
    if (j == N) return c.P(P_BN, s);
    return push2(MS_VEL0SQ, relaxL(c.P(P_BMIN, s)), relaxU(c.W(WS_TRK + TRK_BMAX, j, s)));
}
MS_HD double init_t(const Ctx& c, int j, int s, int N) {
    double t0 = c.P(P_T0, s), T = c.P(P_T, s);
    if (j == 0) return t0;
    return push2(t0 + j * ((T - t0) / N), relaxL(t0), relaxU(T));
}

// values of the inequality rows at a point                                  (ocp.py:189,199,225-226)
MS_HD void ineq_values(const Ctx& c, int s, double fel, double fpb, double sl, double b0, double b1,
                       const IntervalCoef& q, double* d) {
    d[R_P0] = fel * sqrt(b0);
    d[R_P1] = fel * sqrt(b1);
    d[R_ACC] = accel(b0, fel + fpb, q);
    d[R_LTR] = sl - c.P(P_CT, s) * fel;
    d[R_LRG] = sl + c.P(P_CR, s) * fel;
}

MS_HD bool row_on(const Config& g, int j) {
    if (j == R_P0 || j == R_P1) return g.withPower != 0;
    if (j == R_LTR || j == R_LRG) return g.energy != 0;
    return true;
}
MS_HD void row_bounds(const Bnd& B, int j, double& L, double& U, bool& hasU) {
    hasU = true;
    if (j == R_P0 || j == R_P1) { L = B.pL; U = B.pU; }
    else if (j == R_ACC) { L = B.aL; U = B.aU; }
    else { L = B.lL; U = 0.0; hasU = false; }
}

// ------------------------------------------------------------------------------------------------
// initialisation: x0 pushed inside the bounds, slacks from d(x0), multipliers 1 / 0   (IPOPT sec. 3.6)
// ------------------------------------------------------------------------------------------------
MS_HD void cell_init(const Ctx& c, int k, int s) {
    const Config& g = c.cfg;
    const int N = c.I(SI_N_INT, s);
    if (s >= g.nInst || k > N) return;
    const int it = WS_IT0;
    for (int f = 0; f < IT_N; ++f) { c.W(WS_IT0 + f, k, s) = 0.0; c.W(WS_IT1 + f, k, s) = 0.0; }
    for (int f = 0; f < ST_N; ++f) c.W(WS_ST + f, k, s) = 0.0;
    Bnd B = load_bounds(c, k, s);
    double t = init_t(c, k, s, N), b = init_b(c, k, s, N);
    c.W(it + IT_T, k, s) = t;
    c.W(it + IT_B, k, s) = b;
    if (k >= 1 && k <= N) { c.W(it + IT_Z + Z_T_L, k, s) = 1.0; c.W(it + IT_Z + Z_T_U, k, s) = 1.0; }
    if (k >= 1 && k < N) { c.W(it + IT_Z + Z_B_L, k, s) = 1.0; c.W(it + IT_Z + Z_B_U, k, s) = 1.0; }
    if (k == N) return;
    double fel = push2(0.5, B.felL, B.felU);
    double fpb = g.withPn ? push2(-0.1, B.fpbL, B.fpbU) : 0.0;
    double sl = push1(1.0, B.slL);
    c.W(it + IT_FEL, k, s) = fel;
    c.W(it + IT_FPB, k, s) = fpb;
    c.W(it + IT_SL, k, s) = sl;
    c.W(it + IT_Z + Z_FEL_L, k, s) = 1.0;
    c.W(it + IT_Z + Z_FEL_U, k, s) = 1.0;
    if (g.withPn) { c.W(it + IT_Z + Z_FPB_L, k, s) = 1.0; c.W(it + IT_Z + Z_FPB_U, k, s) = 1.0; }
    c.W(it + IT_Z + Z_SL_L, k, s) = 1.0;
    IntervalCoef q = load_coef(c, k, s);
    double d[NROW];
    ineq_values(c, s, fel, fpb, sl, b, init_b(c, k + 1, s, N), q, d);
    for (int j = 0; j < NROW; ++j) {
        if (!row_on(g, j)) continue;
        double L, U; bool hasU;
        row_bounds(B, j, L, U, hasU);
        c.W(it + IT_W + j, k, s) = hasU ? push2(d[j], L, U) : push1(d[j], L);
        const int zl = (j == R_P0) ? Z_P0_L : (j == R_P1) ? Z_P1_L : (j == R_ACC) ? Z_ACC_L : (j == R_LTR) ? Z_LTR_L : Z_LRG_L;
        c.W(it + IT_Z + zl, k, s) = 1.0;
        if (hasU) c.W(it + IT_Z + zl + 1, k, s) = 1.0;
    }
}

MS_HD void inst_init(const Ctx& c, int s) {
    const Config& g = c.cfg;
    if (s >= g.nInst) return;
    for (int f = 0; f < SD_N; ++f) c.D(f, s) = 0.0;
    c.D(SD_MU, s) = g.muInit;
    c.D(SD_TAU, s) = fmax(0.99, 1.0 - g.muInit);
    c.D(SD_THETA_MAX, s) = -1.0;   // set from theta(x0) in the first inst_step
    c.I(SI_PHASE, s) = PH_EVAL;
    c.I(SI_PARITY, s) = 0;
    c.I(SI_ITERS, s) = 0;
    c.I(SI_STATUS, s) = ST_RUNNING;
    c.I(SI_NLS, s) = 0;
    c.I(SI_NFILT, s) = 0;
    c.I(SI_NREG, s) = 0;
    c.I(SI_TICKS, s) = 0;
}

// slack of a bound and the matching barrier pieces
struct BarAcc {
    double slog, sdamp;
    bool ok;
};
MS_HD void bar_add(BarAcc& a, double slack, bool oneSided) {
    if (!(slack > 0.0)) a.ok = false;
    a.slog += log(slack);
    if (oneSided) a.sdamp += slack;
}

// ------------------------------------------------------------------------------------------------
// trial point                                                                 (IPOPT sec. 2.3, Alg. A step A-5)
// ------------------------------------------------------------------------------------------------
MS_HD void cell_trial(const Ctx& c, int k, int s) {
    const Config& g = c.cfg;
    if (s >= g.nInst || c.I(SI_PHASE, s) != PH_TRIAL) return;
    const int N = c.I(SI_N_INT, s);
    if (k > N) return;
    const int cur = c.I(SI_PARITY, s) ? WS_IT1 : WS_IT0;
    const int alt = c.I(SI_PARITY, s) ? WS_IT0 : WS_IT1;
    const double al = c.D(SD_ALPHA, s), az = c.D(SD_ALPHA_Z, s), mu = c.D(SD_MU, s);
    const double scale = c.P(P_SCALE, s);
    Bnd B = load_bounds(c, k, s);
    BarAcc bar{0.0, 0.0, true};
    double th = 0.0, fo = 0.0;

    // multiplier update of one bound: z + az*dz, dz = mu/s - z -+ (z/s) dv, then the kappa_sigma safeguard (eq. 16)
    auto zstep = [&](int zi, double sOld, double sNew, double dvSigned) {
        double z = c.W(cur + IT_Z + zi, k, s);
        double dz = mu / sOld - z - (z / sOld) * dvSigned;
        double zn = z + az * dz;
        zn = fmax(fmin(zn, MS_KAPPA_SIGMA * mu / sNew), mu / (MS_KAPPA_SIGMA * sNew));
        c.W(alt + IT_Z + zi, k, s) = zn;
    };
    auto var2 = [&](int itf, int stf, int zl, double L, double U, bool hasU, bool oneSided) -> double {
        double v = c.W(cur + itf, k, s), dv = c.W(WS_ST + stf, k, s);
        double vn = v + al * dv;
        c.W(alt + itf, k, s) = vn;
        zstep(zl, v - L, vn - L, dv);
        bar_add(bar, vn - L, oneSided);
        if (hasU) { zstep(zl + 1, U - v, U - vn, -dv); bar_add(bar, U - vn, false); }
        return vn;
    };

    double t, b;
    if (k == 0) {
        t = c.W(cur + IT_T, k, s); b = c.W(cur + IT_B, k, s);
        c.W(alt + IT_T, k, s) = t; c.W(alt + IT_B, k, s) = b;
    } else {
        t = var2(IT_T, ST_T, Z_T_L, B.tL, B.tU, true, false);
        if (k < N) b = var2(IT_B, ST_B, Z_B_L, B.bL, B.bU, true, false);
        else { b = c.W(cur + IT_B, k, s); c.W(alt + IT_B, k, s) = b; }
    }
    if (k == N) {
        if (!g.energy) fo += t / scale;
        c.W(WS_PART + PT_TH, k, s) = 0.0;
        c.W(WS_PART + PT_F, k, s) = fo;
        c.W(WS_PART + PT_SLOG, k, s) = bar.ok ? bar.slog : NAN;
        c.W(WS_PART + PT_SDAMP, k, s) = 0.0;
        return;
    }
    double fel = var2(IT_FEL, ST_FEL, Z_FEL_L, B.felL, B.felU, true, false);
    double fpb = 0.0;
    if (g.withPn) fpb = var2(IT_FPB, ST_FPB, Z_FPB_L, B.fpbL, B.fpbU, true, false);
    else c.W(alt + IT_FPB, k, s) = 0.0;
    double sl = var2(IT_SL, ST_SL, Z_SL_L, B.slL, 0.0, false, true);
    // neighbours at the trial point
    double t1 = c.W(cur + IT_T, k + 1, s) + al * c.W(WS_ST + ST_T, k + 1, s);
    double b1 = (k + 1 < N) ? c.W(cur + IT_B, k + 1, s) + al * c.W(WS_ST + ST_B, k + 1, s) : c.W(cur + IT_B, k + 1, s);
    IntervalCoef q = load_coef(c, k, s);
    double tau, phib;
    shoot<double>(b, fel + fpb, q, g.numSteps, g.numApprox, tau, phib);
    double ct = t1 - t - tau, cb = b1 - phib;
    th += fabs(ct) + fabs(cb);
    c.W(alt + IT_YT, k, s) = c.W(cur + IT_YT, k, s) + al * c.W(WS_ST + ST_YT, k, s);
    c.W(alt + IT_YB, k, s) = c.W(cur + IT_YB, k, s) + al * c.W(WS_ST + ST_YB, k, s);
    double d[NROW];
    ineq_values(c, s, fel, fpb, sl, b, b1, q, d);
    for (int j = 0; j < NROW; ++j) {
        if (!row_on(g, j)) continue;
        double L, U; bool hasU;
        row_bounds(B, j, L, U, hasU);
        const int zl = (j == R_P0) ? Z_P0_L : (j == R_P1) ? Z_P1_L : (j == R_ACC) ? Z_ACC_L : (j == R_LTR) ? Z_LTR_L : Z_LRG_L;
        double w = c.W(cur + IT_W + j, k, s), dw = c.W(WS_ST + ST_W + j, k, s);
        double wn = w + al * dw;
        c.W(alt + IT_W + j, k, s) = wn;
        zstep(zl, w - L, wn - L, dw);
        bar_add(bar, wn - L, !hasU);
        if (hasU) { zstep(zl + 1, U - w, U - wn, -dw); bar_add(bar, U - wn, false); }
        c.W(alt + IT_YD + j, k, s) = c.W(cur + IT_YD + j, k, s) + al * c.W(WS_ST + ST_YD + j, k, s);
        th += fabs(d[j] - wn);
    }
    if (g.energy) {                                                            // ocp.py:223,243-245
        fo += q.ds * (fel + sl) / scale;
        if (k >= 1) {
            double fprev = c.W(cur + IT_FEL, k - 1, s) + al * c.W(WS_ST + ST_FEL, k - 1, s);
            fo += 1e-3 * (fel - fprev) * (fel - fprev) / scale;
        }
    } else {                                                                   // ocp.py:146-150
        fo += 1e-4 * (fel * fel + fpb * fpb) / scale;
    }
    c.W(WS_PART + PT_TH, k, s) = th;
    c.W(WS_PART + PT_F, k, s) = fo;
    c.W(WS_PART + PT_SLOG, k, s) = bar.ok ? bar.slog : NAN;
    c.W(WS_PART + PT_SDAMP, k, s) = bar.sdamp;
}

// ------------------------------------------------------------------------------------------------
// filter acceptance                                                         (IPOPT sec. 2.3, eqs. 18-20)
// ------------------------------------------------------------------------------------------------
MS_HD void inst_decide(const Ctx& c, int s) {
    const Config& g = c.cfg;
    if (s >= g.nInst || c.I(SI_PHASE, s) != PH_TRIAL) return;
    const int N = c.I(SI_N_INT, s);
    double tht = 0.0, ft = 0.0, slog = 0.0, sdamp = 0.0;
    for (int k = 0; k <= N; ++k) {
        tht += c.W(WS_PART + PT_TH, k, s);
        ft += c.W(WS_PART + PT_F, k, s);
        slog += c.W(WS_PART + PT_SLOG, k, s);
        sdamp += c.W(WS_PART + PT_SDAMP, k, s);
    }
    const double mu = c.D(SD_MU, s);
    const double pht = ft - mu * slog + MS_KAPPA_D * mu * sdamp;
    const double theta = c.D(SD_THETA, s);
    const double phi = c.D(SD_FOBJ, s) - mu * c.D(SD_SLOG, s) + MS_KAPPA_D * mu * c.D(SD_SDAMP, s);
    const double gphid = c.D(SD_GPHID, s), alpha = c.D(SD_ALPHA, s);
    const double thmin = c.D(SD_THETA_MIN, s), thmax = c.D(SD_THETA_MAX, s);
    bool ok = isfinite(pht) && isfinite(tht) && tht < thmax;
    const int nf = c.I(SI_NFILT, s);
    for (int i = 0; ok && i < nf; ++i)
        if (tht >= c.D(SD_FILTER + 2 * i, s) && pht >= c.D(SD_FILTER + 2 * i + 1, s)) ok = false;
    bool armijo = false;
    if (ok) {
        const double eps10 = 10.0 * 2.220446049250313e-16 * fabs(phi);
        bool switching = gphid < 0.0 && theta <= thmin && alpha * pow(-gphid, 2.3) > pow(theta, 1.1);
        if (switching) {
            ok = (pht - phi - eps10 <= 1e-8 * alpha * gphid);
            armijo = true;
        } else {
            ok = (tht <= (1.0 - 1e-5) * theta) || (pht - phi - eps10 <= -1e-8 * theta);
        }
    }
    c.I(SI_TICKS, s) += 1;
    count_cells(c, 0, N + 1);
    if (ok) {
        if (!armijo) {                                    // augment the filter (eq. 22)
            int n = nf;
            if (n == 12) { for (int i = 0; i < 22; ++i) c.D(SD_FILTER + i, s) = c.D(SD_FILTER + i + 2, s); n = 11; }
            c.D(SD_FILTER + 2 * n, s) = (1.0 - 1e-5) * theta;
            c.D(SD_FILTER + 2 * n + 1, s) = phi - 1e-8 * theta;
            c.I(SI_NFILT, s) = n + 1;
        }
        c.I(SI_PARITY, s) ^= 1;
        c.I(SI_ITERS, s) += 1;
        c.I(SI_PHASE, s) = PH_EVAL;
    } else {
        double a2 = 0.5 * alpha;
        c.I(SI_NLS, s) += 1;
        if (a2 < c.D(SD_ALPHA_MIN, s)) {
            finish(c, s, ST_RESTORATION_FAILED);   // no restoration phase: report like IPOPT would
        } else {
            c.D(SD_ALPHA, s) = a2;
        }
    }
}

// symmetric 7x7 in packed upper storage, index order (t,b,f,Fel,Fpb,sl,b+)
MS_HD int sidx(int i, int j) { return (i <= j) ? i * 7 - i * (i - 1) / 2 + (j - i) : j * 7 - j * (j - 1) / 2 + (i - j); }
enum { V_T = 0, V_B, V_F, V_FEL, V_FPB, V_SL, V_BN, NV7 };

// ------------------------------------------------------------------------------------------------
// interval evaluation: RK4 + sensitivities, Hessian of the Lagrangian, condensed stage QP, KKT partials
// ------------------------------------------------------------------------------------------------
MS_HD void cell_eval(const Ctx& c, int k, int s) {
    const Config& g = c.cfg;
    if (s >= g.nInst || c.I(SI_PHASE, s) != PH_EVAL) return;
    const int N = c.I(SI_N_INT, s);
    if (k > N) return;
    const int it = c.I(SI_PARITY, s) ? WS_IT1 : WS_IT0;
    const double scale = c.P(P_SCALE, s);
    Bnd B = load_bounds(c, k, s);
    double H[28], g0[NV7], g1[NV7];
    for (int i = 0; i < 28; ++i) H[i] = 0.0;
    for (int i = 0; i < NV7; ++i) { g0[i] = 0.0; g1[i] = 0.0; }
    BarAcc bar{0.0, 0.0, true};
    double cmin = 1e300, cmax = 0.0, zsum = 0.0;

    auto bound = [&](int vi, int zi, double slack, double sign, bool oneSided) {
        // sign = +1 lower bound (slack = v-L), -1 upper bound (slack = U-v)
        double z = c.W(it + IT_Z + zi, k, s);
        H[sidx(vi, vi)] += z / slack;
        g1[vi] += -sign / slack + (oneSided ? MS_KAPPA_D : 0.0);
        double pr = z * slack;
        cmin = fmin(cmin, pr); cmax = fmax(cmax, pr); zsum += z;
        bar_add(bar, slack, oneSided);
    };

    const double t = c.W(it + IT_T, k, s), b = c.W(it + IT_B, k, s);
    double own_t = 0.0, own_b = 0.0;
    if (k >= 1) {
        bound(V_T, Z_T_L, t - B.tL, 1.0, false);
        bound(V_T, Z_T_U, B.tU - t, -1.0, false);
        own_t = -c.W(it + IT_Z + Z_T_L, k, s) + c.W(it + IT_Z + Z_T_U, k, s);
        if (k < N) {
            bound(V_B, Z_B_L, b - B.bL, 1.0, false);
            bound(V_B, Z_B_U, B.bU - b, -1.0, false);
            own_b = -c.W(it + IT_Z + Z_B_L, k, s) + c.W(it + IT_Z + Z_B_U, k, s);
        }
    }
    if (k == N) {
        // terminal node: only t_N carries a barrier; stored in the same QP planes for the Riccati start
        double fo = 0.0;
        if (!g.energy) { own_t += 1.0 / scale; fo = t / scale; }
        c.W(WS_QP + QP_H_TT, k, s) = H[sidx(V_T, V_T)];
        c.W(WS_QP + QP_G1_T, k, s) = g1[V_T];
        c.W(WS_PART + PC_TH, k, s) = 0.0;
        c.W(WS_PART + PC_F, k, s) = fo;
        c.W(WS_PART + PC_SLOG, k, s) = bar.ok ? bar.slog : NAN;
        c.W(WS_PART + PC_SDAMP, k, s) = 0.0;
        c.W(WS_PART + PC_DINF, k, s) = 0.0;
        c.W(WS_PART + PC_PINF, k, s) = 0.0;
        c.W(WS_PART + PC_CMIN, k, s) = cmin;
        c.W(WS_PART + PC_CMAX, k, s) = cmax;
        c.W(WS_PART + PC_ZSUM, k, s) = zsum;
        c.W(WS_PART + PC_YSUM, k, s) = 0.0;
        c.W(WS_PART + PC_OWN_B, k, s) = 0.0;
        c.W(WS_PART + PC_CN_B, k, s) = 0.0;
        c.W(WS_PART + PC_OWN_T, k, s) = own_t;
        return;
    }
    const double fel = c.W(it + IT_FEL, k, s), fpb = c.W(it + IT_FPB, k, s), sl = c.W(it + IT_SL, k, s);
    const double t1 = c.W(it + IT_T, k + 1, s), b1 = c.W(it + IT_B, k + 1, s);
    const double yt = c.W(it + IT_YT, k, s), yb = c.W(it + IT_YB, k, s);
    bound(V_FEL, Z_FEL_L, fel - B.felL, 1.0, false);
    bound(V_FEL, Z_FEL_U, B.felU - fel, -1.0, false);
    if (g.withPn) {
        bound(V_FPB, Z_FPB_L, fpb - B.fpbL, 1.0, false);
        bound(V_FPB, Z_FPB_U, B.fpbU - fpb, -1.0, false);
    } else {
        H[sidx(V_FPB, V_FPB)] = 1.0;        // dummy control, keeps the stage Hessian regular
    }
    bound(V_SL, Z_SL_L, sl - B.slL, 1.0, true);

    // ---- shooting: values + exact first/second sensitivities w.r.t. (b_k, F)
    IntervalCoef q = load_coef(c, k, s);
    Jet2 tau, phi;
    shoot<Jet2>(jvar0(b), jvar1(fel + fpb), q, g.numSteps, g.numApprox, tau, phi);
    const double ct = t1 - t - tau.v, cb = b1 - phi.v;
    const double v0 = sqrt(b), v1 = sqrt(b1);
    const double a_b = -(0.5 * q.sr1 / v0 + q.sr2), a_bb = 0.25 * q.sr1 / (b * v0);

    // ---- Hessian of the Lagrangian: coupling rows (ct = t1 - t - tau, cb = b1 - phi)
    {
        double hbb = -yt * tau.h00 - yb * phi.h00, hbF = -yt * tau.h01 - yb * phi.h01, hFF = -yt * tau.h11 - yb * phi.h11;
        H[sidx(V_B, V_B)] += hbb;
        H[sidx(V_B, V_FEL)] += hbF;
        H[sidx(V_FEL, V_FEL)] += hFF;
        if (g.withPn) { H[sidx(V_B, V_FPB)] += hbF; H[sidx(V_FEL, V_FPB)] += hFF; H[sidx(V_FPB, V_FPB)] += hFF; }
    }
    // ---- inequality rows: value, gradient over the 7 local variables, Lagrangian-Hessian contribution
    double d[NROW], J[NROW][NV7];
    for (int j = 0; j < NROW; ++j) for (int i = 0; i < NV7; ++i) J[j][i] = 0.0;
    ineq_values(c, s, fel, fpb, sl, b, b1, q, d);
    double ydv[NROW];
    for (int j = 0; j < NROW; ++j) ydv[j] = c.W(it + IT_YD + j, k, s);
    J[R_P0][V_B] = 0.5 * fel / v0; J[R_P0][V_FEL] = v0;
    J[R_P1][V_BN] = 0.5 * fel / v1; J[R_P1][V_FEL] = v1;
    J[R_ACC][V_B] = a_b; J[R_ACC][V_FEL] = 1.0; J[R_ACC][V_FPB] = g.withPn ? 1.0 : 0.0;
    J[R_LTR][V_FEL] = -c.P(P_CT, s); J[R_LTR][V_SL] = 1.0;
    J[R_LRG][V_FEL] = c.P(P_CR, s); J[R_LRG][V_SL] = 1.0;
    if (g.withPower) {
        H[sidx(V_B, V_B)] += ydv[R_P0] * (-0.25 * fel / (b * v0));
        H[sidx(V_B, V_FEL)] += ydv[R_P0] * (0.5 / v0);
        H[sidx(V_BN, V_BN)] += ydv[R_P1] * (-0.25 * fel / (b1 * v1));
        H[sidx(V_FEL, V_BN)] += ydv[R_P1] * (0.5 / v1);
    }
    H[sidx(V_B, V_B)] += ydv[R_ACC] * a_bb;

    // ---- objective                                                           (ocp.py:146-154,223,243-245)
    double fo = 0.0, gf_fel = 0.0, gf_fpb = 0.0, gf_sl = 0.0;
    if (g.energy) {
        const double w2 = 2e-3 / scale;
        fo = q.ds * (fel + sl) / scale;
        gf_fel = q.ds / scale; gf_sl = q.ds / scale;
        g0[V_FEL] += q.ds / scale; g0[V_SL] += q.ds / scale;
        if (k >= 1) {
            double df = fel - c.W(it + IT_FEL, k - 1, s);
            fo += 1e-3 * df * df / scale;
            H[sidx(V_FEL, V_FEL)] += w2; H[sidx(V_F, V_F)] += w2; H[sidx(V_F, V_FEL)] -= w2;
            g0[V_FEL] += w2 * df; g0[V_F] -= w2 * df;
            gf_fel += w2 * df;
        }
        if (k + 1 < N) gf_fel -= w2 * (c.W(it + IT_FEL, k + 1, s) - fel);
    } else {
        const double w4 = 2e-4 / scale;
        fo = 1e-4 * (fel * fel + fpb * fpb) / scale;
        H[sidx(V_FEL, V_FEL)] += w4; g0[V_FEL] += w4 * fel; gf_fel = w4 * fel;
        if (g.withPn) { H[sidx(V_FPB, V_FPB)] += w4; g0[V_FPB] += w4 * fpb; gf_fpb = w4 * fpb; }
    }

    // ---- condensation of the inequality rows (slack w, multiplier v_L/v_U)
    double th = fabs(ct) + fabs(cb), pinf = fmax(fabs(ct), fabs(cb));
    double ysum = fabs(yt) + fabs(yb), dinf = 0.0;
    double rx_fel = gf_fel - tau.g1 * yt - phi.g1 * yb - c.W(it + IT_Z + Z_FEL_L, k, s) + c.W(it + IT_Z + Z_FEL_U, k, s);
    double rx_fpb = gf_fpb - tau.g1 * yt - phi.g1 * yb - c.W(it + IT_Z + Z_FPB_L, k, s) + c.W(it + IT_Z + Z_FPB_U, k, s);
    double rx_sl = gf_sl - c.W(it + IT_Z + Z_SL_L, k, s);
    own_b += -tau.g0 * yt - phi.g0 * yb;
    own_t += -yt;
    double cn_b = yb;
    for (int j = 0; j < NROW; ++j) {
        c.W(WS_QP + QP_RES + j, k, s) = 0.0;
        if (!row_on(g, j)) continue;
        double L, U; bool hasU;
        row_bounds(B, j, L, U, hasU);
        const int zl = (j == R_P0) ? Z_P0_L : (j == R_P1) ? Z_P1_L : (j == R_ACC) ? Z_ACC_L : (j == R_LTR) ? Z_LTR_L : Z_LRG_L;
        const double w = c.W(it + IT_W + j, k, s);
        const double vL = c.W(it + IT_Z + zl, k, s), sL = w - L;
        double sig = vL / sL, coef = -1.0 / sL + (hasU ? 0.0 : MS_KAPPA_D);
        double rw = -ydv[j] - vL;
        double pr = vL * sL;
        cmin = fmin(cmin, pr); cmax = fmax(cmax, pr); zsum += vL;
        bar_add(bar, sL, !hasU);
        if (hasU) {
            const double vU = c.W(it + IT_Z + zl + 1, k, s), sU = U - w;
            sig += vU / sU; coef += 1.0 / sU; rw += vU;
            pr = vU * sU;
            cmin = fmin(cmin, pr); cmax = fmax(cmax, pr); zsum += vU;
            bar_add(bar, sU, false);
        }
        const double res = d[j] - w;
        c.W(WS_QP + QP_RES + j, k, s) = res;
        th += fabs(res); pinf = fmax(pinf, fabs(res)); ysum += fabs(ydv[j]);
        dinf = fmax(dinf, fabs(rw));
        for (int a = 0; a < NV7; ++a) {
            if (J[j][a] == 0.0) continue;
            g0[a] += sig * res * J[j][a];
            g1[a] += coef * J[j][a];
            for (int e = a; e < NV7; ++e) H[sidx(a, e)] += sig * J[j][a] * J[j][e];
        }
        rx_fel += ydv[j] * J[j][V_FEL];
        rx_fpb += ydv[j] * J[j][V_FPB];
        rx_sl += ydv[j] * J[j][V_SL];
        own_b += ydv[j] * J[j][V_B];
        cn_b += ydv[j] * J[j][V_BN];
    }
    dinf = fmax(dinf, fmax(fabs(rx_fel), fabs(rx_sl)));
    if (g.withPn) dinf = fmax(dinf, fabs(rx_fpb));

    // ---- fold the b_{k+1} column through the linearised coupling row (exact Newton step of the reference NLP)
    const double rt = -ct, rb = -cb;
    double av[6] = {0.0, phi.g0, 0.0, phi.g1, g.withPn ? phi.g1 : 0.0, 0.0};
    double hc[6];
    for (int i = 0; i < 6; ++i) hc[i] = H[sidx(i, V_BN)];
    const double hpp = H[sidx(V_BN, V_BN)], gp0 = g0[V_BN], gp1 = g1[V_BN];
    if (k + 1 < N) {
        for (int i = 0; i < 6; ++i) {
            for (int j = i; j < 6; ++j) H[sidx(i, j)] += hc[i] * av[j] + av[i] * hc[j] + hpp * av[i] * av[j];
            g0[i] += hc[i] * rb + (hpp * rb + gp0) * av[i];
            g1[i] += gp1 * av[i];
        }
    }
    // ---- store
    c.W(WS_QP + QP_H_TT, k, s) = H[sidx(V_T, V_T)];
    c.W(WS_QP + QP_H_BB, k, s) = H[sidx(V_B, V_B)];
    c.W(WS_QP + QP_H_BFEL, k, s) = H[sidx(V_B, V_FEL)];
    c.W(WS_QP + QP_H_BFPB, k, s) = H[sidx(V_B, V_FPB)];
    c.W(WS_QP + QP_H_BSL, k, s) = H[sidx(V_B, V_SL)];
    c.W(WS_QP + QP_H_FF, k, s) = H[sidx(V_F, V_F)];
    c.W(WS_QP + QP_H_FFEL, k, s) = H[sidx(V_F, V_FEL)];
    c.W(WS_QP + QP_H_FELFEL, k, s) = H[sidx(V_FEL, V_FEL)];
    c.W(WS_QP + QP_H_FELFPB, k, s) = H[sidx(V_FEL, V_FPB)];
    c.W(WS_QP + QP_H_FELSL, k, s) = H[sidx(V_FEL, V_SL)];
    c.W(WS_QP + QP_H_FPBFPB, k, s) = H[sidx(V_FPB, V_FPB)];
    c.W(WS_QP + QP_H_FPBSL, k, s) = H[sidx(V_FPB, V_SL)];
    c.W(WS_QP + QP_H_SLSL, k, s) = H[sidx(V_SL, V_SL)];
    c.W(WS_QP + QP_TAU_B, k, s) = tau.g0;
    c.W(WS_QP + QP_TAU_F, k, s) = tau.g1;
    c.W(WS_QP + QP_PHI_B, k, s) = phi.g0;
    c.W(WS_QP + QP_PHI_F, k, s) = phi.g1;
    c.W(WS_QP + QP_RT, k, s) = rt;
    c.W(WS_QP + QP_RB, k, s) = rb;
    c.W(WS_QP + QP_G0_B, k, s) = g0[V_B];
    c.W(WS_QP + QP_G0_F, k, s) = g0[V_F];
    c.W(WS_QP + QP_G0_FEL, k, s) = g0[V_FEL];
    c.W(WS_QP + QP_G0_FPB, k, s) = g0[V_FPB];
    c.W(WS_QP + QP_G0_SL, k, s) = g0[V_SL];
    c.W(WS_QP + QP_G1_T, k, s) = g1[V_T];
    c.W(WS_QP + QP_G1_B, k, s) = g1[V_B];
    c.W(WS_QP + QP_G1_FEL, k, s) = g1[V_FEL];
    c.W(WS_QP + QP_G1_FPB, k, s) = g1[V_FPB];
    c.W(WS_QP + QP_G1_SL, k, s) = g1[V_SL];
    c.W(WS_QP + QP_HC_B, k, s) = hc[V_B];
    c.W(WS_QP + QP_HC_FEL, k, s) = hc[V_FEL];
    c.W(WS_QP + QP_HC_FPB, k, s) = hc[V_FPB];
    c.W(WS_QP + QP_HC_SL, k, s) = hc[V_SL];
    c.W(WS_QP + QP_HPP, k, s) = hpp;
    c.W(WS_QP + QP_GP0, k, s) = gp0;
    c.W(WS_QP + QP_GP1, k, s) = gp1;
    c.W(WS_QP + QP_J_P0_B, k, s) = J[R_P0][V_B];
    c.W(WS_QP + QP_J_P0_FEL, k, s) = J[R_P0][V_FEL];
    c.W(WS_QP + QP_J_P1_FEL, k, s) = J[R_P1][V_FEL];
    c.W(WS_QP + QP_J_P1_BN, k, s) = J[R_P1][V_BN];
    c.W(WS_QP + QP_J_ACC_B, k, s) = J[R_ACC][V_B];
    c.W(WS_QP + QP_J_LTR_FEL, k, s) = J[R_LTR][V_FEL];
    c.W(WS_QP + QP_J_LTR_B, k, s) = J[R_LTR][V_B];
    c.W(WS_QP + QP_J_LTR_BN, k, s) = J[R_LTR][V_BN];
    c.W(WS_QP + QP_J_LRG_FEL, k, s) = J[R_LRG][V_FEL];
    c.W(WS_QP + QP_J_LRG_B, k, s) = J[R_LRG][V_B];
    c.W(WS_QP + QP_J_LRG_BN, k, s) = J[R_LRG][V_BN];
    c.W(WS_PART + PC_TH, k, s) = th;
    c.W(WS_PART + PC_F, k, s) = fo;
    c.W(WS_PART + PC_SLOG, k, s) = bar.ok ? bar.slog : NAN;
    c.W(WS_PART + PC_SDAMP, k, s) = bar.sdamp;
    c.W(WS_PART + PC_DINF, k, s) = dinf;
    c.W(WS_PART + PC_PINF, k, s) = pinf;
    c.W(WS_PART + PC_CMIN, k, s) = cmin;
    c.W(WS_PART + PC_CMAX, k, s) = cmax;
    c.W(WS_PART + PC_ZSUM, k, s) = zsum;
    c.W(WS_PART + PC_YSUM, k, s) = ysum;
    c.W(WS_PART + PC_OWN_B, k, s) = own_b;
    c.W(WS_PART + PC_CN_B, k, s) = cn_b;
    c.W(WS_PART + PC_OWN_T, k, s) = own_t;
}

// ------------------------------------------------------------------------------------------------
// Riccati recursion
// ------------------------------------------------------------------------------------------------
struct Sym3 { double tt, tb, tf, bb, bf, ff; };

// dense symmetric 6x6 (t,b,f,Fel,Fpb,sl) in full storage; small enough to live in registers
MS_HD void load_stage(const Ctx& c, int k, int s, double mu, double delta, double M[6][6], double m[6]) {
    for (int i = 0; i < 6; ++i) { for (int j = 0; j < 6; ++j) M[i][j] = 0.0; }
    M[0][0] = c.W(WS_QP + QP_H_TT, k, s) + delta;
    M[1][1] = c.W(WS_QP + QP_H_BB, k, s) + delta;
    M[1][3] = M[3][1] = c.W(WS_QP + QP_H_BFEL, k, s);
    M[1][4] = M[4][1] = c.W(WS_QP + QP_H_BFPB, k, s);
    M[1][5] = M[5][1] = c.W(WS_QP + QP_H_BSL, k, s);
    M[2][2] = c.W(WS_QP + QP_H_FF, k, s);
    M[2][3] = M[3][2] = c.W(WS_QP + QP_H_FFEL, k, s);
    M[3][3] = c.W(WS_QP + QP_H_FELFEL, k, s) + delta;
    M[3][4] = M[4][3] = c.W(WS_QP + QP_H_FELFPB, k, s);
    M[3][5] = M[5][3] = c.W(WS_QP + QP_H_FELSL, k, s);
    M[4][4] = c.W(WS_QP + QP_H_FPBFPB, k, s) + delta;
    M[4][5] = M[5][4] = c.W(WS_QP + QP_H_FPBSL, k, s);
    M[5][5] = c.W(WS_QP + QP_H_SLSL, k, s) + delta;
    m[0] = mu * c.W(WS_QP + QP_G1_T, k, s);
    m[1] = c.W(WS_QP + QP_G0_B, k, s) + mu * c.W(WS_QP + QP_G1_B, k, s);
    m[2] = c.W(WS_QP + QP_G0_F, k, s);
    m[3] = c.W(WS_QP + QP_G0_FEL, k, s) + mu * c.W(WS_QP + QP_G1_FEL, k, s);
    m[4] = c.W(WS_QP + QP_G0_FPB, k, s) + mu * c.W(WS_QP + QP_G1_FPB, k, s);
    m[5] = c.W(WS_QP + QP_G0_SL, k, s) + mu * c.W(WS_QP + QP_G1_SL, k, s);
}

// backward sweep; returns false when a reduced control Hessian is not positive definite (wrong inertia)
MS_HD bool riccati_backward(const Ctx& c, int s, int N, double mu, double delta) {
    const Config& g = c.cfg;
    // terminal value function: only t_N is free (b_N fixed, f_N costless)
    double P[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, p[3] = {0, 0, 0};
    P[0][0] = c.W(WS_QP + QP_H_TT, N, s) + delta;
    p[0] = (g.energy ? 0.0 : 1.0 / c.P(P_SCALE, s)) + mu * c.W(WS_QP + QP_G1_T, N, s);
    for (int i = 0; i < 6; ++i) c.W(WS_RIC + RIC_P + i, N, s) = 0.0;
    c.W(WS_RIC + RIC_P + 0, N, s) = P[0][0];
    c.W(WS_RIC + RIC_PV + 0, N, s) = p[0];
    c.W(WS_RIC + RIC_PV + 1, N, s) = 0.0;
    c.W(WS_RIC + RIC_PV + 2, N, s) = 0.0;
    for (int k = N - 1; k >= 0; --k) {
        double M[6][6], m[6];
        load_stage(c, k, s, mu, delta, M, m);
        const double tb = c.W(WS_QP + QP_TAU_B, k, s), tF = c.W(WS_QP + QP_TAU_F, k, s);
        const double pb = c.W(WS_QP + QP_PHI_B, k, s), pF = c.W(WS_QP + QP_PHI_F, k, s);
        const double rt = c.W(WS_QP + QP_RT, k, s), rb = c.W(WS_QP + QP_RB, k, s);
        const double pn = g.withPn ? 1.0 : 0.0;
        // G = [A B]: rows t,b,f of the next state
        double G[3][6] = {{1.0, tb, 0.0, tF, pn * tF, 0.0}, {0.0, pb, 0.0, pF, pn * pF, 0.0}, {0.0, 0.0, 0.0, 1.0, 0.0, 0.0}};
        double r[3] = {rt, rb, 0.0};
        const bool last = (k == N - 1);
        if (last) { for (int j = 0; j < 6; ++j) G[1][j] = 0.0; r[1] = 0.0; }   // db_N = 0 handled by elimination
        // M += G' P G ; m += G' (P r + p)
        double Y[3][6], pr[3];
        for (int a = 0; a < 3; ++a) {
            pr[a] = p[a] + P[a][0] * r[0] + P[a][1] * r[1] + P[a][2] * r[2];
            for (int j = 0; j < 6; ++j) Y[a][j] = P[a][0] * G[0][j] + P[a][1] * G[1][j] + P[a][2] * G[2][j];
        }
        for (int i = 0; i < 6; ++i) {
            m[i] += G[0][i] * pr[0] + G[1][i] * pr[1] + G[2][i] * pr[2];
            for (int j = i; j < 6; ++j) {
                double v = M[i][j] + G[0][i] * Y[0][j] + G[1][i] * Y[1][j] + G[2][i] * Y[2][j];
                M[i][j] = v; M[j][i] = v;
            }
        }
        double eB = 0.0, ePn = 0.0, e0 = 0.0;
        if (last) {
            // terminal speed fixed: Phi_b db + Phi_F (dFel + dFpb) + rb = 0  ->  dFel = eB db + ePn dFpb + e0
            eB = -pb / pF; ePn = -pn; e0 = -rb / pF;
            // substitute: column/row Fel distributed onto b and Fpb, then Fel becomes a dummy control
            double colF[6];
            for (int i = 0; i < 6; ++i) colF[i] = M[i][3];
            const double mFF = M[3][3];
            for (int i = 0; i < 6; ++i) m[i] += colF[i] * e0;
            const double mF = m[3];
            double ev[6] = {0.0, eB, 0.0, 0.0, ePn, 0.0};
            for (int i = 0; i < 6; ++i) m[i] += ev[i] * mF;
            for (int i = 0; i < 6; ++i)
                for (int j = 0; j < 6; ++j) M[i][j] += colF[i] * ev[j] + ev[i] * colF[j] + mFF * ev[i] * ev[j];
            // (m already carries M*t0 through colF*e0; add the (Fel,Fel) part routed through ev)
            for (int i = 0; i < 6; ++i) { M[i][3] = 0.0; M[3][i] = 0.0; }
            M[3][3] = 1.0; m[3] = 0.0;
        }
        // Cholesky of the control block (indices 3..5)
        double l00 = M[3][3];
        if (!(l00 > 0.0) || !isfinite(l00)) return false;
        l00 = sqrt(l00);
        double l10 = M[4][3] / l00, l20 = M[5][3] / l00;
        double l11 = M[4][4] - l10 * l10;
        if (!(l11 > 0.0) || !isfinite(l11)) return false;
        l11 = sqrt(l11);
        double l21 = (M[5][4] - l20 * l10) / l11;
        double l22 = M[5][5] - l20 * l20 - l21 * l21;
        if (!(l22 > 0.0) || !isfinite(l22)) return false;
        l22 = sqrt(l22);
        // solve Muu X = [Mux mu]  (4 right-hand sides)
        double K[3][3], kf[3];
        for (int j = 0; j < 4; ++j) {
            double r0 = (j < 3) ? M[3][j] : m[3], r1 = (j < 3) ? M[4][j] : m[4], r2 = (j < 3) ? M[5][j] : m[5];
            double y0 = r0 / l00, y1 = (r1 - l10 * y0) / l11, y2 = (r2 - l20 * y0 - l21 * y1) / l22;
            double x2 = y2 / l22, x1 = (y1 - l21 * x2) / l11, x0 = (y0 - l10 * x1 - l20 * x2) / l00;
            if (j < 3) { K[0][j] = -x0; K[1][j] = -x1; K[2][j] = -x2; }
            else { kf[0] = -x0; kf[1] = -x1; kf[2] = -x2; }
        }
        // P = Mxx + Mxu K ; p = mx + Mxu kf
        double Pn[3][3], pnv[3];
        for (int i = 0; i < 3; ++i) {
            pnv[i] = m[i] + M[i][3] * kf[0] + M[i][4] * kf[1] + M[i][5] * kf[2];
            for (int j = 0; j < 3; ++j) Pn[i][j] = M[i][j] + M[i][3] * K[0][j] + M[i][4] * K[1][j] + M[i][5] * K[2][j];
        }
        for (int i = 0; i < 3; ++i) { p[i] = pnv[i]; for (int j = 0; j < 3; ++j) P[i][j] = 0.5 * (Pn[i][j] + Pn[j][i]); }
        if (last) {   // recover the eliminated control's feedback row
            for (int j = 0; j < 3; ++j) K[0][j] = ePn * K[1][j];
            K[0][1] += eB;
            kf[0] = e0 + ePn * kf[1];
        }
        for (int i = 0; i < 3; ++i) {
            for (int j = 0; j < 3; ++j) c.W(WS_RIC + RIC_K + 3 * i + j, k, s) = K[i][j];
            c.W(WS_RIC + RIC_KF + i, k, s) = kf[i];
            c.W(WS_RIC + RIC_PV + i, k, s) = p[i];
        }
        c.W(WS_RIC + RIC_P + 0, k, s) = P[0][0]; c.W(WS_RIC + RIC_P + 1, k, s) = P[0][1];
        c.W(WS_RIC + RIC_P + 2, k, s) = P[0][2]; c.W(WS_RIC + RIC_P + 3, k, s) = P[1][1];
        c.W(WS_RIC + RIC_P + 4, k, s) = P[1][2]; c.W(WS_RIC + RIC_P + 5, k, s) = P[2][2];
    }
    return true;
}

struct Ftb {
    double aP, aZ, gphid;
};
MS_HD void ftb_bound(Ftb& f, double tau, double mu, double z, double slack, double dvSigned, bool oneSided) {
    // dvSigned = change of the slack; primal fraction-to-boundary, dual step, barrier directional derivative
    if (dvSigned < 0.0) f.aP = fmin(f.aP, -tau * slack / dvSigned);
    double dz = mu / slack - z - (z / slack) * dvSigned;
    if (dz < 0.0) f.aZ = fmin(f.aZ, -tau * z / dz);
    f.gphid += (-mu / slack + (oneSided ? MS_KAPPA_D * mu : 0.0)) * dvSigned;
}

// forward sweep: primal step, new multipliers, slack / bound-multiplier steps, step-size limits
MS_HD void riccati_forward(const Ctx& c, int s, int N, double mu, double tauF, double delta, Ftb& f) {
    const Config& g = c.cfg;
    const int it = c.I(SI_PARITY, s) ? WS_IT1 : WS_IT0;
    const double scale = c.P(P_SCALE, s);
    double dx[3] = {0.0, 0.0, 0.0};
    double felPrev = 0.0, dfelPrev = 0.0;
    f.aP = 1.0; f.aZ = 1.0; f.gphid = 0.0;
    c.W(WS_ST + ST_T, 0, s) = 0.0;
    c.W(WS_ST + ST_B, 0, s) = 0.0;
    for (int k = 0; k < N; ++k) {
        Bnd B = load_bounds(c, k, s);
        double du[3];
        for (int i = 0; i < 3; ++i)
            du[i] = c.W(WS_RIC + RIC_KF + i, k, s) + c.W(WS_RIC + RIC_K + 3 * i + 0, k, s) * dx[0]
                  + c.W(WS_RIC + RIC_K + 3 * i + 1, k, s) * dx[1] + c.W(WS_RIC + RIC_K + 3 * i + 2, k, s) * dx[2];
        if (!g.withPn) du[1] = 0.0;
        const double tb = c.W(WS_QP + QP_TAU_B, k, s), tF = c.W(WS_QP + QP_TAU_F, k, s);
        const double pb = c.W(WS_QP + QP_PHI_B, k, s), pF = c.W(WS_QP + QP_PHI_F, k, s);
        const double dF = du[0] + du[1];
        double dxn[3];
        dxn[0] = dx[0] + tb * dx[1] + tF * dF + c.W(WS_QP + QP_RT, k, s);
        dxn[1] = (k + 1 < N) ? pb * dx[1] + pF * dF + c.W(WS_QP + QP_RB, k, s) : 0.0;
        dxn[2] = du[0];
        // costates of the next node
        double pit = c.W(WS_RIC + RIC_PV + 0, k + 1, s) + c.W(WS_RIC + RIC_P + 0, k + 1, s) * dxn[0]
                   + c.W(WS_RIC + RIC_P + 1, k + 1, s) * dxn[1] + c.W(WS_RIC + RIC_P + 2, k + 1, s) * dxn[2];
        double pib;
        if (k + 1 < N) {
            pib = c.W(WS_RIC + RIC_PV + 1, k + 1, s) + c.W(WS_RIC + RIC_P + 1, k + 1, s) * dxn[0]
                + c.W(WS_RIC + RIC_P + 3, k + 1, s) * dxn[1] + c.W(WS_RIC + RIC_P + 4, k + 1, s) * dxn[2];
            pib += c.W(WS_QP + QP_HC_B, k, s) * dx[1] + c.W(WS_QP + QP_HC_FEL, k, s) * du[0]
                 + c.W(WS_QP + QP_HC_FPB, k, s) * du[1] + c.W(WS_QP + QP_HC_SL, k, s) * du[2]
                 + c.W(WS_QP + QP_HPP, k, s) * dxn[1] + c.W(WS_QP + QP_GP0, k, s) + mu * c.W(WS_QP + QP_GP1, k, s);
        } else {
            // b_N is fixed: its row multiplier follows from stationarity w.r.t. Fel of the last interval
            double gF = c.W(WS_QP + QP_G0_FEL, k, s) + mu * c.W(WS_QP + QP_G1_FEL, k, s)
                      + c.W(WS_QP + QP_H_BFEL, k, s) * dx[1] + c.W(WS_QP + QP_H_FFEL, k, s) * dx[2]
                      + (c.W(WS_QP + QP_H_FELFEL, k, s) + delta) * du[0] + c.W(WS_QP + QP_H_FELFPB, k, s) * du[1]
                      + c.W(WS_QP + QP_H_FELSL, k, s) * du[2];
            pib = -(gF + tF * pit) / pF;
        }
        const double ytNew = -pit, ybNew = -pib;
        // ---- store the primal / equality-multiplier step
        c.W(WS_ST + ST_FEL, k, s) = du[0];
        c.W(WS_ST + ST_FPB, k, s) = du[1];
        c.W(WS_ST + ST_SL, k, s) = du[2];
        c.W(WS_ST + ST_T, k + 1, s) = dxn[0];
        c.W(WS_ST + ST_B, k + 1, s) = dxn[1];
        c.W(WS_ST + ST_YT, k, s) = ytNew - c.W(it + IT_YT, k, s);
        c.W(WS_ST + ST_YB, k, s) = ybNew - c.W(it + IT_YB, k, s);
        // ---- bounds on the variables of this interval
        const double fel = c.W(it + IT_FEL, k, s), fpb = c.W(it + IT_FPB, k, s), sl = c.W(it + IT_SL, k, s);
        if (k >= 1) {
            const double t = c.W(it + IT_T, k, s), b = c.W(it + IT_B, k, s);
            ftb_bound(f, tauF, mu, c.W(it + IT_Z + Z_T_L, k, s), t - B.tL, dx[0], false);
            ftb_bound(f, tauF, mu, c.W(it + IT_Z + Z_T_U, k, s), B.tU - t, -dx[0], false);
            ftb_bound(f, tauF, mu, c.W(it + IT_Z + Z_B_L, k, s), b - B.bL, dx[1], false);
            ftb_bound(f, tauF, mu, c.W(it + IT_Z + Z_B_U, k, s), B.bU - b, -dx[1], false);
        }
        ftb_bound(f, tauF, mu, c.W(it + IT_Z + Z_FEL_L, k, s), fel - B.felL, du[0], false);
        ftb_bound(f, tauF, mu, c.W(it + IT_Z + Z_FEL_U, k, s), B.felU - fel, -du[0], false);
        if (g.withPn) {
            ftb_bound(f, tauF, mu, c.W(it + IT_Z + Z_FPB_L, k, s), fpb - B.fpbL, du[1], false);
            ftb_bound(f, tauF, mu, c.W(it + IT_Z + Z_FPB_U, k, s), B.fpbU - fpb, -du[1], false);
        }
        ftb_bound(f, tauF, mu, c.W(it + IT_Z + Z_SL_L, k, s), sl - B.slL, du[2], true);
        // ---- objective part of the barrier directional derivative
        if (g.energy) {
            f.gphid += c.W(WS_TRK + TRK_DS, k, s) * (du[0] + du[2]) / scale;
            if (k >= 1) f.gphid += (2e-3 / scale) * (fel - felPrev) * (du[0] - dfelPrev);
        } else {
            f.gphid += (2e-4 / scale) * (fel * du[0] + fpb * du[1]);
        }
        felPrev = fel; dfelPrev = du[0];
        // ---- inequality rows: slack step, multiplier step
        for (int j = 0; j < NROW; ++j) {
            c.W(WS_ST + ST_W + j, k, s) = 0.0;
            c.W(WS_ST + ST_YD + j, k, s) = 0.0;
            if (!row_on(g, j)) continue;
            double jd;
            if (j == R_P0) jd = c.W(WS_QP + QP_J_P0_B, k, s) * dx[1] + c.W(WS_QP + QP_J_P0_FEL, k, s) * du[0];
            else if (j == R_P1) jd = c.W(WS_QP + QP_J_P1_FEL, k, s) * du[0] + c.W(WS_QP + QP_J_P1_BN, k, s) * dxn[1];
            else if (j == R_ACC) jd = c.W(WS_QP + QP_J_ACC_B, k, s) * dx[1] + du[0] + du[1];
            else if (j == R_LTR) jd = du[2] + c.W(WS_QP + QP_J_LTR_FEL, k, s) * du[0] + c.W(WS_QP + QP_J_LTR_B, k, s) * dx[1]
                                    + c.W(WS_QP + QP_J_LTR_BN, k, s) * dxn[1];
            else jd = du[2] + c.W(WS_QP + QP_J_LRG_FEL, k, s) * du[0] + c.W(WS_QP + QP_J_LRG_B, k, s) * dx[1]
                    + c.W(WS_QP + QP_J_LRG_BN, k, s) * dxn[1];
            const double dw = jd + c.W(WS_QP + QP_RES + j, k, s);
            double L, U; bool hasU;
            row_bounds(B, j, L, U, hasU);
            const int zl = (j == R_P0) ? Z_P0_L : (j == R_P1) ? Z_P1_L : (j == R_ACC) ? Z_ACC_L : (j == R_LTR) ? Z_LTR_L : Z_LRG_L;
            const double w = c.W(it + IT_W + j, k, s);
            const double vL = c.W(it + IT_Z + zl, k, s), sL = w - L;
            double sig = vL / sL, gw = -mu / sL + (hasU ? 0.0 : MS_KAPPA_D * mu);
            ftb_bound(f, tauF, mu, vL, sL, dw, !hasU);
            if (hasU) {
                const double vU = c.W(it + IT_Z + zl + 1, k, s), sU = U - w;
                sig += vU / sU; gw += mu / sU;
                ftb_bound(f, tauF, mu, vU, sU, -dw, false);
            }
            c.W(WS_ST + ST_W + j, k, s) = dw;
            c.W(WS_ST + ST_YD + j, k, s) = sig * dw + gw - c.W(it + IT_YD + j, k, s);
        }
        dx[0] = dxn[0]; dx[1] = dxn[1]; dx[2] = dxn[2];
    }
    // terminal node: t_N
    {
        Bnd B = load_bounds(c, N, s);
        const double t = c.W(it + IT_T, N, s);
        ftb_bound(f, tauF, mu, c.W(it + IT_Z + Z_T_L, N, s), t - B.tL, dx[0], false);
        ftb_bound(f, tauF, mu, c.W(it + IT_Z + Z_T_U, N, s), B.tU - t, -dx[0], false);
        if (!g.energy) f.gphid += dx[0] / scale;
    }
}

// ------------------------------------------------------------------------------------------------
// per-instance driver: KKT error, termination, barrier update, search direction   (IPOPT Alg. A, steps A-1..A-4)
// ------------------------------------------------------------------------------------------------
MS_HD void inst_step(const Ctx& c, int s) {
    const Config& g = c.cfg;
    if (s >= g.nInst || c.I(SI_PHASE, s) != PH_EVAL) return;
    const int N = c.I(SI_N_INT, s);
    const int it = c.I(SI_PARITY, s) ? WS_IT1 : WS_IT0;
    count_cells(c, 1, N + 1);
    double th = 0.0, fo = 0.0, slog = 0.0, sdamp = 0.0, dinf = 0.0, pinf = 0.0, cmin = 1e300, cmax = 0.0, zsum = 0.0, ysum = 0.0;
    double cnPrev = 0.0, ytPrev = 0.0;
    for (int k = 0; k <= N; ++k) {
        th += c.W(WS_PART + PC_TH, k, s);
        fo += c.W(WS_PART + PC_F, k, s);
        slog += c.W(WS_PART + PC_SLOG, k, s);
        sdamp += c.W(WS_PART + PC_SDAMP, k, s);
        dinf = fmax(dinf, c.W(WS_PART + PC_DINF, k, s));
        pinf = fmax(pinf, c.W(WS_PART + PC_PINF, k, s));
        cmin = fmin(cmin, c.W(WS_PART + PC_CMIN, k, s));
        cmax = fmax(cmax, c.W(WS_PART + PC_CMAX, k, s));
        zsum += c.W(WS_PART + PC_ZSUM, k, s);
        ysum += c.W(WS_PART + PC_YSUM, k, s);
        if (k >= 1) {
            dinf = fmax(dinf, fabs(c.W(WS_PART + PC_OWN_T, k, s) + ytPrev));
            if (k < N) dinf = fmax(dinf, fabs(c.W(WS_PART + PC_OWN_B, k, s) + cnPrev));
        }
        if (k < N) { cnPrev = c.W(WS_PART + PC_CN_B, k, s); ytPrev = c.W(it + IT_YT, k, s); }
    }
    // counts for the IPOPT error scaling s_d, s_c (eq. 6)
    const int nrow = (g.withPower ? 2 : 0) + 1 + (g.energy ? 2 : 0);
    const int nbRow = (g.withPower ? 4 : 0) + 2 + (g.energy ? 2 : 0);
    const int nb = N * (2 + (g.withPn ? 2 : 0) + 1 + nbRow) + (N - 1) * 4 + 2;
    const int mrows = N * (nrow + 2);
    const double sd = fmax(100.0, (ysum + zsum) / (mrows + nb)) / 100.0;
    const double sc = fmax(100.0, zsum / nb) / 100.0;
    const double E0 = fmax(fmax(dinf / sd, pinf), cmax / sc);
    c.D(SD_THETA, s) = th; c.D(SD_FOBJ, s) = fo; c.D(SD_SLOG, s) = slog; c.D(SD_SDAMP, s) = sdamp;
    c.D(SD_KKT, s) = E0; c.D(SD_DINF, s) = dinf; c.D(SD_PINF, s) = pinf; c.D(SD_CINF, s) = cmax;
    if (c.D(SD_THETA_MAX, s) < 0.0) {
        c.D(SD_THETA_MAX, s) = 1e4 * fmax(1.0, th);
        c.D(SD_THETA_MIN, s) = 1e-4 * fmax(1.0, th);
    }
    if (!isfinite(E0) || !isfinite(fo)) { finish(c, s, ST_INVALID_NUMBER); return; }
    if (E0 <= g.tol) { finish(c, s, ST_SOLVE_SUCCEEDED); return; }
    if (c.I(SI_ITERS, s) >= g.maxIter) { finish(c, s, ST_MAXITER); return; }
    // ---- monotone barrier update (eq. 7), filter reset
    double mu = c.D(SD_MU, s);
    for (;;) {
        double cinf = fmax(cmax - mu, mu - cmin);
        double Emu = fmax(fmax(dinf / sd, pinf), cinf / sc);
        if (Emu <= 10.0 * mu && mu > g.tol / 10.0 * (1.0 + 1e-12)) {
            mu = fmax(g.tol / 10.0, fmin(0.2 * mu, pow(mu, 1.5)));
            c.I(SI_NFILT, s) = 0;
        } else break;
    }
    const double tauF = fmax(0.99, 1.0 - mu);
    c.D(SD_MU, s) = mu; c.D(SD_TAU, s) = tauF;
    // ---- search direction with inertia correction (IPOPT Alg. IC)
    double delta = 0.0;
    const double dlast = c.D(SD_DELTA_LAST, s);
    bool ok = false;
    for (int tries = 0; tries < 40; ++tries) {
        count_cells(c, 2, N);
        if (riccati_backward(c, s, N, mu, delta)) { ok = true; break; }
        c.I(SI_NREG, s) += 1;
        if (delta == 0.0) delta = (dlast == 0.0) ? 1e-4 : fmax(1e-20, dlast / 3.0);
        else delta *= (dlast == 0.0) ? 100.0 : 8.0;
        if (delta > 1e40) break;
    }
    if (!ok) { finish(c, s, ST_STEP_FAILED); return; }
    if (delta > 0.0) c.D(SD_DELTA_LAST, s) = delta;
    Ftb f;
    count_cells(c, 3, N);
    riccati_forward(c, s, N, mu, tauF, delta, f);
    if (!isfinite(f.gphid) || !isfinite(f.aP)) { finish(c, s, ST_STEP_FAILED); return; }
    double amin;
    if (f.gphid < 0.0) {
        amin = 1e-5;
        if (th > 0.0) amin = fmin(amin, 1e-8 * th / (-f.gphid));
        if (th <= c.D(SD_THETA_MIN, s)) amin = fmin(amin, pow(th, 1.1) / pow(-f.gphid, 2.3));
        amin *= 0.05;
    } else amin = 0.05 * 1e-5;
    c.D(SD_GPHID, s) = f.gphid;
    c.D(SD_ALPHA, s) = f.aP;
    c.D(SD_ALPHA_Z, s) = f.aZ;
    c.D(SD_ALPHA_MIN, s) = amin;
    c.I(SI_NLS, s) = 0;
    c.I(SI_PHASE, s) = PH_TRIAL;
}

}  // namespace mseetc
