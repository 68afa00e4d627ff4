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

// initial guess pushed into the relaxed bounds: the reference's (ocp.py:325-339) or, with initMode 1, the speed-envelope
// profile that inst_profile left in the planes of iterate buffer 0
MS_HD double init_b(const Ctx& c, int j, int s, int N) {
    if (j == 0) return c.P(P_B0, s);
    if (j == N) return c.P(P_BN, s);
    const double guess = c.cfg.initMode ? c.W(WS_IT1 + IT_B, j, s) : MS_VEL0SQ;
    return push2(guess, relaxL(c.P(P_BMIN, s)), relaxU(c.W(WS_TRK + TRK_BMAX, j, s)));
}
MS_HD double init_t(const Ctx& c, int j, int s, int N) {
    double t0 = c.P(P_T0, s), T = c.P(P_T, s);
    if (j == 0) return t0;
    const double guess = c.cfg.initMode ? c.W(WS_IT1 + IT_T, j, s) : t0 + j * ((T - t0) / N);
    return push2(guess, relaxL(t0), relaxU(T));
}

// Early infeasibility screening (only when the caller asked for screening by passing a tmin plane): a lower bound on the
// trip duration from the speed envelope at 100 % of the force, power, acceleration and speed limits -- fastest
// acceleration from b_0, latest braking into b_N, never above the limits.  No admissible run is faster, up to
// discretisation effects that MS_SCREEN_MARGIN covers (with piecewise-constant forces b is monotone inside an interval, so
// the discrete run respects the limits between the nodes too and is a restriction of the continuous problem; what remains
// are the Euler steps of the envelope and the RK4 / quadrature errors of the NLP, O(1e-4) relative at the usual grids),
// and terminalTime is an upper bound on t_N (ocp.py:260-261): an
// instance whose available time is below (1 - margin) * bound is reported infeasible before a single iteration is
// spent on it.  The exact certificate is the minimum trip time (time-optimal solve, running concurrently); the host
// layer checks every early flag against it and re-solves an instance without screening should a flag ever be wrong.
#define MS_SCREEN_MARGIN 0.01
// Relative margin of the exact certificate (minimum trip time from the time-optimal solve): that solve minimises
// t_N + 1e-4 * sum(F^2) (ocp.py:146-150), so its t_N can sit slightly above the true minimum; trips within the margin below
// it are left to the iteration instead of being reported infeasible.
#define MS_TMIN_MARGIN 1e-6
// Both per-instance set-up routines below walk the track sequentially (one thread per instance).  Their loops work on chunks
// of MS_PCH intervals -- all loads of a chunk are issued before the first dependent operation and the results are stored after
// the last one -- so a pass costs one memory latency per chunk instead of one per interval.  Square roots and reciprocals are the
// short hardware-approximation + Newton sequences of jet.cuh (1-2 ulp): these are dependent chains, one link per interval, and the
// values only shape the starting point and the 1 % screening test.
#define MS_PCH 8
struct TrainLimits {
    double sr0, sr1, sr2, felU, felL, fpbL, pUp, pLo, aLo, aUp, b0, bN, bmin;
};
MS_HD TrainLimits load_limits(const Ctx& c, int s) {
    const Config& g = c.cfg;
    TrainLimits q;
    q.sr0 = c.P(P_SR0, s); q.sr1 = c.P(P_SR1, s); q.sr2 = c.P(P_SR2, s);
    q.felU = c.P(P_FEL_U, s); q.felL = c.P(P_FEL_L, s); q.fpbL = g.withPn ? c.P(P_FPB_L, s) : 0.0;
    q.pUp = g.withPower ? c.P(P_P_UP, s) : 1e30; q.pLo = g.withPower ? c.P(P_P_LO, s) : -1e30;
    q.aLo = c.P(P_A_LO, s); q.aUp = c.P(P_A_UP, s);
    q.b0 = c.P(P_B0, s); q.bN = c.P(P_BN, s); q.bmin = c.P(P_BMIN, s);
    return q;
}
// speed envelope: fastest acceleration from b_0 (forward), latest braking into b_N (backward), with the fraction `frac` of the
// force, power and acceleration limits and the speed limits scaled by `flim`; the node values go to plane `dst`.
// Returns the trip time of the envelope (trapezoidal in 1/v, exact for b linear in s).
MS_HD double speed_envelope(const Ctx& c, int s, int N, const TrainLimits& q, double frac, double flim, double vfloor, int brakeIters, int dst,
                             double* bmaxAll) {
    double b = q.b0;
    c.W(dst, 0, s) = q.b0;
    for (int k0 = 0; k0 < N; k0 += MS_PCH) {
        double ds[MS_PCH], c0[MS_PCH], lim[MS_PCH], out[MS_PCH];
#pragma unroll
        for (int i = 0; i < MS_PCH; ++i) {
            const int k = k0 + i;
            if (k < N) {
                ds[i] = c.W(WS_TRK + TRK_DS, k, s); c0[i] = c.W(WS_TRK + TRK_C0, k, s);
                lim[i] = (k + 1 < N) ? flim * c.W(WS_TRK + TRK_BMAX, k + 1, s) : q.bN;
            }
        }
#pragma unroll
        for (int i = 0; i < MS_PCH; ++i) {
            if (k0 + i < N) {
                const double v = fsqrt(b);
                const double r = q.sr0 + q.sr1 * v + q.sr2 * b + c0[i];
                const double a = fmin(frac * (fmin(q.felU, q.pUp * rcp_slack(fmax(v, vfloor))) - r), frac * q.aUp);
                b = fmax(q.bmin, fmin(lim[i], b + 2.0 * ds[i] * a));
                out[i] = b;
            }
        }
#pragma unroll
        for (int i = 0; i < MS_PCH; ++i) if (k0 + i < N) c.W(dst, k0 + i + 1, s) = out[i];
    }
    b = q.bN;
    c.W(dst, N, s) = q.bN;
    double vn = fsqrt(q.bN), tt = 0.0, bmx = 0.0;
    for (int k1 = N - 1; k1 >= 0; k1 -= MS_PCH) {
        double ds[MS_PCH], c0[MS_PCH], fw[MS_PCH], out[MS_PCH];
#pragma unroll
        for (int i = 0; i < MS_PCH; ++i) {
            const int k = k1 - i;
            if (k >= 0) { ds[i] = c.W(WS_TRK + TRK_DS, k, s); c0[i] = c.W(WS_TRK + TRK_C0, k, s); fw[i] = c.W(dst, k, s); }
        }
#pragma unroll
        for (int i = 0; i < MS_PCH; ++i) {
            const int k = k1 - i;
            if (k >= 0) {
                double bk = q.b0;
                if (k >= 1) {
                    bk = b;
                    for (int it = 0; it < brakeIters; ++it) {   // the braking deceleration depends on the speed at the start of the interval
                        const double v = fsqrt(bk);
                        const double r = q.sr0 + q.sr1 * v + q.sr2 * bk + c0[i];
                        const double a = fmax(frac * (fmax(q.felL, q.pLo * rcp_slack(fmax(v, vfloor))) + q.fpbL - r), frac * q.aLo);      // negative
                        bk = b - 2.0 * ds[i] * a;
                    }
                    bk = fmin(fw[i], bk);
                    bmx = fmax(bmx, bk);
                }
                const double vk = fsqrt(bk);
                tt += 2.0 * ds[i] * rcp_slack(vk + vn);
                b = bk; vn = vk;
                out[i] = bk;
            }
        }
#pragma unroll
        for (int i = 0; i < MS_PCH; ++i) if (k1 - i >= 1) c.W(dst, k1 - i, s) = out[i];
    }
    if (bmaxAll) *bmaxAll = bmx;
    return tt;
}

MS_HD void inst_screen(const Ctx& c, int s) {
    const Config& g = c.cfg;
    if (s >= g.nInst || !g.energy || !c.tmin || c.I(SI_PHASE, s) == PH_DONE) return;
    if (c.tmin[c.I(SI_ORIG, s)] < 0.0) return;        // the caller asserted that this instance is feasible: never screened
    const int N = c.I(SI_N_INT, s);
    const TrainLimits q = load_limits(c, s);
    // scratch plane: a step plane (zeroed by cell_init afterwards); inst_profile, which may run concurrently in another block,
    // works in the planes of iterate buffer 1
    const double tLow = speed_envelope(c, s, N, q, 1.0, 1.0, 1e-3, 2, WS_ST + ST_B, nullptr);
    if (isfinite(tLow) && (c.P(P_T, s) - c.P(P_T0, s)) < (1.0 - MS_SCREEN_MARGIN) * tLow) finish(c, s, ST_INFEASIBLE);
}

// Dynamically consistent starting profile (initMode 1): speed envelope from the limits with bounded acceleration and
// braking, cruise speed capped so that the trip takes the available time, times and forces from the ODE.  It replaces
// the constant-speed guess of the reference only as a starting point; the NLP and its optimum are unchanged.
// `smv`, `smd` (optional, stride `sms` doubles between consecutive k): shared-memory columns for the node speeds and the
// interval lengths, which the cruise-cap search reads once per evaluation.
MS_HD void inst_profile(const Ctx& c, int s, double* smv = nullptr, double* smd = nullptr, int sms = 0) {
    const Config& g = c.cfg;
    if (s >= g.nInst || !g.initMode || c.I(SI_PHASE, s) == PH_DONE) return;
    const int N = c.I(SI_N_INT, s);
    const int P = WS_IT1;                       // scratch: buffer 1 holds the profile until cell_init has consumed it
    const TrainLimits q = load_limits(c, s);
    const double Tav = c.P(P_T, s) - c.P(P_T0, s);
    // ---- fastest admissible profile: accelerate / brake with 80 % of what the force, power and acceleration limits allow
    double bmaxAll = 0.0;
    const double tFast = speed_envelope(c, s, N, q, 0.8, 0.97, 1.0, 1, P + IT_B, &bmaxAll);
    // ---- energy mode: cap the cruise speed so that the trip uses (almost all of) the available time.  The trip time is a
    // decreasing function of the cap: bracketing + regula falsi (Illinois) on the speed
    double cap = 1e30;
    const double target = 0.995 * Tav;
    if (g.energy && tFast < target) {
        const int V = P + IT_SL;                  // scratch when no shared memory is given (IT_SL is written last, below)
        if (!smv) { smv = &c.W(V, 0, s); smd = &c.W(WS_TRK + TRK_DS, 0, s); sms = WS_FIELDS * 32; }
        const bool copyDs = (smd != &c.W(WS_TRK + TRK_DS, 0, s));
        for (int k0 = 0; k0 <= N; k0 += MS_PCH) {
            double bb[MS_PCH], dd[MS_PCH];
#pragma unroll
            for (int i = 0; i < MS_PCH; ++i)
                if (k0 + i <= N) { bb[i] = c.W(P + IT_B, k0 + i, s); dd[i] = (copyDs && k0 + i < N) ? c.W(WS_TRK + TRK_DS, k0 + i, s) : 0.0; }
#pragma unroll
            for (int i = 0; i < MS_PCH; ++i)
                if (k0 + i <= N) { smv[(size_t)(k0 + i) * sms] = fsqrt(bb[i]); if (copyDs && k0 + i < N) smd[(size_t)(k0 + i) * sms] = dd[i]; }
        }
        auto trip = [&](double vcap) {            // interior nodes capped, boundary speeds kept
            double tt = 0.0, vp = smv[0];
#pragma unroll 8
            for (int k = 0; k < N; ++k) {
                const double vk = smv[(size_t)(k + 1) * sms];
                const double vn = (k + 1 < N) ? fmin(vk, vcap) : vk;
                tt += 2.0 * smd[(size_t)k * sms] * rcp_slack(vp + vn);
                vp = vn;
            }
            return tt;
        };
        double lo = fsqrt(q.bmin), hi = fsqrt(bmaxAll);
        double fhi = tFast - target;              // fastest profile: negative, there is slack in the timetable
        if (hi > lo) {
            double flo = trip(lo) - target;       // slowest cap: positive unless even crawling is too fast
            if (flo > 0.0) {
                double vc = hi;
                int side = 0;
                for (int it = 0; it < 40 && hi - lo > 5e-4 * hi; ++it) {
                    const double x = fmin(fmax((lo * fhi - hi * flo) / (fhi - flo), lo + 1e-3 * (hi - lo)), hi - 1e-3 * (hi - lo));
                    const double fx = trip(x) - target;
                    if (fabs(fx) <= 2e-4 * target) { vc = x; break; }      // within 0.02 % of the target time: good enough for a start
                    if (fx > 0.0) { lo = x; flo = fx; if (side < 0) fhi *= 0.5; side = -1; }
                    else { hi = x; fhi = fx; if (side > 0) flo *= 0.5; side = 1; }
                    vc = hi;
                }
                cap = vc * vc;
            } else cap = lo * lo;
        }
    }
    // ---- times, forces and epigraph variable of the profile
    double t = c.P(P_T0, s);
    c.W(P + IT_T, 0, s) = t;
    double b = q.b0, vb = fsqrt(q.b0);
    for (int k0 = 0; k0 < N; k0 += MS_PCH) {
        double bnx[MS_PCH], ds[MS_PCH], c0[MS_PCH], ot[MS_PCH], of[MS_PCH], op[MS_PCH];
#pragma unroll
        for (int i = 0; i < MS_PCH; ++i) {
            const int k = k0 + i;
            if (k < N) { bnx[i] = c.W(P + IT_B, k + 1, s); ds[i] = c.W(WS_TRK + TRK_DS, k, s); c0[i] = c.W(WS_TRK + TRK_C0, k, s); }
        }
#pragma unroll
        for (int i = 0; i < MS_PCH; ++i) {
            const int k = k0 + i;
            if (k < N) {
                const double bn = (k + 1 < N) ? fmin(bnx[i], cap) : bnx[i];
                bnx[i] = bn;
                const double vbn = fsqrt(bn);
                t += 2.0 * ds[i] * rcp_slack(vb + vbn);
                ot[i] = t;
                const double bm = 0.5 * (b + bn);
                const double F = (bn - b) * rcp_slack(2.0 * ds[i]) + q.sr0 + q.sr1 * fsqrt(bm) + q.sr2 * bm + c0[i];
                const double ivmx = 0.97 * rcp_slack(fmax(fmax(vb, vbn), 1.0));
                const double fel = fmin(fmax(F, fmax(0.97 * q.felL, q.pLo * ivmx)), fmin(0.97 * q.felU, q.pUp * ivmx));
                of[i] = fel;
                op[i] = g.withPn ? fmin(0.0, fmax(0.97 * q.fpbL, F - fel)) : 0.0;
                b = bn; vb = vbn;
            }
        }
#pragma unroll
        for (int i = 0; i < MS_PCH; ++i) {
            const int k = k0 + i;
            if (k < N) {
                if (k + 1 < N) c.W(P + IT_B, k + 1, s) = bnx[i];
                c.W(P + IT_T, k + 1, s) = ot[i];
                c.W(P + IT_FEL, k, s) = of[i];
                c.W(P + IT_FPB, k, s) = op[i];
                c.W(P + IT_SL, k, s) = 0.5 * fabs(of[i]) + 0.02;
            }
        }
    }
}

MS_HD LossPar load_losspar(const Ctx& c, int s) {
    LossPar p;
    const double eg = c.P(P_DYN_ETAG, s);
    p.M = c.P(P_MASS, s); p.aux = c.P(P_DYN_AUX, s); p.cgT = (1.0 - eg) / eg; p.cgB = 1.0 - eg;
    p.fMax = c.P(P_DYN_FMAX, s); p.pMax = c.P(P_DYN_PMAX, s); p.scale = c.P(P_DYN_SCALE, s);
    return p;
}

// values of the inequality rows at a point                                  (ocp.py:189,199,225-226)
template <bool DYN, bool INTL = false>
MS_HD void ineq_values(const Ctx& c, int s, double fel, double fpb, double sl, double b0, double b1,
                       const IntervalCoef& q, double* d) {
    // every operation individually rounded (see mul_rn): cell_eval and cell_step both evaluate these rows
    const double r0 = fsqrt(b0);
    d[R_P0] = mul_rn(fel, r0);
    d[R_P1] = mul_rn(fel, fsqrt(b1));
    d[R_ACC] = sub_rn(sub_rn(add_rn(fel, fpb), fma(q.sr1, r0, mul_rn(q.sr2, b0))), add_rn(q.sr0, q.c0));
    if (INTL) {
        d[R_LTR] = 0.0; d[R_LRG] = 0.0;      // integrated losses: filled in by the caller (loss_energy_rows)
    } else if (DYN) {
        LossRow tr, rg;
        loss_rows_dynamic(c.lm, load_losspar(c, s), fel, b0, b1, tr, rg);
        d[R_LTR] = sl - tr.v;
        d[R_LRG] = sl - rg.v;
    } else {
        d[R_LTR] = fma(-c.P(P_CT, s), fel, sl);
        d[R_LRG] = fma(c.P(P_CR, s), fel, sl);
    }
}

// ---- integrateLosses = True                                            (ocp.py:231-241, train.py:367-413)
// The reference's rows are  s_k - E(sqrt(b_k), t_{k+1} - t_k, Fel_k, Fpb_k) >= 0  with E = (eTr, eBr) at the end of the time-domain
// integration of  dv/dt = a(v^2, F),  d eTr/dt = PLtr(Fel, v),  d eBr/dt = PLrgb(Fel, v)  over the interval's duration from v =
// sqrt(b_k), eTr = eBr = 0 (specific quantities; CVODES with relTol 1e-6 in the reference).  Here the duration is taken from the
// shooting function of the same interval, t_{k+1} - t_k = tau(b_k, Fel_k + Fpb_k): the shooting row t_{k+1} - t_k - tau = 0 is an
// equality constraint of the NLP, so feasible set, objective and minimisers are those of the reference's formulation, while the
// rows stay functions of (b_k, Fel_k, Fpb_k, s_k) -- the variables the stage QP of the sweeps carries.
// The integration runs over the normalised time theta in [0,1] with MS_INTL_STEPS classic RK4 steps and second-order jets in
// (b_k, Fel_k, Fpb_k); MS_INTL_STEPS_KINK steps in an interval in which the speed crosses a kink of the spline loss map (the ends of
// its speed range and the turning speed Pmax/Fmax, efficiency.py:10-12,40): the integrand is only continuous there, and the error of
// 4 steps (6e-5 of the total losses on a 60-interval grid) would exceed the tolerance of the reference's CVODES integration.
#define MS_INTL_STEPS 4
#ifndef MS_INTL_CLAMP
#define MS_INTL_CLAMP true
#endif
#ifndef MS_INTL_STEPS_KINK
#define MS_INTL_STEPS_KINK 32
#endif
MS_HD void power_loss_jets(const Ctx& c, int s, const LossPar& lp, bool pos, const Jet3& FEL, const Jet3& v, Jet3& ptr, Jet3& prg) {
    if (c.cfg.lossKind == 2) {
        const Jet2 vv = jvar0(v.v), ff = jvar1(FEL.v);       // the map returns PL/v with partials w.r.t. (v, specific force)
        // The motor map is zero outside its grid (efficiency.py:40-51,137); its upper load edge is the power hyperbola F v = Pmax, where
        // the power rows are active at the nodes.  The last stage of the integration reproduces the node speed only to the integration
        // error and may land 1e-6 beyond the edge: the load is clamped to the edge there (an adaptive integrator like the reference's
        // CVODES loses a vanishing sliver at that point, a fixed-step stage would lose a sixth of a step).
        const Jet2 qt = pos ? loss_full(c.lm, lp, vv, ff, true, MS_INTL_CLAMP) : loss_tangent(c.lm, lp, vv, ff, true);
        const Jet2 qr = pos ? loss_tangent(c.lm, lp, vv, ff, false) : loss_full(c.lm, lp, vv, ff, false, MS_INTL_CLAMP);
        ptr = j3compose(qt * vv, v, FEL);
        prg = j3compose(qr * vv, v, FEL);
    } else if (c.cfg.lossKind == 1) {                        // constant efficiencies: PLtr = cT f v, PLrgb = -cR f v
        const Jet3 fv = FEL * v;
        ptr = c.P(P_CT, s) * fv;
        prg = (-c.P(P_CR, s)) * fv;
    } else { ptr = j3const(0.0); prg = j3const(0.0); }
}
// wrtDuration = false: jets w.r.t. (b_k, Fel_k, Fpb_k), the duration being tau(b_k, Fel_k + Fpb_k) -- the rows of the NLP.
// wrtDuration = true: (b_k, Fel_k, Fpb_k) held fixed, component 0 of the jets = derivative w.r.t. the duration (of this discrete
// map, not of the continuous integral) -- for the multipliers of the time rows in the reference's formulation (io.cuh).
MS_HD void loss_energy_rows(const Ctx& c, int s, const IntervalCoef& q, double b0, double b1, double fel, double fpb, const Jet2& tau, Jet3& etr, Jet3& erg,
                            bool wrtDuration = false) {
    const Jet3 B = wrtDuration ? j3const(b0) : j3var(b0, 0), FEL = wrtDuration ? j3const(fel) : j3var(fel, 1);
    const Jet3 F = c.cfg.withPn ? FEL + (wrtDuration ? j3const(fpb) : j3var(fpb, 2)) : FEL;
    const Jet3 e = wrtDuration ? j3var(tau.v, 0) : j3compose(tau, B, F);      // duration of the interval
    LossPar lp;
    if (c.cfg.lossKind == 2) lp = load_losspar(c, s);
    const bool pos = fel >= 0.0;
    Jet3 v = j3sqrt(B);
    etr = j3const(0.0); erg = j3const(0.0);
    int nsteps = MS_INTL_STEPS;
    if (c.cfg.lossKind == 2) {        // does the speed cross a kink of the map between the nodes (with a 3 % margin)?  (v is monotone in t)
        const double va = fmin(v.v, sqrt(b1)) * 0.97, vb = fmax(v.v, sqrt(b1)) * 1.03;
        const double kinks[3] = {c.lm.tv[0], lp.pMax / lp.fMax, c.lm.tv[c.lm.nv + 3]};
        for (int i = 0; i < 3; ++i) if (va < kinks[i] && kinks[i] < vb) nsteps = MS_INTL_STEPS_KINK;
    }
    const double h = 1.0 / nsteps, r0 = q.sr0 + q.c0;
    auto rhs = [&](const Jet3& vv, Jet3& dv, Jet3& dtr, Jet3& drg) {
        dv = e * (F - (q.sr1 * vv + q.sr2 * (vv * vv)) + (-r0));      // train.py:377: acceleration with b -> v^2
        Jet3 pt, pr;
        power_loss_jets(c, s, lp, pos, FEL, vv, pt, pr);
        dtr = e * pt; drg = e * pr;
    };
    for (int st = 0; st < nsteps; ++st) {
        Jet3 k1v, k1t, k1r, k2v, k2t, k2r, k3v, k3t, k3r, k4v, k4t, k4r;
        rhs(v, k1v, k1t, k1r);
        rhs(v + (0.5 * h) * k1v, k2v, k2t, k2r);
        rhs(v + (0.5 * h) * k2v, k3v, k3t, k3r);
        rhs(v + h * k3v, k4v, k4t, k4r);
        v = v + (h / 6.0) * (k1v + 2.0 * k2v + 2.0 * k3v + k4v);
        etr = etr + (h / 6.0) * (k1t + 2.0 * k2t + 2.0 * k3t + k4t);
        erg = erg + (h / 6.0) * (k1r + 2.0 * k2r + 2.0 * k3r + k4r);
    }
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
template <bool DYN, bool INTL = false, bool IRK = false>
MS_HD void cell_init(const Ctx& c, int k, int s) {
    const Config& g = c.cfg;
    const int N = c.I(SI_N_INT, s);
    if (s >= g.nInst || k > N) return;
    const int it = WS_IT0;
    for (int f = 0; f < IT_N; ++f) c.W(WS_IT0 + f, k, s) = 0.0;      // buffer 1 may hold the starting profile (read by neighbours)
    for (int f = 0; f < ST_N; ++f) c.W(WS_ST + f, k, s) = 0.0;
    Bnd B = load_bounds(c, k, s);
    double t = init_t(c, k, s, N), b = init_b(c, k, s, N);
    c.W(it + IT_T, k, s) = t;
    c.W(it + IT_B, k, s) = b;
    if (k >= 1 && k <= N) { c.W(it + IT_Z + Z_T_L, k, s) = 1.0; c.W(it + IT_Z + Z_T_U, k, s) = 1.0; }
    if (k >= 1 && k < N) { c.W(it + IT_Z + Z_B_L, k, s) = 1.0; c.W(it + IT_Z + Z_B_U, k, s) = 1.0; }
    if (k == N) return;
    double fel = push2(g.initMode ? c.W(WS_IT1 + IT_FEL, k, s) : 0.5, B.felL, B.felU);
    double fpb = g.withPn ? push2(g.initMode ? c.W(WS_IT1 + IT_FPB, k, s) : -0.1, B.fpbL, B.fpbU) : 0.0;
    IntervalCoef q = load_coef(c, k, s);
    // (integrateLosses: s_k is the loss energy of the interval, not a force -- the profile's guess is scaled by the interval length)
    double sl = push1(g.initMode ? c.W(WS_IT1 + IT_SL, k, s) * (INTL ? q.ds : 1.0) : 1.0, B.slL);
    c.W(it + IT_FEL, k, s) = fel;
    c.W(it + IT_FPB, k, s) = fpb;
    c.W(it + IT_SL, k, s) = sl;
    c.W(it + IT_Z + Z_FEL_L, k, s) = 1.0;
    c.W(it + IT_Z + Z_FEL_U, k, s) = 1.0;
    if (g.withPn) { c.W(it + IT_Z + Z_FPB_L, k, s) = 1.0; c.W(it + IT_Z + Z_FPB_U, k, s) = 1.0; }
    c.W(it + IT_Z + Z_SL_L, k, s) = 1.0;
    double d[NROW];
    ineq_values<DYN, INTL>(c, s, fel, fpb, sl, b, init_b(c, k + 1, s, N), q, d);
    if (INTL && g.energy) {
        Jet2 tau, phi;
        if (IRK) shoot_irk(jvar0(b), jvar1(fel + fpb), q, g.numSteps, g.numApprox, *c.irk, tau, phi);
        else shoot<Jet2>(jvar0(b), jvar1(fel + fpb), q, g.numSteps, g.numApprox, tau, phi);
        Jet3 etr, erg;
        loss_energy_rows(c, s, q, b, init_b(c, k + 1, s, N), fel, fpb, tau, etr, erg);
        d[R_LTR] = sl - etr.v; d[R_LRG] = sl - erg.v;
    }
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
    c.I(SI_LAST_GAIN, s) = 0;
    c.I(SI_FACT, s) = 0;
    c.I(SI_NACC, s) = 0;
    c.I(SI_ORIG, s) = s;
    c.I(SI_EXTRACTED, s) = 0;
}

// barrier sum of one cell: sum(log slack) is accumulated as the log of ONE product -- mantissa and exponent kept apart (the
// product of up to 6 slacks, each in [1e-20, 1e4], cannot leave the range; then it is renormalised), one log per cell
struct BarAcc {
    double slog, sdamp;
    bool ok;
    double prod;
    int nprod, expo;
};
MS_HD void bar_add(BarAcc& a, double slack, bool oneSided) {
    if (!(slack > 0.0)) a.ok = false;
    a.prod *= slack;
    if (++a.nprod == 6) { int e; a.prod = frexp(a.prod, &e); a.expo += e; a.nprod = 0; }
    if (oneSided) a.sdamp += slack;
}
MS_HD double bar_finish(BarAcc& a) {
    a.slog = log(a.prod) + 0.6931471805599453 * a.expo;
    return a.ok ? a.slog : NAN;
}

// ------------------------------------------------------------------------------------------------
// trial point                                                                 (IPOPT sec. 2.3, Alg. A step A-5)
// ------------------------------------------------------------------------------------------------
// x + alpha dx of one cell (all primal, slack and multiplier planes, kappa_sigma safeguard of eq. 16) formed in registers
// from the current iterate and the step; AO = the trial iterate of the cell, nb = what cell_eval needs of the neighbouring
// cells at the trial point (t, b of node k+1; Fel of intervals k-1 and k+1).  The trial point is not a kernel of its own: it
// is formed by the interval evaluation (cell_eval<DYN, true>), which evaluates the NLP there at once -- constraint violation,
// objective and barrier sums for the filter test AND the stage QP of the next iteration -- because the first trial point is
// accepted in ~95 % of the iterations; a rejected one costs one more evaluation at the halved step.
struct TrialNeighbours {
    double nT, nB, pFel, nFel;
};
MS_HD void form_trial_cell(const Ctx& c, int k, int s, int N, int cur, const Bnd& B, double* AO, TrialNeighbours& nb) {
    const Config& g = c.cfg;
    const int kn = (k < N) ? k + 1 : N, km = (k > 0) ? k - 1 : 0, kp = (k + 1 < N) ? k + 1 : k;
    double CI[IT_N], CS[ST_N];
    {
        const double* ip = &c.W(cur, k, s);
        const double* sp = &c.W(WS_ST, k, s);
#pragma unroll
        for (int f = 0; f < IT_N; ++f) CI[f] = ip[f * 32];
#pragma unroll
        for (int f = 0; f < ST_N; ++f) CS[f] = sp[f * 32];
    }
    const double nT = c.W(cur + IT_T, kn, s), nB = c.W(cur + IT_B, kn, s);
    const double nDT = c.W(WS_ST + ST_T, kn, s), nDB = c.W(WS_ST + ST_B, kn, s);
    const double pFel = c.W(cur + IT_FEL, km, s), pDFel = c.W(WS_ST + ST_FEL, km, s);
    const double nFel = c.W(cur + IT_FEL, kp, s), nDFel = c.W(WS_ST + ST_FEL, kp, s);
    const double al = c.D(SD_ALPHA, s), az = c.D(SD_ALPHA_Z, s), mu = c.D(SD_MU, s);
#pragma unroll
    for (int f = 0; f < IT_N; ++f) AO[f] = 0.0;
    // multiplier update of one bound: z + az*dz, dz = mu/s - z -+ (z/s) dv, then the kappa_sigma safeguard (eq. 16)
    auto zstep = [&](int zi, double sOld, double sNew, double dvSigned) {
        const double z = CI[IT_Z + zi];
        const double rOld = rcp_slack(sOld);
        const double dz = mu * rOld - z - (z * rOld) * dvSigned;
        double zn = z + az * dz;
        const double pr = zn * sNew;                 // kappa_sigma safeguard: mu/kappa <= z*slack <= kappa*mu
        if (pr > MS_KAPPA_SIGMA * mu) zn = MS_KAPPA_SIGMA * mu * rcp_slack(sNew);
        else if (pr < mu * (1.0 / MS_KAPPA_SIGMA)) zn = mu * (1.0 / MS_KAPPA_SIGMA) * rcp_slack(sNew);
        AO[IT_Z + zi] = zn;
    };
    auto var2 = [&](int itf, int stf, int zl, double L, double U, bool hasU) {
        const double v = CI[itf], dv = CS[stf];
        const double vn = v + al * dv;
        AO[itf] = vn;
        zstep(zl, v - L, vn - L, dv);
        if (hasU) zstep(zl + 1, U - v, U - vn, -dv);
    };
    if (k == 0) {
        AO[IT_T] = CI[IT_T]; AO[IT_B] = CI[IT_B];
    } else {
        var2(IT_T, ST_T, Z_T_L, B.tL, B.tU, true);
        if (k < N) var2(IT_B, ST_B, Z_B_L, B.bL, B.bU, true);
        else AO[IT_B] = CI[IT_B];
    }
    if (k < N) {
        var2(IT_FEL, ST_FEL, Z_FEL_L, B.felL, B.felU, true);
        if (g.withPn) var2(IT_FPB, ST_FPB, Z_FPB_L, B.fpbL, B.fpbU, true);
        var2(IT_SL, ST_SL, Z_SL_L, B.slL, 0.0, false);
        AO[IT_YT] = CI[IT_YT] + al * CS[ST_YT];
        AO[IT_YB] = CI[IT_YB] + al * CS[ST_YB];
#pragma unroll
        for (int j = 0; j < NROW; ++j) {
            if (!row_on(g, j)) continue;
            double L, U; bool hasU;
            row_bounds(B, j, L, U, hasU);
            const int zl = (j == R_P0) ? Z_P0_L : (j == R_P1) ? Z_P1_L : (j == R_ACC) ? Z_ACC_L : (j == R_LTR) ? Z_LTR_L : Z_LRG_L;
            const double w = CI[IT_W + j], dw = CS[ST_W + j];
            const double wn = w + al * dw;
            AO[IT_W + j] = wn;
            zstep(zl, w - L, wn - L, dw);
            if (hasU) zstep(zl + 1, U - w, U - wn, -dw);
            AO[IT_YD + j] = CI[IT_YD + j] + al * CS[ST_YD + j];
        }
    }
    nb.nT = nT + al * nDT;
    nb.nB = (k + 1 < N) ? nB + al * nDB : nB;
    nb.pFel = pFel + al * pDFel;
    nb.nFel = nFel + al * nDFel;
}

// ------------------------------------------------------------------------------------------------
// filter acceptance                                                         (IPOPT sec. 2.3, eqs. 18-20)
// ------------------------------------------------------------------------------------------------
// The per-instance reductions read their partial planes MS_RED_U intervals at a time (all loads of a group in flight, then
// summed in the original order: same bits as a one-by-one loop).
#define MS_RED_U 4
// IPOPT's second termination test with its default thresholds (acceptable_tol 1e-6, acceptable_constr_viol_tol 1e-2,
// acceptable_compl_inf_tol 1e-2, acceptable_dual_inf_tol 1e10) on the errors of the current iterate (set by inst_kkt); CasADi
// counts "Solved_To_Acceptable_Level" as success (reference ocp.py:364 reads stats()['success'])
#define MS_ACCEPTABLE_TOL 1e-6
#define MS_ACCEPTABLE_ITER 15
MS_HD bool acceptable_point(const Ctx& c, int s) {
    return c.D(SD_KKT, s) <= MS_ACCEPTABLE_TOL && c.D(SD_DINF, s) <= 1e10 && c.D(SD_PINF, s) <= 1e-2 && c.D(SD_CINF, s) <= 1e-2;
}

MS_HD void inst_decide(const Ctx& c, int s, const double* sums) {
    const Config& g = c.cfg;
    if (s >= g.nInst || c.I(SI_PHASE, s) != PH_TRIAL) return;
    const int N = c.I(SI_N_INT, s);
    const double tht = sums[0], ft = sums[1], slog = sums[2], sdamp = sums[3];
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
            // no restoration phase: report like IPOPT would -- which, at a point of "acceptable" quality, is success
            finish(c, s, acceptable_point(c, s) ? ST_ACCEPTABLE : ST_RESTORATION_FAILED);
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
// TRIAL = false: at the current iterate (first iteration).  TRIAL = true: at the trial point x + alpha dx, which is formed here
// and written to the other iterate buffer (see form_trial_cell); inst_decide then reads theta, the objective and the barrier sums
// from the same partial planes that inst_kkt uses when the point is accepted.
template <bool DYN, bool TRIAL, bool IRK = false, bool INTL = false>
MS_HD void cell_eval(const Ctx& c, int k, int s) {
    const Config& g = c.cfg;
    if (s >= g.nInst || c.I(SI_PHASE, s) != (TRIAL ? PH_TRIAL : PH_EVAL)) return;
    const int N = c.I(SI_N_INT, s);
    if (k > N) return;
    const int it = c.I(SI_PARITY, s) ? WS_IT1 : WS_IT0;
    // ---- all loads first (independent, in flight together); the stores of the stage QP are at the very end
    const int kn = (k < N) ? k + 1 : N, km = (k > 0) ? k - 1 : 0;
    double CI[IT_N];
    Bnd B = load_bounds(c, k, s);
    double nT, nB, pFel, nFel;
    if (TRIAL) {
        TrialNeighbours nb;
        form_trial_cell(c, k, s, N, it, B, CI, nb);
        nT = nb.nT; nB = nb.nB; pFel = nb.pFel; nFel = nb.nFel;
        double* op = &c.W(c.I(SI_PARITY, s) ? WS_IT0 : WS_IT1, k, s);
#pragma unroll
        for (int f = 0; f < IT_N; ++f) op[f * 32] = CI[f];
    } else {
        const double* ip = &c.W(it, k, s);
#pragma unroll
        for (int f = 0; f < IT_N; ++f) CI[f] = ip[f * 32];
        nT = c.W(it + IT_T, kn, s); nB = c.W(it + IT_B, kn, s);
        pFel = c.W(it + IT_FEL, km, s); nFel = c.W(it + IT_FEL, (k + 1 < N) ? k + 1 : k, s);
    }
    const double iscale = rcp_slack(c.P(P_SCALE, s));          // objective scaling: multiplications instead of divisions below
    double H[28], g0[NV7], g1[NV7];
    #pragma unroll
    for (int i = 0; i < 28; ++i) H[i] = 0.0;
    #pragma unroll
    for (int i = 0; i < NV7; ++i) { g0[i] = 0.0; g1[i] = 0.0; }
    BarAcc bar{0.0, 0.0, true, 1.0, 0, 0};
    double cmin = 1e300, cmax = 0.0, zsum = 0.0;

    auto bound = [&](int vi, int zi, double slack, double sign, bool oneSided) {
        // sign = +1 lower bound (slack = v-L), -1 upper bound (slack = U-v)
        const double z = CI[IT_Z + zi];
        const double r = rcp_slack(slack);
        H[sidx(vi, vi)] += z * r;
        g1[vi] += -sign * r + (oneSided ? MS_KAPPA_D : 0.0);
        double pr = z * slack;
        cmin = fmin(cmin, pr); cmax = fmax(cmax, pr); zsum += z;
        bar_add(bar, slack, oneSided);
    };

    const double t = CI[IT_T], b = CI[IT_B];
    double own_t = 0.0, own_b = 0.0;
    if (k >= 1) {
        bound(V_T, Z_T_L, t - B.tL, 1.0, false);
        bound(V_T, Z_T_U, B.tU - t, -1.0, false);
        own_t = -CI[IT_Z + Z_T_L] + CI[IT_Z + Z_T_U];
        if (k < N) {
            bound(V_B, Z_B_L, b - B.bL, 1.0, false);
            bound(V_B, Z_B_U, B.bU - b, -1.0, false);
            own_b = -CI[IT_Z + Z_B_L] + CI[IT_Z + Z_B_U];
        }
    }
    if (k == N) {
        // terminal node: only t_N carries a barrier; stored in the same QP planes for the Riccati start
        double fo = 0.0;
        if (!g.energy) { own_t += iscale; fo = t * iscale; }
        c.W(WS_QP + QP_H_TT, k, s) = H[sidx(V_T, V_T)];
        c.W(WS_QP + QP_G1_T, k, s) = g1[V_T];
        c.W(WS_PART + PC_TH, k, s) = 0.0;
        c.W(WS_PART + PC_F, k, s) = fo;
        c.W(WS_PART + PC_SLOG, k, s) = bar_finish(bar);
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
    const double fel = CI[IT_FEL], fpb = CI[IT_FPB], sl = CI[IT_SL];
    const double t1 = nT, b1 = nB;
    const double yt = CI[IT_YT], yb = CI[IT_YB];
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
    if (IRK) shoot_irk(jvar0(b), jvar1(fel + fpb), q, g.numSteps, g.numApprox, *c.irk, tau, phi);
    else shoot<Jet2>(jvar0(b), jvar1(fel + fpb), q, g.numSteps, g.numApprox, tau, phi);
    const double ct = t1 - t - tau.v, cb = b1 - phi.v;
    double v0, v1, iv0, iv1;
    sqrt_inv(b, v0, iv0);
    sqrt_inv(b1, v1, iv1);
    const double ib0 = iv0 * iv0, ib1 = iv1 * iv1;
    const double a_b = -(0.5 * q.sr1 * iv0 + q.sr2), a_bb = 0.25 * q.sr1 * ib0 * iv0;

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
    #pragma unroll
    for (int j = 0; j < NROW; ++j) for (int i = 0; i < NV7; ++i) J[j][i] = 0.0;
    ineq_values<false, INTL>(c, s, fel, fpb, sl, b, b1, q, d);
    double ydv[NROW];
    #pragma unroll
    for (int j = 0; j < NROW; ++j) ydv[j] = CI[IT_YD + j];
    J[R_P0][V_B] = 0.5 * fel * iv0; J[R_P0][V_FEL] = v0;
    J[R_P1][V_BN] = 0.5 * fel * iv1; J[R_P1][V_FEL] = v1;
    J[R_ACC][V_B] = a_b; J[R_ACC][V_FEL] = 1.0; J[R_ACC][V_FPB] = g.withPn ? 1.0 : 0.0;
    J[R_LTR][V_FEL] = -c.P(P_CT, s); J[R_LTR][V_SL] = 1.0;
    J[R_LRG][V_FEL] = c.P(P_CR, s); J[R_LRG][V_SL] = 1.0;
    if (INTL && g.energy) {
        // integrateLosses: rows s - E(b_k, Fel_k, Fpb_k) with E the loss energies integrated over the interval in the time domain
        Jet3 en[2];
        loss_energy_rows(c, s, q, b, b1, fel, fpb, tau, en[0], en[1]);
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            const int row = (a == 0) ? R_LTR : R_LRG;
            d[row] = sl - en[a].v;
            J[row][V_B] = -en[a].g[0]; J[row][V_FEL] = -en[a].g[1]; J[row][V_FPB] = g.withPn ? -en[a].g[2] : 0.0; J[row][V_BN] = 0.0;
            const double y = ydv[row];
            H[sidx(V_B, V_B)] -= y * en[a].h[0];
            H[sidx(V_B, V_FEL)] -= y * en[a].h[1];
            H[sidx(V_FEL, V_FEL)] -= y * en[a].h[3];
            if (g.withPn) { H[sidx(V_B, V_FPB)] -= y * en[a].h[2]; H[sidx(V_FEL, V_FPB)] -= y * en[a].h[4]; H[sidx(V_FPB, V_FPB)] -= y * en[a].h[5]; }
        }
    } else if (DYN && g.energy) {
        // loss map of efficiency.py: rows s - G(Fel, b_k, b_{k+1}) with full first and second derivatives
        LossRow lr[2];
        loss_rows_dynamic(c.lm, load_losspar(c, s), fel, b, b1, lr[0], lr[1]);
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            const int row = (a == 0) ? R_LTR : R_LRG;
            d[row] = sl - lr[a].v;
            J[row][V_FEL] = -lr[a].gF; J[row][V_B] = -lr[a].g0; J[row][V_BN] = -lr[a].g1;
            const double y = ydv[row];
            H[sidx(V_FEL, V_FEL)] -= y * lr[a].hFF;
            H[sidx(V_B, V_FEL)] -= y * lr[a].hF0;
            H[sidx(V_FEL, V_BN)] -= y * lr[a].hF1;
            H[sidx(V_B, V_B)] -= y * lr[a].h00;
            H[sidx(V_B, V_BN)] -= y * lr[a].h01;
            H[sidx(V_BN, V_BN)] -= y * lr[a].h11;
        }
    }
    if (g.withPower) {
        H[sidx(V_B, V_B)] += ydv[R_P0] * (-0.25 * fel * ib0 * iv0);
        H[sidx(V_B, V_FEL)] += ydv[R_P0] * (0.5 * iv0);
        H[sidx(V_BN, V_BN)] += ydv[R_P1] * (-0.25 * fel * ib1 * iv1);
        H[sidx(V_FEL, V_BN)] += ydv[R_P1] * (0.5 * iv1);
    }
    H[sidx(V_B, V_B)] += ydv[R_ACC] * a_bb;

    // ---- objective                                                           (ocp.py:146-154,223,243-245)
    double fo = 0.0, gf_fel = 0.0, gf_fpb = 0.0, gf_sl = 0.0;
    if (g.energy) {
        const double w2 = 2e-3 * iscale, dsc = q.ds * iscale, ssc = INTL ? iscale : dsc;      // ocp.py:223 / :235
        fo = dsc * fel + ssc * sl;
        gf_fel = dsc; gf_sl = ssc;
        g0[V_FEL] += dsc; g0[V_SL] += ssc;
        if (k >= 1) {
            double df = fel - pFel;
            fo += 1e-3 * df * df * iscale;
            H[sidx(V_FEL, V_FEL)] += w2; H[sidx(V_F, V_F)] += w2; H[sidx(V_F, V_FEL)] -= w2;
            g0[V_FEL] += w2 * df; g0[V_F] -= w2 * df;
            gf_fel += w2 * df;
        }
        if (k + 1 < N) gf_fel -= w2 * (nFel - fel);
    } else {
        const double w4 = 2e-4 * iscale;
        fo = 1e-4 * (fel * fel + fpb * fpb) * iscale;
        H[sidx(V_FEL, V_FEL)] += w4; g0[V_FEL] += w4 * fel; gf_fel = w4 * fel;
        if (g.withPn) { H[sidx(V_FPB, V_FPB)] += w4; g0[V_FPB] += w4 * fpb; gf_fpb = w4 * fpb; }
    }

    // ---- condensation of the inequality rows (slack w, multiplier v_L/v_U)
    double th = fabs(ct) + fabs(cb), pinf = fmax(fabs(ct), fabs(cb));
    double ysum = fabs(yt) + fabs(yb), dinf = 0.0;
    double rx_fel = gf_fel - tau.g1 * yt - phi.g1 * yb - CI[IT_Z + Z_FEL_L] + CI[IT_Z + Z_FEL_U];
    double rx_fpb = gf_fpb - tau.g1 * yt - phi.g1 * yb - CI[IT_Z + Z_FPB_L] + CI[IT_Z + Z_FPB_U];
    double rx_sl = gf_sl - CI[IT_Z + Z_SL_L];
    own_b += -tau.g0 * yt - phi.g0 * yb;
    own_t += -yt;
    double cn_b = yb;
    #pragma unroll
    for (int j = 0; j < NROW; ++j) {
        if (DYN) c.W(WS_QP + QP_RES + j, k, s) = 0.0;
        if (!row_on(g, j)) continue;
        double L, U; bool hasU;
        row_bounds(B, j, L, U, hasU);
        const int zl = (j == R_P0) ? Z_P0_L : (j == R_P1) ? Z_P1_L : (j == R_ACC) ? Z_ACC_L : (j == R_LTR) ? Z_LTR_L : Z_LRG_L;
        const double w = CI[IT_W + j];
        const double vL = CI[IT_Z + zl], sL = w - L;
        const double rL = rcp_slack(sL);
        double sig = vL * rL, coef = -rL + (hasU ? 0.0 : MS_KAPPA_D);
        double rw = -ydv[j] - vL;
        double pr = vL * sL;
        cmin = fmin(cmin, pr); cmax = fmax(cmax, pr); zsum += vL;
        bar_add(bar, sL, !hasU);
        if (hasU) {
            const double vU = CI[IT_Z + zl + 1], sU = U - w;
            const double rU = rcp_slack(sU);
            sig += vU * rU; coef += rU; rw += vU;
            pr = vU * sU;
            cmin = fmin(cmin, pr); cmax = fmax(cmax, pr); zsum += vU;
            bar_add(bar, sU, false);
        }
        const double res = sub_rn(d[j], w);
        if (DYN) c.W(WS_QP + QP_RES + j, k, s) = res;      // otherwise cell_step recomputes it (bit for bit)
        th += fabs(res); pinf = fmax(pinf, fabs(res)); ysum += fabs(ydv[j]);
        dinf = fmax(dinf, fabs(rw));
        #pragma unroll
        for (int a = 0; a < NV7; ++a) {
            if (J[j][a] == 0.0) continue;
            g0[a] += sig * res * J[j][a];
            g1[a] += coef * J[j][a];
            #pragma unroll
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
    #pragma unroll
    for (int i = 0; i < 6; ++i) hc[i] = H[sidx(i, V_BN)];
    const double hpp = H[sidx(V_BN, V_BN)], gp0 = g0[V_BN], gp1 = g1[V_BN];
    if (k + 1 < N) {
        #pragma unroll
        for (int i = 0; i < 6; ++i) {
            #pragma unroll
            for (int j = i; j < 6; ++j) H[sidx(i, j)] += hc[i] * av[j] + av[i] * hc[j] + hpp * av[i] * av[j];
            g0[i] += hc[i] * rb + (hpp * rb + gp0) * av[i];
            g1[i] += gp1 * av[i];
        }
    }
    // ---- static condensation of s (its pivot H_ss is a sum of barrier terms, > 0, and independent of the value function):
    // the sweep works on the Schur complement; the column of s is stored next to it for cell_step and the delta_w != 0 path
    const double hb = H[sidx(V_B, V_SL)], hF = H[sidx(V_FEL, V_SL)], hQ = H[sidx(V_FPB, V_SL)], hss = H[sidx(V_SL, V_SL)];
    {
        const double is = rcp(hss);
        const double eb = hb * is, eF = hF * is, eQ = hQ * is;
        H[sidx(V_B, V_B)] -= eb * hb; H[sidx(V_B, V_FEL)] -= eb * hF; H[sidx(V_B, V_FPB)] -= eb * hQ;
        H[sidx(V_FEL, V_FEL)] -= eF * hF; H[sidx(V_FEL, V_FPB)] -= eF * hQ; H[sidx(V_FPB, V_FPB)] -= eQ * hQ;
        g0[V_B] -= eb * g0[V_SL]; g0[V_FEL] -= eF * g0[V_SL]; g0[V_FPB] -= eQ * g0[V_SL];
        g1[V_B] -= eb * g1[V_SL]; g1[V_FEL] -= eF * g1[V_SL]; g1[V_FPB] -= eQ * g1[V_SL];
    }
    // ---- store
    c.W(WS_QP + QP_H_TT, k, s) = H[sidx(V_T, V_T)];
    c.W(WS_QP + QP_H_BB, k, s) = H[sidx(V_B, V_B)];
    c.W(WS_QP + QP_H_BFEL, k, s) = H[sidx(V_B, V_FEL)];
    c.W(WS_QP + QP_H_BFPB, k, s) = H[sidx(V_B, V_FPB)];
    c.W(WS_QP + QP_H_BSL, k, s) = hb;
    c.W(WS_QP + QP_H_FF, k, s) = H[sidx(V_F, V_F)];
    c.W(WS_QP + QP_H_FFEL, k, s) = H[sidx(V_F, V_FEL)];
    c.W(WS_QP + QP_H_FELFEL, k, s) = H[sidx(V_FEL, V_FEL)];
    c.W(WS_QP + QP_H_FELFPB, k, s) = H[sidx(V_FEL, V_FPB)];
    c.W(WS_QP + QP_H_FELSL, k, s) = hF;
    c.W(WS_QP + QP_H_FPBFPB, k, s) = H[sidx(V_FPB, V_FPB)];
    c.W(WS_QP + QP_H_FPBSL, k, s) = hQ;
    c.W(WS_QP + QP_H_SLSL, k, s) = hss;
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
    if (DYN) {      // constant-efficiency rows: cell_step recomputes the row gradients
    c.W(WS_QP + QP_J_P0_B, k, s) = J[R_P0][V_B];
    c.W(WS_QP + QP_J_P0_FEL, k, s) = J[R_P0][V_FEL];
    c.W(WS_QP + QP_J_P1_FEL, k, s) = J[R_P1][V_FEL];
    c.W(WS_QP + QP_J_P1_BN, k, s) = J[R_P1][V_BN];
    c.W(WS_QP + QP_J_ACC_B, k, s) = J[R_ACC][V_B];
    c.W(WS_QP + QP_J_LTR_FEL, k, s) = J[R_LTR][V_FEL];
    c.W(WS_QP + QP_J_LTR_B, k, s) = J[R_LTR][V_B];
    c.W(WS_QP + QP_J_LTR_BN, k, s) = INTL ? J[R_LTR][V_FPB] : J[R_LTR][V_BN];      // integrated losses: the rows depend on Fpb_k, not on b_{k+1}
    c.W(WS_QP + QP_J_LRG_FEL, k, s) = J[R_LRG][V_FEL];
    c.W(WS_QP + QP_J_LRG_B, k, s) = J[R_LRG][V_B];
    c.W(WS_QP + QP_J_LRG_BN, k, s) = INTL ? J[R_LRG][V_FPB] : J[R_LRG][V_BN];
    }
    c.W(WS_PART + PC_TH, k, s) = th;
    c.W(WS_PART + PC_F, k, s) = fo;
    c.W(WS_PART + PC_SLOG, k, s) = bar_finish(bar);
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


}  // namespace mseetc
