// TEST HARNESS ONLY -- not part of the product, never loaded by the ms-eetc_b200 package.
//
// Compiles the solver's host/device headers (ms-eetc_b200/csrc/*.cuh) with g++ and runs the identical
// lock-step tick sequence with plain loops in place of CUDA thread grids, so that the algorithmic logic
// (interval sensitivities, stage-QP condensation, Riccati recursion, filter line search) can be unit-tested
// against the oracle on a machine without a GPU.  The C-ABI library (libmseetc_b200.so) contains no such
// path: it launches CUDA kernels or fails.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../ms-eetc_b200/csrc/io.cuh"

using namespace mseetc;


static void lin_residual(const Ctx& c, int s, int N, const char* tag) {
    const double mu = c.D(SD_MU, s), pn = c.cfg.withPn ? 1.0 : 0.0;
    double worstU = 0, worstX = 0, worstD = 0; int kU = -1, kX = -1, kD = -1;
    for (int k = 0; k < N - 1; ++k) {
        double v[QP_SWEEP_N];
        for (int f = 0; f < QP_SWEEP_N; ++f) v[f] = c.W(WS_QP + f, k, s);
        const double dt = c.W(WS_ST + ST_T, k, s), db = c.W(WS_ST + ST_B, k, s), df = (k > 0) ? c.W(WS_ST + ST_FEL, k - 1, s) : 0.0;
        const double dF = c.W(WS_ST + ST_FEL, k, s), dQ = c.W(WS_ST + ST_FPB, k, s);
        const double dtn = c.W(WS_ST + ST_T, k + 1, s), dbn = c.W(WS_ST + ST_B, k + 1, s);
        double Pn[6], pnx[3], Pk[6], pk[3];
        for (int i = 0; i < 6; ++i) { Pn[i] = c.W(WS_RIC + RIC_P + i, k + 1, s); Pk[i] = c.W(WS_RIC + RIC_P + i, k, s); }
        for (int i = 0; i < 3; ++i) { pnx[i] = c.W(WS_RIC + RIC_PV + i, k + 1, s); pk[i] = c.W(WS_RIC + RIC_PV + i, k, s); }
        const double lt = pnx[0] + Pn[0] * dtn + Pn[1] * dbn + Pn[2] * dF;
        const double lb = pnx[1] + Pn[1] * dtn + Pn[3] * dbn + Pn[4] * dF;
        const double lf = pnx[2] + Pn[2] * dtn + Pn[4] * dbn + Pn[5] * dF;
        const double tb = v[QP_TAU_B], tF = v[QP_TAU_F], pb = v[QP_PHI_B], pF = v[QP_PHI_F];
        const double rF = v[QP_H_BFEL] * db + v[QP_H_FFEL] * df + v[QP_H_FELFEL] * dF + v[QP_H_FELFPB] * dQ + v[QP_G0_FEL] + mu * v[QP_G1_FEL] + tF * lt + pF * lb + lf;
        const double rQ = v[QP_H_BFPB] * db + v[QP_H_FELFPB] * dF + v[QP_H_FPBFPB] * dQ + v[QP_G0_FPB] + mu * v[QP_G1_FPB] + pn * (tF * lt + pF * lb);
        const double ru = fmax(fabs(rF), pn * fabs(rQ));
        if (ru > worstU) { worstU = ru; kU = k; }
        // costate recursion: lambda_k = H_x. d + g_x + A' lambda_{k+1}  vs  P_k dx_k + p_k
        const double l0 = pk[0] + Pk[0] * dt + Pk[1] * db + Pk[2] * df;
        const double l1 = pk[1] + Pk[1] * dt + Pk[3] * db + Pk[4] * df;
        const double l2 = pk[2] + Pk[2] * dt + Pk[4] * db + Pk[5] * df;
        const double x0 = v[QP_H_TT] * dt + mu * v[QP_G1_T] + lt;
        const double x1 = v[QP_H_BB] * db + v[QP_H_BFEL] * dF + v[QP_H_BFPB] * dQ + v[QP_G0_B] + mu * v[QP_G1_B] + tb * lt + pb * lb;
        const double x2 = v[QP_H_FF] * df + v[QP_H_FFEL] * dF + v[QP_G0_F];
        const double rx = fmax(fmax(fabs(l0 - x0), fabs(l1 - x1)), fabs(l2 - x2));
        if (rx > worstX) { worstX = rx; kX = k; }
        const double rd = fmax(fabs(dtn - (dt + tb * db + tF * (dF + pn * dQ) + v[QP_RT])), fabs(dbn - (pb * db + pF * (dF + pn * dQ) + v[QP_RB])));
        if (rd > worstD) { worstD = rd; kD = k; }
    }
    printf("   lin-res %s: controls %.2e at %d | costate %.2e at %d | dynamics %.2e at %d\n", tag, worstU, kU, worstX, kX, worstD, kD);
}

static IrkTab g_irk;
static bool g_irk_on = false;
static int g_int_losses = 0;

extern "C" {

struct hostsim_problem {
    int32_t n_intervals_max, with_pn_brake, with_power_rows, energy_optimal, loss_kind, num_steps, num_approx_steps,
        max_iterations;
    double tol, mu_init;
};

int hostsim_solve_batch(const hostsim_problem* pr, int32_t n, const double* params, const int32_t* nint,
                        const int32_t* trk_of, const int32_t* trk_off, const double* ds, const double* c0,
                        const double* bmax, const double* tmin, double* z_out, double* lam_out, double* obj, double* kkt, int32_t* iters,
                        int32_t* status, int32_t verbose_inst, int32_t* ticks_out, int32_t pit_lanes,
                        int32_t lm_nl, int32_t lm_nv, const double* lm_tl, const double* lm_tv, const double* lm_coef, int32_t init_mode) {
    Config g;
    memset(&g, 0, sizeof g);
    g.S = pad_slots(n);
    g.NK = pr->n_intervals_max + 1;
    g.nInst = n;
    g.withPn = pr->with_pn_brake; g.withPower = pr->with_power_rows; g.energy = pr->energy_optimal;
    g.lossKind = pr->loss_kind; g.numSteps = pr->num_steps; g.numApprox = pr->num_approx_steps;
    g.maxIter = pr->max_iterations; g.tol = pr->tol; g.muInit = pr->mu_init; g.initMode = init_mode; g.intLosses = (g_int_losses && pr->energy_optimal) ? 1 : 0;
    WsPlan plan = plan_workspace(g.S, g.NK);
    // the device workspace is uninitialised memory: poison it here (0xFF bytes = NaN doubles, -1 ints) so that a read of a plane
    // nobody has written shows up in the emulation too; the library clears the integer state and the counters itself
    std::vector<char> buf(plan.total, (char)0xFF);
    std::memset(buf.data() + plan.off_si, 0, sizeof(int) * (size_t)SI_N * g.S);
    std::memset(buf.data() + plan.off_done, 0, 256);
    Ctx c;
    c.cfg = g;
    c.ws = (double*)(buf.data() + plan.off_ws);
    c.par = (double*)(buf.data() + plan.off_par);
    c.sd = (double*)(buf.data() + plan.off_sd);
    c.si = (int*)(buf.data() + plan.off_si);
    c.done = (int*)(buf.data() + plan.off_done);
    c.cnt = (unsigned long long*)(buf.data() + plan.off_done + 64);
    c.tmin = tmin;
    c.plan = nullptr;
    c.irk = g_irk_on ? &g_irk : nullptr;
    c.lm.tl = lm_tl; c.lm.tv = lm_tv; c.lm.coef = lm_coef; c.lm.nl = lm_nl; c.lm.nv = lm_nv;
    const bool dyn = (pr->loss_kind == 2 && pr->energy_optimal);
    BatchIO io{params, nint, trk_of, trk_off, ds, c0, bmax, tmin, z_out, lam_out, obj, kkt, iters, status};
    for (int s = 0; s < g.S; ++s) inst_setup(c, io, s);
    for (int k = 0; k < g.NK; ++k) for (int s = 0; s < g.S; ++s) cell_setup(c, io, k, s);
    for (int s = 0; s < g.S; ++s) { inst_screen(c, s); inst_profile(c, s); }
    const bool intl = g.intLosses != 0;
    for (int k = 0; k < g.NK; ++k) for (int s = 0; s < g.S; ++s) { if (intl) { if (c.irk) cell_init<true, true, true>(c, k, s); else cell_init<true, true>(c, k, s); } else if (dyn) cell_init<true>(c, k, s); else cell_init<false>(c, k, s); }
    int tick = 0;
    long fallbacks = 0;
    const int diag = getenv("HOSTSIM_PIT_DIAG") ? atoi(getenv("HOSTSIM_PIT_DIAG")) : 0;
    double diag_worst = 0.0;
    const int pit_until = getenv("HOSTSIM_PIT_UNTIL") ? atoi(getenv("HOSTSIM_PIT_UNTIL")) : 1 << 30;
    const int maxTicks = 20 * pr->max_iterations + 50;
    auto report = [&](const char* tag) {
        int s = verbose_inst;
        if (s < 0 || s >= n) return;
        printf("%s tick %3d it %3d ph %d f=%.10e th=%.3e dinf=%.2e pinf=%.2e mu=%.1e a=%.3e az=%.3e kkt=%.2e nls=%d nreg=%d\n", tag,
               tick, c.I(SI_ITERS, s), c.I(SI_PHASE, s), c.D(SD_FOBJ, s), c.D(SD_THETA, s), c.D(SD_DINF, s), c.D(SD_PINF, s),
               c.D(SD_MU, s), c.D(SD_ALPHA, s), c.D(SD_ALPHA_Z, s), c.D(SD_KKT, s), c.I(SI_NLS, s), c.I(SI_NREG, s));
    };
    // the same lock-step sequence as mseetc_solve_batch: starting point, then per tick the direction and the evaluation at the trial point
    auto reduce_kkt = [&](bool trial) {
        for (int s = 0; s < g.nInst; ++s) {
            if (c.I(SI_PHASE, s) != (trial ? PH_TRIAL : PH_EVAL)) continue;
            const int N = c.I(SI_N_INT, s), it = ((c.I(SI_PARITY, s) != 0) != trial) ? WS_IT1 : WS_IT0;
            KktAcc tot, part;
            kkt_init(tot);
            for (int w = 0; w < RED_W; ++w) { kkt_partials(c, s, N, it, w, RED_W, part); kkt_combine(tot, part); }
            if (trial) { const double sums[4] = {tot.th, tot.fo, tot.slog, tot.sdamp}; inst_decide(c, s, sums); }
            inst_kkt(c, s, tot);
            if (diag >= 3 && s == verbose_inst && c.I(SI_PHASE, s) != PH_TRIAL) {
                const int itc = c.I(SI_PARITY, s) ? WS_IT1 : WS_IT0;
                double best[3] = {0, 0, 0}; int at[3] = {-1, -1, -1};
                for (int k = 0; k <= N; ++k) {
                    const double d0 = c.W(WS_PART + PC_DINF, k, s);
                    const double d1 = (k >= 1) ? fabs(c.W(WS_PART + PC_OWN_T, k, s) + c.W(itc + IT_YT, k - 1, s)) : 0.0;
                    const double d2 = (k >= 1 && k < N) ? fabs(c.W(WS_PART + PC_OWN_B, k, s) + c.W(WS_PART + PC_CN_B, k - 1, s)) : 0.0;
                    if (d0 > best[0]) { best[0] = d0; at[0] = k; }
                    if (d1 > best[1]) { best[1] = d1; at[1] = k; }
                    if (d2 > best[2]) { best[2] = d2; at[2] = k; }
                }
                printf("   dinf parts: controls %.2e at %d | t-node %.2e at %d | b-node %.2e at %d\n", best[0], at[0], best[1], at[1], best[2], at[2]);
            }
        }
    };
    for (int k = 0; k < g.NK; ++k) for (int s = 0; s < g.S; ++s) {
        if (intl) { if (c.irk) cell_eval<true, false, true, true>(c, k, s); else cell_eval<true, false, false, true>(c, k, s); }
        else if (c.irk) { if (dyn) cell_eval<true, false, true>(c, k, s); else cell_eval<false, false, true>(c, k, s); }
        else if (dyn) cell_eval<true, false>(c, k, s); else cell_eval<false, false>(c, k, s);
    }
    reduce_kkt(false);
    for (;;) {
        for (int s = 0; s < g.S; ++s) {
            DirectFetch<BwdFields> fb; DirectFetch<FwdFields> ff;
            if (pit_lanes > 1 && diag && s < g.nInst && c.I(SI_PHASE, s) == PH_FACTOR) {
                // diagnostic: direction of the chunked sweeps next to the one of the sequential sweeps (same iterate)
                const int N = c.I(SI_N_INT, s);
                inst_step(c, s, fb, ff);
                std::vector<double> ref(4 * (N + 1));
                for (int k = 0; k <= N; ++k) { ref[4*k] = c.W(WS_ST+ST_FEL,k,s); ref[4*k+1] = c.W(WS_ST+ST_FPB,k,s); ref[4*k+2] = c.W(WS_ST+ST_T,k,s); ref[4*k+3] = c.W(WS_ST+ST_B,k,s); }
                std::vector<double> refP(9 * (N + 1));
                for (int k = 0; k <= N; ++k) for (int i = 0; i < 9; ++i) refP[9*k+i] = c.W(WS_RIC + RIC_P + i, k, s);
                if (getenv("HOSTSIM_DUMP") && c.I(SI_ITERS, s) == atoi(getenv("HOSTSIM_DUMP"))) {
                    FILE* f = fopen("/tmp/exp/qp_dump.txt", "w");
                    fprintf(f, "%d %.17g %d\n", N, c.D(SD_MU, s), c.cfg.withPn);
                    for (int k = 0; k <= N; ++k) {
                        for (int q = 0; q < QP_SWEEP_N + 6; ++q) fprintf(f, "%.17g ", c.W(WS_QP + q, k, s));
                        for (int q = 0; q < 9; ++q) fprintf(f, "%.17g ", c.W(WS_RIC + RIC_P + q, k, s));
                        fprintf(f, "\n");
                    }
                    fclose(f);
                }
                if (c.I(SI_PHASE, s) == PH_STEPPED) {
                    if (diag >= 3) lin_residual(c, s, N, "seq");
                    c.I(SI_PHASE, s) = PH_FACTOR;
                    inst_step_pit_emulated(c, s, pit_lanes, fb, ff, &fallbacks);
                    double num[4] = {0,0,0,0}, den[4] = {1e-300,1e-300,1e-300,1e-300};
                    const int fld[4] = {ST_FEL, ST_FPB, ST_T, ST_B};
                    for (int k = 0; k <= N; ++k) for (int i = 0; i < 4; ++i) { num[i] = fmax(num[i], fabs(c.W(WS_ST+fld[i],k,s) - ref[4*k+i])); den[i] = fmax(den[i], fabs(ref[4*k+i])); }
                    double pd = 0.0;
                    for (int k = 0; k <= N; ++k) { double sc = 1e-300, dv = 0; for (int i = 0; i < 6; ++i) { sc = fmax(sc, fabs(refP[9*k+i])); dv = fmax(dv, fabs(c.W(WS_RIC+RIC_P+i,k,s) - refP[9*k+i])); } pd = fmax(pd, dv / sc); }
                    double worst = 0; for (int i = 0; i < 4; ++i) worst = fmax(worst, num[i] / den[i]);
                    if (worst > diag_worst) diag_worst = worst;
                    if (diag >= 3) {
                        lin_residual(c, s, N, "pit");
                        for (int i = 0; i < 4; ++i) {
                            int kk = 0; double w = -1;
                            for (int k = 0; k <= N; ++k) { const double r = fabs(c.W(WS_ST+fld[i],k,s) - ref[4*k+i]) / (fabs(ref[4*k+i]) + 1e-300); if (fabs(ref[4*k+i]) > 0 && r > w) { w = r; kk = k; } }
                            printf("   worst pointwise rel dev field %d: k=%d seq %.6e pit %.6e\n", i, kk, ref[4*kk+i], c.W(WS_ST+fld[i],kk,s));
                        }
                    }
                    if (diag > 1) printf("diag s %d it %d mu %.1e: rel step dev Fel %.1e Fpb %.1e t %.1e b %.1e   P dev %.1e\n", s, c.I(SI_ITERS, s), c.D(SD_MU, s), num[0]/den[0], num[1]/den[1], num[2]/den[2], num[3]/den[3], pd);
                }
                continue;
            }
            if (pit_lanes > 1 && (s >= g.nInst || c.I(SI_ITERS, s) < pit_until)) inst_step_pit_emulated(c, s, pit_lanes, fb, ff, &fallbacks);   // chunked parallel-in-time variant
            else inst_step(c, s, fb, ff);
        }
        for (int k = 0; k < g.NK; ++k) for (int s = 0; s < g.S; ++s) { if (intl) cell_step<true, true>(c, k, s); else if (dyn) cell_step<true>(c, k, s); else cell_step<false>(c, k, s); }
        for (int s = 0; s < g.nInst; ++s) {
            if (c.I(SI_PHASE, s) != PH_STEPPED) continue;
            const int N = c.I(SI_N_INT, s);
            double tot[3] = {1.0, 1.0, 0.0}, part[3];
            for (int w = 0; w < RED_W; ++w) { alpha_partials(c, s, N, w, RED_W, part); tot[0] = fmin(tot[0], part[0]); tot[1] = fmin(tot[1], part[1]); tot[2] += part[2]; }
            inst_alpha(c, s, tot);
        }
        report("step ");
        if (*c.done >= n || tick >= maxTicks) break;
        for (int k = 0; k < g.NK; ++k) for (int s = 0; s < g.S; ++s) {
            if (intl) { if (c.irk) cell_eval<true, true, true, true>(c, k, s); else cell_eval<true, true, false, true>(c, k, s); }
            else if (c.irk) { if (dyn) cell_eval<true, true, true>(c, k, s); else cell_eval<false, true, true>(c, k, s); }
            else if (dyn) cell_eval<true, true>(c, k, s); else cell_eval<false, true>(c, k, s);
        }
        reduce_kkt(true);
        ++tick;
        if (*c.done >= n) break;
    }
    for (int k = 0; k < g.NK; ++k) for (int s = 0; s < g.S; ++s) cell_extract(c, io, k, s);
    if (intl) for (int k = 0; k < g.NK; ++k) for (int s = 0; s < g.S; ++s) { if (c.irk) cell_fix_time_multiplier_intl<true>(c, io, k, s); else cell_fix_time_multiplier_intl<false>(c, io, k, s); }
    if (ticks_out) *ticks_out = tick;
    if (pit_lanes > 1 && getenv("HOSTSIM_PIT_DIAG")) printf("pit: fallbacks %ld  worst relative step deviation %.2e\n", fallbacks, diag_worst);
    return 0;
}

int hostsim_eval_loss_rows(int32_t n, int32_t nl, int32_t nv, const double* tl, const double* tv, const double* coef,
                           const double* in, const double* par, double* out) {
    LossMapDev lm{tl, tv, coef, nl, nv};
    for (int i = 0; i < n; ++i) eval_loss_rows_point(lm, i, n, in, par, out);
    return 0;
}

int hostsim_eval_interval(int32_t n, int32_t num_steps, int32_t num_approx, const double* in, double* out) {
    for (int i = 0; i < n; ++i) eval_interval_point<false>(i, n, num_steps, num_approx, in, out, nullptr);
    return 0;
}

int hostsim_set_integrate_losses(int on) { g_int_losses = on; return 0; }

// collocation integrator (mseetc_set_integrator): stages = 0 switches it off again
int hostsim_set_integrator(int32_t stages, const double* A, const double* w, int32_t max_newton) {
    g_irk_on = stages > 0;
    if (!g_irk_on) return 0;
    g_irk.d = stages; g_irk.maxNewton = max_newton;
    for (int i = 0; i < stages * stages; ++i) g_irk.A[i] = A[i];
    for (int i = 0; i < stages; ++i) g_irk.w[i] = w[i];
    return 0;
}
int hostsim_eval_interval_irk(int32_t n, int32_t num_steps, int32_t num_approx, const double* in, double* out) {
    for (int i = 0; i < n; ++i) eval_interval_point<true>(i, n, num_steps, num_approx, in, out, &g_irk);
    return 0;
}

}  // extern "C"
