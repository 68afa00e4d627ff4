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
    g.maxIter = pr->max_iterations; g.tol = pr->tol; g.muInit = pr->mu_init; g.initMode = init_mode;
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
    c.lm.tl = lm_tl; c.lm.tv = lm_tv; c.lm.coef = lm_coef; c.lm.nl = lm_nl; c.lm.nv = lm_nv;
    const bool dyn = (pr->loss_kind == 2 && pr->energy_optimal);
    BatchIO io{params, nint, trk_of, trk_off, ds, c0, bmax, tmin, z_out, lam_out, obj, kkt, iters, status};
    for (int s = 0; s < g.S; ++s) inst_setup(c, io, s);
    for (int k = 0; k < g.NK; ++k) for (int s = 0; s < g.S; ++s) cell_setup(c, io, k, s);
    for (int s = 0; s < g.S; ++s) { inst_screen(c, s); inst_profile(c, s); }
    for (int k = 0; k < g.NK; ++k) for (int s = 0; s < g.S; ++s) { if (dyn) cell_init<true>(c, k, s); else cell_init<false>(c, k, s); }
    int tick = 0;
    const int maxTicks = 20 * pr->max_iterations + 50;
    auto report = [&](const char* tag) {
        int s = verbose_inst;
        if (s < 0 || s >= n) return;
        printf("%s tick %3d it %3d ph %d f=%.10e th=%.3e dinf=%.2e pinf=%.2e mu=%.1e a=%.3e az=%.3e kkt=%.2e nls=%d nreg=%d\n", tag,
               tick, c.I(SI_ITERS, s), c.I(SI_PHASE, s), c.D(SD_FOBJ, s), c.D(SD_THETA, s), c.D(SD_DINF, s), c.D(SD_PINF, s),
               c.D(SD_MU, s), c.D(SD_ALPHA, s), c.D(SD_ALPHA_Z, s), c.D(SD_KKT, s), c.I(SI_NLS, s), c.I(SI_NREG, s));
    };
    for (;;) {
        for (int k = 0; k < g.NK; ++k) for (int s = 0; s < g.S; ++s) { if (dyn) cell_eval<true>(c, k, s); else cell_eval<false>(c, k, s); }
        for (int s = 0; s < g.nInst; ++s) {
            if (c.I(SI_PHASE, s) != PH_EVAL) continue;
            const int N = c.I(SI_N_INT, s), it = c.I(SI_PARITY, s) ? WS_IT1 : WS_IT0;
            KktAcc tot, part;
            kkt_init(tot);
            for (int w = 0; w < RED_W; ++w) { kkt_partials(c, s, N, it, w, RED_W, part); kkt_combine(tot, part); }
            inst_kkt(c, s, tot);
        }
        for (int s = 0; s < g.S; ++s) {
            DirectFetch<BwdFields> fb; DirectFetch<FwdFields> ff;
            if (pit_lanes > 1) inst_step_pit_emulated(c, s, pit_lanes, fb, ff);   // lanes-per-instance (parallel-in-time) variant
            else inst_step(c, s, fb, ff);
        }
        for (int k = 0; k < g.NK; ++k) for (int s = 0; s < g.S; ++s) { if (dyn) cell_step<true>(c, k, s); else cell_step<false>(c, k, s); }
        for (int s = 0; s < g.nInst; ++s) {
            if (c.I(SI_PHASE, s) != PH_STEPPED) continue;
            const int N = c.I(SI_N_INT, s);
            double tot[3] = {1.0, 1.0, 0.0}, part[3];
            for (int w = 0; w < RED_W; ++w) { alpha_partials(c, s, N, w, RED_W, part); tot[0] = fmin(tot[0], part[0]); tot[1] = fmin(tot[1], part[1]); tot[2] += part[2]; }
            inst_alpha(c, s, tot);
        }
        report("step ");
        if (*c.done >= n || tick >= maxTicks) break;
        for (int k = 0; k < g.NK; ++k) for (int s = 0; s < g.S; ++s) { if (dyn) cell_trial<true>(c, k, s); else cell_trial<false>(c, k, s); }
        for (int s = 0; s < g.nInst; ++s) {
            if (c.I(SI_PHASE, s) != PH_TRIAL) continue;
            const int N = c.I(SI_N_INT, s);
            double tot[4] = {0, 0, 0, 0}, part[4];
            for (int w = 0; w < RED_W; ++w) { trial_partials(c, s, N, w, RED_W, part); for (int f = 0; f < 4; ++f) tot[f] += part[f]; }
            inst_decide(c, s, tot);
        }
        ++tick;
        if (*c.done >= n) break;
    }
    for (int k = 0; k < g.NK; ++k) for (int s = 0; s < g.S; ++s) cell_extract(c, io, k, s);
    if (ticks_out) *ticks_out = tick;
    return 0;
}

int hostsim_eval_loss_rows(int32_t n, int32_t nl, int32_t nv, const double* tl, const double* tv, const double* coef,
                           const double* in, const double* par, double* out) {
    LossMapDev lm{tl, tv, coef, nl, nv};
    for (int i = 0; i < n; ++i) eval_loss_rows_point(lm, i, n, in, par, out);
    return 0;
}

int hostsim_eval_interval(int32_t n, int32_t num_steps, int32_t num_approx, const double* in, double* out) {
    for (int i = 0; i < n; ++i) eval_interval_point(i, n, num_steps, num_approx, in, out);
    return 0;
}

}  // extern "C"
