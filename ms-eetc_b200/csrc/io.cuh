// Batch input scatter (caller arrays -> SoA planes) and result gather in the reference's own layout
// (zOpt of ocp.py:360, interleaved as ocp.py:166-181,248-249; g-row order of ocp.py:183-241).
#pragma once
#include "pit.cuh"

namespace mseetc {

struct BatchIO {
    const double* params;
    const int32_t* nint;
    const int32_t* trk_of;
    const int32_t* trk_off;
    const double* ds;
    const double* c0;
    const double* bmax;
    const double* tmin;
    double* z_out;
    double* lam_out;
    double* obj;
    double* kkt;
    int32_t* iters;
    int32_t* status;
};

MS_HD void inst_setup(const Ctx& c, const BatchIO& io, int s) {
    const Config& g = c.cfg;
    if (s >= g.nInst) return;
    for (int f = 0; f < PAR_N; ++f) c.P(f, s) = io.params[(size_t)f * g.nInst + s];
    c.I(SI_N_INT, s) = io.nint[s];
    inst_init(c, s);
    // screening with a known minimum trip duration: terminalTime is an upper bound on t_N (ocp.py:260-261)
    if (io.tmin && io.tmin[s] > 0.0 && (c.P(P_T, s) - c.P(P_T0, s)) < io.tmin[s] * (1.0 - MS_TMIN_MARGIN)) finish(c, s, ST_INFEASIBLE);
}

MS_HD void cell_setup(const Ctx& c, const BatchIO& io, int k, int s) {
    const Config& g = c.cfg;
    if (s >= g.nInst) return;
    const int N = io.nint[s];
    if (k > N) return;
    const int trk = io.trk_of[s];
    const int off = io.trk_off[trk];
    c.W(WS_TRK + TRK_DS, k, s) = (k < N) ? io.ds[off + k] : 0.0;
    c.W(WS_TRK + TRK_C0, k, s) = (k < N) ? io.c0[off + k] : 0.0;
    c.W(WS_TRK + TRK_BMAX, k, s) = io.bmax[off + trk + k];
}

// finalPass = false: only the instances that have finished (their slots are about to be re-used, see compact.cuh)
MS_HD void cell_extract(const Ctx& c, const BatchIO& io, int k, int s, bool finalPass = true) {
    const Config& g = c.cfg;
    if (s >= g.nInst || c.I(SI_EXTRACTED, s)) return;
    if (!finalPass && c.I(SI_PHASE, s) != PH_DONE) return;
    const int N = c.I(SI_N_INT, s);
    if (k > N) return;
    const int o = c.I(SI_ORIG, s);        // row of the caller's arrays
    const int it = c.I(SI_PARITY, s) ? WS_IT1 : WS_IT0;
    const int nu = 1 + (g.withPn ? 1 : 0), stp = 3 + nu, Nmax = g.NK - 1;
    if (io.z_out) {
        double* z = io.z_out + (size_t)o * ((size_t)Nmax * stp + 2) + (size_t)k * stp;
        if (k < N) {
            int o = 0;
            z[o++] = c.W(it + IT_FEL, k, s);
            if (g.withPn) z[o++] = c.W(it + IT_FPB, k, s);
            z[o++] = c.W(it + IT_SL, k, s);
            z[o++] = c.W(it + IT_T, k, s);
            z[o++] = c.W(it + IT_B, k, s);
        } else {
            z[0] = c.W(it + IT_T, k, s);
            z[1] = c.W(it + IT_B, k, s);
        }
    }
    if (io.lam_out && k < N) {
        const int rows = (g.withPower ? 2 : 0) + 3 + (g.energy ? 2 : 0);
        double* l = io.lam_out + (size_t)o * ((size_t)Nmax * rows) + (size_t)k * rows;
        int o = 0;
        if (g.withPower) { l[o++] = c.W(it + IT_YD + R_P0, k, s); l[o++] = c.W(it + IT_YD + R_P1, k, s); }
        l[o++] = c.W(it + IT_YD + R_ACC, k, s);
        l[o++] = c.W(it + IT_YT, k, s);
        l[o++] = c.W(it + IT_YB, k, s);
        if (g.energy) { l[o++] = c.W(it + IT_YD + R_LTR, k, s); l[o++] = c.W(it + IT_YD + R_LRG, k, s); }
    }
    if (k == 0) {
        if (io.obj) io.obj[o] = c.D(SD_FOBJ, s);
        if (io.kkt) io.kkt[o] = c.D(SD_KKT, s);
        if (io.iters) io.iters[o] = c.I(SI_ITERS, s);
        int st = c.I(SI_STATUS, s);
        io.status[o] = (st == ST_RUNNING) ? (int)ST_MAXITER : st;
    }
}

// integrateLosses = True: the multipliers of the shooting-time rows in the reference's formulation.  The device takes the duration in
// the loss rows from the shooting function (core.cuh, loss_energy_rows), the reference from t_{k+1} - t_k; the two Lagrangians have
// the same stationary points with  y_t(reference) = y_t(device) + y_ltr dE_tr/d(duration) + y_lrg dE_rgb/d(duration),  and
// dE/d(duration) is the loss power at the end of the interval.  Applied to lam_out after cell_extract (final pass only).
template <bool IRK>
MS_HD void cell_fix_time_multiplier_intl(const Ctx& c, const BatchIO& io, int k, int s) {
    const Config& g = c.cfg;
    if (s >= g.nInst || !io.lam_out || !g.energy) return;
    const int N = c.I(SI_N_INT, s);
    if (k >= N) return;
    const int o = c.I(SI_ORIG, s);
    const int it = c.I(SI_PARITY, s) ? WS_IT1 : WS_IT0;
    const int rows = (g.withPower ? 2 : 0) + 3 + 2, Nmax = g.NK - 1;
    const double b = c.W(it + IT_B, k, s), b1 = c.W(it + IT_B, k + 1, s), fel = c.W(it + IT_FEL, k, s), fpb = g.withPn ? c.W(it + IT_FPB, k, s) : 0.0;
    const IntervalCoef q = load_coef(c, k, s);
    Jet2 tau, phi;
    if (IRK) shoot_irk(jvar0(b), jvar1(fel + fpb), q, g.numSteps, g.numApprox, *c.irk, tau, phi);
    else shoot<Jet2>(jvar0(b), jvar1(fel + fpb), q, g.numSteps, g.numApprox, tau, phi);
    Jet3 etr, erg;
    loss_energy_rows(c, s, q, b, b1, fel, fpb, tau, etr, erg, true);
    const double pl[2] = {etr.g[0], erg.g[0]};
    double* l = io.lam_out + (size_t)o * ((size_t)Nmax * rows) + (size_t)k * rows;
    const int iyt = (g.withPower ? 2 : 0) + 1;
    l[iyt] += c.W(it + IT_YD + R_LTR, k, s) * pl[0] + c.W(it + IT_YD + R_LRG, k, s) * pl[1];
}

// kernel-level parity hook: one shooting interval with sensitivities (train.py:347-364)
template <bool IRK>
MS_HD void eval_interval_point(int i, int n, int numSteps, int numApprox, const double* in, double* out, const IrkTab* irk) {
    if (i >= n) return;
    IntervalCoef q;
    const double b0 = in[i], F = in[n + i];
    q.ds = in[2 * n + i]; q.c0 = in[3 * n + i]; q.sr0 = in[4 * n + i]; q.sr1 = in[5 * n + i]; q.sr2 = in[6 * n + i];
    Jet2 tau, phi;
    if (IRK) shoot_irk(jvar0(b0), jvar1(F), q, numSteps, numApprox, *irk, tau, phi);
    else shoot<Jet2>(jvar0(b0), jvar1(F), q, numSteps, numApprox, tau, phi);
    const Jet2* js[2] = {&tau, &phi};
    for (int a = 0; a < 2; ++a) {
        out[(6 * a + 0) * (size_t)n + i] = js[a]->v;
        out[(6 * a + 1) * (size_t)n + i] = js[a]->g0;
        out[(6 * a + 2) * (size_t)n + i] = js[a]->g1;
        out[(6 * a + 3) * (size_t)n + i] = js[a]->h00;
        out[(6 * a + 4) * (size_t)n + i] = js[a]->h01;
        out[(6 * a + 5) * (size_t)n + i] = js[a]->h11;
    }
}

// kernel-level parity hook of the dynamic loss rows
MS_HD void eval_loss_rows_point(const LossMapDev& lm, int i, int n, const double* in, const double* par, double* out) {
    if (i >= n) return;
    LossPar p;
    const double eg = par[(size_t)P_DYN_ETAG * n + i];
    p.M = par[(size_t)P_MASS * n + i]; p.aux = par[(size_t)P_DYN_AUX * n + i]; p.cgT = (1.0 - eg) / eg; p.cgB = 1.0 - eg;
    p.fMax = par[(size_t)P_DYN_FMAX * n + i]; p.pMax = par[(size_t)P_DYN_PMAX * n + i]; p.scale = par[(size_t)P_DYN_SCALE * n + i];
    LossRow r[2];
    loss_rows_dynamic(lm, p, in[i], in[n + i], in[2 * (size_t)n + i], r[0], r[1]);
    for (int a = 0; a < 2; ++a) {
        const double vals[10] = {r[a].v, r[a].gF, r[a].g0, r[a].g1, r[a].hFF, r[a].hF0, r[a].hF1, r[a].h00, r[a].h01, r[a].h11};
        for (int j = 0; j < 10; ++j) out[(size_t)(10 * a + j) * n + i] = vals[j];
    }
}

// workspace carving (all sizes in bytes, 256-byte aligned sections)
struct WsPlan {
    size_t off_ws, off_par, off_sd, off_si, off_done, off_plan, total;
};
inline WsPlan plan_workspace(int S, int NK) {
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    WsPlan p;
    size_t o = 0;
    p.off_ws = o; o = al(o + sizeof(double) * (size_t)WS_FIELDS * NK * S);
    p.off_par = o; o = al(o + sizeof(double) * (size_t)PAR_N * S);
    p.off_sd = o; o = al(o + sizeof(double) * (size_t)SD_N * S);
    p.off_si = o; o = al(o + sizeof(int) * (size_t)SI_N * S);
    p.off_done = o; o = al(o + 256);   // int done; then 4 x uint64 cell counters at +64
    p.off_plan = o; o = al(o + sizeof(int) * (size_t)(2 * S + 8));      // compaction plan (compact.cuh)
    p.total = o;
    return p;
}
inline int pad_slots(int n) { return (n + 31) & ~31; }

}  // namespace mseetc
