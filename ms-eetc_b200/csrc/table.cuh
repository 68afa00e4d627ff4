// Trajectory tables of a batch: the derived columns that utils.postProcessDataFrame (reference utils.py:223-336) adds to every
// solution returned by casadiSolver.solve (ocp.py:407), for all instances at once.
//   cell_table    thread = (node k, instance): every column that is a function of rows k and k+1
//   inst_resim    thread = instance: time-domain re-simulation of the optimal controls, interval by interval with accumulated
//                 errors (reference utils.py:110-194: CVODES at 1e-12 / 1e-14; here classic RK4 with 24 and with 48 sub-steps
//                 per interval combined by Richardson extrapolation, |error| ~ 1e-13 on these intervals)
// Output: out[instance][node][TAB_NCOL], row-major, columns in the reference's order (index column first).
#pragma once
#include "core.cuh"

namespace mseetc {

enum TableCol {
    TAB_TIME = 0, TAB_POS, TAB_VEL, TAB_FEL, TAB_FPB, TAB_SLACK,                  // ocp.py:386-405
    TAB_LIMIT, TAB_GRAD, TAB_CURV,                                                // utils.py:230-232
    TAB_FACC, TAB_FRGB, TAB_FORCE, TAB_PMAX, TAB_PMIN,                            // utils.py:233-241
    TAB_LOSSES, TAB_ENERGY, TAB_EPNB, TAB_EKIN, TAB_ACC,                          // utils.py:247-259,291-294,322-330
    TAB_POS_SIM, TAB_VEL_SIM, TAB_ERR_POS, TAB_ERR_VEL,                           // utils.py:188-192
    TAB_NCOL
};

struct TableIO {
    const double* z;          // [n][Nmax*stp+2] solution vectors (reference variable order)
    const double* params;     // [PAR_N][n]
    const int32_t* nint;      // [n]
    const int32_t* trk_of;    // [n]
    const int32_t* trk_off;   // [tracks+1]
    const double* ds;         // [sum N_j]
    const double* c0;         // [sum N_j]
    const double* nodes;      // 4 planes of [sum (N_j+1)]: position, speed limit, gradient [permil], curvature
    size_t node_stride;       // doubles between those planes
    const double* mass;       // [n] train.mass (without the rotating-mass factor: Energy (kin), utils.py:294)
    const int32_t* status;    // [n] or null
    double* out;              // [n][Nmax+1][TAB_NCOL]
    int n, Nmax, withPn, lossKind;
};

MS_HD double tab_nan() {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double(0x7ff8000000000000LL);
#else
    return NAN;
#endif
}

MS_HD void cell_table(const TableIO& io, const LossMapDev& lm, int k, int s) {
    if (s >= io.n) return;
    const int N = io.nint[s];
    if (k > io.Nmax) return;
    const int stp = 4 + (io.withPn ? 1 : 0);
    double* row = io.out + ((size_t)s * (io.Nmax + 1) + k) * TAB_NCOL;
    const double nan = tab_nan();
    const bool ok = !io.status || io.status[s] == ST_SOLVE_SUCCEEDED || io.status[s] == ST_ACCEPTABLE;
    if (k > N || !ok) { for (int c = 0; c < TAB_NCOL; ++c) row[c] = nan; return; }
    const double* z = io.z + (size_t)s * ((size_t)io.Nmax * stp + 2);
    const int trk = io.trk_of[s], off = io.trk_off[trk], noff = off + trk;
    const double M = io.params[(size_t)P_MASS * io.n + s];
    const double kWh = 1e-6 / 3.6;
    const double t = (k < N) ? z[k * stp + stp - 2] : z[N * stp];
    const double b = (k < N) ? z[k * stp + stp - 1] : z[N * stp + 1];
    const double v = sqrt(b);
    row[TAB_TIME] = t;
    row[TAB_POS] = io.nodes[noff + k];
    row[TAB_VEL] = v;
    row[TAB_LIMIT] = io.nodes[io.node_stride + noff + k];
    row[TAB_GRAD] = io.nodes[2 * io.node_stride + noff + k];
    row[TAB_CURV] = io.nodes[3 * io.node_stride + noff + k];
    row[TAB_EKIN] = kWh * 0.5 * io.mass[s] * b;
    if (k == N) {                       // terminal row: no control (NaN, ocp.py:396-399); the pneumatic-brake column is 0 when there is no such brake
        const int cols[] = {TAB_FEL, TAB_SLACK, TAB_FACC, TAB_FRGB, TAB_FORCE, TAB_PMAX, TAB_PMIN, TAB_LOSSES, TAB_ENERGY, TAB_ACC};
        for (int c : cols) row[c] = nan;
        row[TAB_FPB] = io.withPn ? nan : 0.0;
        row[TAB_EPNB] = nan;
        return;
    }
    const double fs = z[k * stp], qs = io.withPn ? z[k * stp + 1] : 0.0, sl = z[k * stp + stp - 3];
    const double bn = (k + 1 < N) ? z[(k + 1) * stp + stp - 1] : z[N * stp + 1];
    const double vn = sqrt(bn);
    const double fel = fs * M, fpb = qs * M;
    const double facc = fel >= 0.0 ? fel : 0.0 * fel, frgb = fel < 0.0 ? fel : 0.0 * fel;
    const double dsk = io.nodes[noff + k + 1] - io.nodes[noff + k];
    row[TAB_FEL] = fel; row[TAB_FPB] = fpb; row[TAB_SLACK] = sl * M;
    row[TAB_FACC] = facc; row[TAB_FRGB] = frgb; row[TAB_FORCE] = facc + frgb + fpb;
    row[TAB_PMAX] = fmax(facc * v / 1e3, facc * vn / 1e3);
    row[TAB_PMIN] = fmin(frgb * v / 1e3, frgb * vn / 1e3);
    // losses of the interval by the mid-point rule with the UNSPLIT loss function (utils.py:243-259)
    double lossPerMetre;          // PL(f, vMid) / vMid, specific [N/kg]
    if (io.lossKind == 2) {
        LossPar p;
        const double eg = io.params[(size_t)P_DYN_ETAG * io.n + s];
        p.M = M; p.aux = io.params[(size_t)P_DYN_AUX * io.n + s]; p.cgT = (1.0 - eg) / eg; p.cgB = 1.0 - eg;
        p.fMax = io.params[(size_t)P_DYN_FMAX * io.n + s]; p.pMax = io.params[(size_t)P_DYN_PMAX * io.n + s]; p.scale = io.params[(size_t)P_DYN_SCALE * io.n + s];
        lossPerMetre = loss_full(lm, p, jvar0(0.5 * (v + vn)), jvar1(fs), fs >= 0.0).v;
    } else {
        lossPerMetre = fs > 0.0 ? io.params[(size_t)P_CT * io.n + s] * fs : -io.params[(size_t)P_CR * io.n + s] * fs;
    }
    const double losses = kWh * dsk * M * lossPerMetre;
    row[TAB_LOSSES] = losses;
    row[TAB_ENERGY] = kWh * dsk * facc + kWh * dsk * frgb + losses;
    row[TAB_EPNB] = -kWh * dsk * fpb;
    const double sr0 = io.params[(size_t)P_SR0 * io.n + s], sr1 = io.params[(size_t)P_SR1 * io.n + s], sr2 = io.params[(size_t)P_SR2 * io.n + s];
    row[TAB_ACC] = (facc + frgb + fpb) / M - (sr0 + sr1 * v + sr2 * b) - io.c0[off + k];
}

// one RK4 run over [0, dt] in m steps of d s/dt = v, d v/dt = f - (sr0 + sr1 v + sr2 v^2) - c0
MS_HD void resim_rk4(double& s, double& v, double f, double c0, double sr0, double sr1, double sr2, double dt, int m) {
    const double h = dt / m;
    for (int i = 0; i < m; ++i) {
        const double k1s = v, k1v = f - (sr0 + sr1 * v + sr2 * v * v) - c0;
        const double v2 = v + 0.5 * h * k1v;
        const double k2s = v2, k2v = f - (sr0 + sr1 * v2 + sr2 * v2 * v2) - c0;
        const double v3 = v + 0.5 * h * k2v;
        const double k3s = v3, k3v = f - (sr0 + sr1 * v3 + sr2 * v3 * v3) - c0;
        const double v4 = v + h * k3v;
        const double k4s = v4, k4v = f - (sr0 + sr1 * v4 + sr2 * v4 * v4) - c0;
        s += h / 6.0 * (k1s + 2.0 * k2s + 2.0 * k3s + k4s);
        v += h / 6.0 * (k1v + 2.0 * k2v + 2.0 * k3v + k4v);
    }
}

MS_HD void inst_resim(const TableIO& io, int s) {
    if (s >= io.n) return;
    const bool ok = !io.status || io.status[s] == ST_SOLVE_SUCCEEDED || io.status[s] == ST_ACCEPTABLE;
    if (!ok) return;                                    // cell_table has filled the rows with NaN
    const int N = io.nint[s];
    const int trk = io.trk_of[s], off = io.trk_off[trk];
    double* tab = io.out + (size_t)s * (io.Nmax + 1) * TAB_NCOL;
    const double M = io.params[(size_t)P_MASS * io.n + s];
    const double sr0 = io.params[(size_t)P_SR0 * io.n + s], sr1 = io.params[(size_t)P_SR1 * io.n + s], sr2 = io.params[(size_t)P_SR2 * io.n + s];
    double ps = tab[TAB_POS], pv = tab[TAB_VEL];
    tab[TAB_POS_SIM] = ps; tab[TAB_VEL_SIM] = pv; tab[TAB_ERR_POS] = 0.0; tab[TAB_ERR_VEL] = 0.0;
    for (int k = 0; k < N; ++k) {
        const double* row = tab + (size_t)k * TAB_NCOL;
        double* nxt = tab + (size_t)(k + 1) * TAB_NCOL;
        const double dt = nxt[TAB_TIME] - row[TAB_TIME], f = row[TAB_FORCE] / M, c0 = io.c0[off + k];
        double as = ps, av = pv, bs = ps, bv = pv;
        resim_rk4(as, av, f, c0, sr0, sr1, sr2, dt, 24);
        resim_rk4(bs, bv, f, c0, sr0, sr1, sr2, dt, 48);
        ps = bs + (bs - as) / 15.0;
        pv = bv + (bv - av) / 15.0;
        nxt[TAB_POS_SIM] = ps; nxt[TAB_VEL_SIM] = pv;
        nxt[TAB_ERR_POS] = fabs(ps - nxt[TAB_POS]); nxt[TAB_ERR_VEL] = fabs(pv - nxt[TAB_VEL]);
    }
}

}  // namespace mseetc
