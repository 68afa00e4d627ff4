// Parallel-in-time Riccati: G lanes per instance, each lane owns a contiguous chunk of shooting intervals.
//
// The backward recursion over a chunk maps the value function at its end to the value function at its start by a
// linear-fractional transformation.  Following Sarkka & Garcia-Fernandez ("Temporal parallelization of dynamic
// programming and linear quadratic control", IEEE TAC 2023) a chunk is summarised by the element
//     V_{i->j}(x_i, x_j) = max_lam  1/2 x_i' J x_i - x_i' eta - 1/2 lam' C lam - lam' (x_j - A x_i - b)
// and elements combine associatively, so the value functions at all chunk boundaries follow from a parallel
// suffix scan (warp shuffles on the device).  Each lane then runs the ordinary Riccati recursion inside its chunk
// (exact inertia test, same feedback gains as the sequential sweep), the closed-loop chunk transitions are
// combined by a prefix scan to give the state step at every chunk start, and the forward sweep runs chunk-local.
//   phase A  chunk element by sequential accumulation (needs the stage-local control Hessian positive definite)
//   scan 1   suffix scan of elements            -> (P, p) at every chunk end
//   phase C  Riccati backward inside the chunk  -> K, k, P, p per interval; closed-loop chunk transition
//   scan 2   prefix scan of affine transitions  -> d x at every chunk start
//   phase F  forward sweep inside the chunk
// Used for small batches, single instances and long horizons; G = 1 degenerates to the sequential sweeps.
#pragma once
#include "riccati.cuh"

namespace mseetc {

#ifndef PIT_REFINE
#define PIT_REFINE 1
#endif

struct Elem {
    double A[9], b[3], C[6], eta[3], J[6];   // C, J symmetric: (00,01,02,11,12,22)
};

MS_HD void sym_to_full(const double* s, double F[3][3]) {
    F[0][0] = s[0]; F[0][1] = F[1][0] = s[1]; F[0][2] = F[2][0] = s[2];
    F[1][1] = s[3]; F[1][2] = F[2][1] = s[4]; F[2][2] = s[5];
}
MS_HD void full_to_sym(const double F[3][3], double* s) {
    s[0] = F[0][0]; s[1] = 0.5 * (F[0][1] + F[1][0]); s[2] = 0.5 * (F[0][2] + F[2][0]);
    s[3] = F[1][1]; s[4] = 0.5 * (F[1][2] + F[2][1]); s[5] = F[2][2];
}

MS_HD void elem_identity(Elem& e) {
    for (int i = 0; i < 9; ++i) e.A[i] = 0.0;
    e.A[0] = e.A[4] = e.A[8] = 1.0;
    for (int i = 0; i < 3; ++i) { e.b[i] = 0.0; e.eta[i] = 0.0; }
    for (int i = 0; i < 6; ++i) { e.C[i] = 0.0; e.J[i] = 0.0; }
}
// value function 1/2 x'Px + p'x seen as an element that ends the horizon
MS_HD void elem_from_value(Elem& e, const double P[3][3], const double p[3]) {
    for (int i = 0; i < 9; ++i) e.A[i] = 0.0;
    for (int i = 0; i < 3; ++i) { e.b[i] = 0.0; e.eta[i] = -p[i]; }
    for (int i = 0; i < 6; ++i) e.C[i] = 0.0;
    full_to_sym(P, e.J);
}

// element of one interval: eliminate the controls with the stage-local R (false if R is not positive definite)
MS_HD bool elem_from_stage(const StageQP& q, Elem& e) {
    const double (*M)[6] = q.M;
    const double d0 = M[3][3];
    if (!(d0 > 0.0) || !isfinite(d0)) return false;
    const double i00 = rcp(sqrt(d0));
    const double l10 = M[4][3] * i00, l20 = M[5][3] * i00;
    const double d1 = M[4][4] - l10 * l10;
    if (!(d1 > 0.0) || !isfinite(d1)) return false;
    const double i11 = rcp(sqrt(d1));
    const double l21 = (M[5][4] - l20 * l10) * i11;
    const double d2 = M[5][5] - l20 * l20 - l21 * l21;
    if (!(d2 > 0.0) || !isfinite(d2)) return false;
    const double i22 = rcp(sqrt(d2));
    // X = R^{-1} [S | r_u | B']   (3 + 1 + 3 columns)
    double X[3][7];
    for (int j = 0; j < 7; ++j) {
        double r0, r1, r2;
        if (j < 3) { r0 = M[3][j]; r1 = M[4][j]; r2 = M[5][j]; }
        else if (j == 3) { r0 = q.m[3]; r1 = q.m[4]; r2 = q.m[5]; }
        else { r0 = q.G[j - 4][3]; r1 = q.G[j - 4][4]; r2 = q.G[j - 4][5]; }
        const double y0 = r0 * i00, y1 = (r1 - l10 * y0) * i11, y2 = (r2 - l20 * y0 - l21 * y1) * i22;
        const double x2 = y2 * i22, x1 = (y1 - l21 * x2) * i11, x0 = (y0 - l10 * x1 - l20 * x2) * i00;
        X[0][j] = x0; X[1][j] = x1; X[2][j] = x2;
    }
    double Cf[3][3], Jf[3][3];
    for (int i = 0; i < 3; ++i) {
        // A_e = A - B R^{-1} S ; b = r - B R^{-1} r_u ; C = B R^{-1} B'
        e.b[i] = q.r[i] - (q.G[i][3] * X[0][3] + q.G[i][4] * X[1][3] + q.G[i][5] * X[2][3]);
        for (int j = 0; j < 3; ++j) {
            e.A[3 * i + j] = q.G[i][j] - (q.G[i][3] * X[0][j] + q.G[i][4] * X[1][j] + q.G[i][5] * X[2][j]);
            Cf[i][j] = q.G[i][3] * X[0][4 + j] + q.G[i][4] * X[1][4 + j] + q.G[i][5] * X[2][4 + j];
            // J = Q - S' R^{-1} S
            Jf[i][j] = M[i][j] - (M[3][i] * X[0][j] + M[4][i] * X[1][j] + M[5][i] * X[2][j]);
        }
        // eta = -(q - S' R^{-1} r_u)
        e.eta[i] = -(q.m[i] - (M[3][i] * X[0][3] + M[4][i] * X[1][3] + M[5][i] * X[2][3]));
    }
    full_to_sym(Cf, e.C);
    full_to_sym(Jf, e.J);
    return true;
}

// out = first (i->j) (x) second (j->k); false if I + C1 J2 is (numerically) singular
MS_HD bool elem_combine(const Elem& e1, const Elem& e2, Elem& out) {
    double C1[3][3], J2[3][3], C2[3][3], J1[3][3];
    sym_to_full(e1.C, C1); sym_to_full(e2.J, J2); sym_to_full(e2.C, C2); sym_to_full(e1.J, J1);
    // X = I + C1 J2, T = X^{-1} by the adjugate
    double X[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) X[i][j] = (i == j ? 1.0 : 0.0) + C1[i][0] * J2[0][j] + C1[i][1] * J2[1][j] + C1[i][2] * J2[2][j];
    // T = X^{-1} by Gauss-Jordan elimination with partial pivoting (X is badly scaled late in the IP iteration:
    // barrier terms of active bounds put entries of 1e9 next to O(1) ones)
    double T[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int col = 0; col < 3; ++col) {
        int piv = col;
        double best = fabs(X[col][col]);
        for (int r = col + 1; r < 3; ++r) if (fabs(X[r][col]) > best) { best = fabs(X[r][col]); piv = r; }
        if (!(best > 1e-300) || !isfinite(best)) return false;
        if (piv != col)
            for (int j = 0; j < 3; ++j) {
                double t = X[col][j]; X[col][j] = X[piv][j]; X[piv][j] = t;
                t = T[col][j]; T[col][j] = T[piv][j]; T[piv][j] = t;
            }
        const double ip = rcp(X[col][col]);
        for (int j = 0; j < 3; ++j) { X[col][j] *= ip; T[col][j] *= ip; }
        for (int r = 0; r < 3; ++r) {
            if (r == col) continue;
            const double f = X[r][col];
            for (int j = 0; j < 3; ++j) { X[r][j] -= f * X[col][j]; T[r][j] -= f * T[col][j]; }
        }
    }
    // A2T = A2 T
    double A2T[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) A2T[i][j] = e2.A[3 * i] * T[0][j] + e2.A[3 * i + 1] * T[1][j] + e2.A[3 * i + 2] * T[2][j];
    // A = A2 T A1 ; b = A2 T (b1 + C1 eta2) + b2 ; C = A2 T C1 A2' + C2
    double w[3];
    for (int i = 0; i < 3; ++i) w[i] = e1.b[i] + C1[i][0] * e2.eta[0] + C1[i][1] * e2.eta[1] + C1[i][2] * e2.eta[2];
    double TC1[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) TC1[i][j] = A2T[i][0] * C1[0][j] + A2T[i][1] * C1[1][j] + A2T[i][2] * C1[2][j];
    double Cn[3][3];
    for (int i = 0; i < 3; ++i) {
        out.b[i] = A2T[i][0] * w[0] + A2T[i][1] * w[1] + A2T[i][2] * w[2] + e2.b[i];
        for (int j = 0; j < 3; ++j) {
            out.A[3 * i + j] = A2T[i][0] * e1.A[j] + A2T[i][1] * e1.A[3 + j] + A2T[i][2] * e1.A[6 + j];
            Cn[i][j] = TC1[i][0] * e2.A[3 * j] + TC1[i][1] * e2.A[3 * j + 1] + TC1[i][2] * e2.A[3 * j + 2] + C2[i][j];
        }
    }
    // A1T' = A1' T' ; eta = A1' T' (eta2 - J2 b1) + eta1 ; J = A1' T' J2 A1 + J1
    double A1Tt[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) A1Tt[i][j] = e1.A[i] * T[j][0] + e1.A[3 + i] * T[j][1] + e1.A[6 + i] * T[j][2];
    double u[3];
    for (int i = 0; i < 3; ++i) u[i] = e2.eta[i] - (J2[i][0] * e1.b[0] + J2[i][1] * e1.b[1] + J2[i][2] * e1.b[2]);
    double TJ[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) TJ[i][j] = A1Tt[i][0] * J2[0][j] + A1Tt[i][1] * J2[1][j] + A1Tt[i][2] * J2[2][j];
    double Jn[3][3];
    for (int i = 0; i < 3; ++i) {
        out.eta[i] = A1Tt[i][0] * u[0] + A1Tt[i][1] * u[1] + A1Tt[i][2] * u[2] + e1.eta[i];
        for (int j = 0; j < 3; ++j) Jn[i][j] = TJ[i][0] * e1.A[j] + TJ[i][1] * e1.A[3 + j] + TJ[i][2] * e1.A[6 + j] + J1[i][j];
    }
    full_to_sym(Cn, out.C);
    full_to_sym(Jn, out.J);
    return true;
}

// chunk geometry: the generic intervals 0 .. N-2 are split over G lanes; interval N-1 (terminal-speed elimination)
// belongs to lane G-1 in addition to its chunk
MS_HD void pit_chunk(int N, int G, int lane, int& kLo, int& kHi) {
    const int Ng = N - 1;
    const int L = (Ng + G - 1) / G;
    kLo = lane * L; if (kLo > Ng) kLo = Ng;
    kHi = kLo + L; if (kHi > Ng) kHi = Ng;
}

// phase A for one lane: E = e_{kLo} (x) ... (x) e_{kHi-1} [(x) Eend]
template <class Fetch>
MS_HD bool pit_phase_a(const Ctx& c, int s, int kLo, int kHi, double mu, double delta, Fetch& fetch, bool haveEnd, Elem& E) {
    const double pn = c.cfg.withPn ? 1.0 : 0.0;
    if (!haveEnd) elem_identity(E);
    if (kHi <= kLo) return true;
    bool have = haveEnd;
    fetch.start(c, s, kHi - 1, kLo, -1);
    for (int k = kHi - 1; k >= kLo; --k) {
        double v[BwdFields::NF];
        fetch.get(c, k, s, v);
        StageQP q;
        double vs[6];
        load_scol(c, k, s, vs);
        stage_build(v, vs, mu, delta, pn, false, q);
        Elem e;
        if (!elem_from_stage(q, e)) return false;
        if (!have) { E = e; have = true; }
        else {
            Elem t;
            if (!elem_combine(e, E, t)) return false;
            E = t;
        }
    }
    return true;
}

// affine map x -> M x + m
struct Aff {
    double M[9], m[3];
};
MS_HD void aff_identity(Aff& a) {
    for (int i = 0; i < 9; ++i) a.M[i] = 0.0;
    a.M[0] = a.M[4] = a.M[8] = 1.0;
    a.m[0] = a.m[1] = a.m[2] = 0.0;
}
// out = second o first  (first applied first)
MS_HD void aff_compose(const Aff& first, const Aff& second, Aff& out) {
    for (int i = 0; i < 3; ++i) {
        out.m[i] = second.m[i] + second.M[3 * i] * first.m[0] + second.M[3 * i + 1] * first.m[1] + second.M[3 * i + 2] * first.m[2];
        for (int j = 0; j < 3; ++j)
            out.M[3 * i + j] = second.M[3 * i] * first.M[j] + second.M[3 * i + 1] * first.M[3 + j] + second.M[3 * i + 2] * first.M[6 + j];
    }
}

// ---- host emulation of the lane-parallel driver (tests/hostsim); the device version is in mseetc_b200.cu -------
// Runs the identical phases with the lanes of one instance in a loop.  Returns false on wrong inertia.
template <class FetchB, class FetchF>
inline bool pit_direction_emulated(const Ctx& c, int s, int N, int G, double mu, double delta, FetchB& fb, FetchF& ff) {
    Elem E[32];
    bool ok = true;
    // last interval (terminal-speed elimination) by lane G-1: value function of node N-1
    double Pl[3][3], pl[3];
    terminal_value(c, s, N, mu, delta, Pl, pl);
    ok = riccati_backward_range(c, s, N, N - 1, N, mu, delta, fb, Pl, pl, nullptr, nullptr);
    if (!ok) return false;
    for (int l = 0; l < G; ++l) {
        int kLo, kHi;
        pit_chunk(N, G, l, kLo, kHi);
        bool haveEnd = (l == G - 1);
        if (haveEnd) elem_from_value(E[l], Pl, pl);
        if (!pit_phase_a(c, s, kLo, kHi, mu, delta, fb, haveEnd, E[l])) return false;
    }
    // suffix scan (Hillis-Steele): after it E[l] = E_l (x) ... (x) E_{G-1}
    for (int d = 1; d < G; d <<= 1) {
        Elem old[32];
        for (int l = 0; l < G; ++l) old[l] = E[l];
        for (int l = 0; l + d < G; ++l)
            if (!elem_combine(old[l], old[l + d], E[l])) return false;
    }
    Aff T[32];
    for (int pass = 0; pass <= PIT_REFINE; ++pass) {
        // pass 0: chunk-end value functions from the scan; later passes: from the neighbour's stable in-chunk
        // recursion of the previous pass (block-Jacobi refinement; the Riccati map contracts errors of its end value)
        double Pe[32][3][3], pe[32][3];
        for (int l = 0; l < G; ++l) {
            int kLo, kHi;
            pit_chunk(N, G, l, kLo, kHi);
            if (l == G - 1) { for (int i = 0; i < 3; ++i) { pe[l][i] = pl[i]; for (int j = 0; j < 3; ++j) Pe[l][i][j] = Pl[i][j]; } }
            else if (pass == 0) { sym_to_full(E[l + 1].J, Pe[l]); for (int i = 0; i < 3; ++i) pe[l][i] = -E[l + 1].eta[i]; }
            else {
                double sy[6];
                for (int i = 0; i < 6; ++i) sy[i] = c.W(WS_RIC + RIC_P + i, kHi, s);
                sym_to_full(sy, Pe[l]);
                for (int i = 0; i < 3; ++i) pe[l][i] = c.W(WS_RIC + RIC_PV + i, kHi, s);
            }
        }
        for (int l = 0; l < G; ++l) {
            int kLo, kHi;
            pit_chunk(N, G, l, kLo, kHi);
            aff_identity(T[l]);
            if (!riccati_backward_range(c, s, N, kLo, kHi, mu, delta, fb, Pe[l], pe[l], T[l].M, T[l].m)) return false;
        }
    }
    // prefix scan of the chunk transitions: after it T[l] = T_l o ... o T_0
    for (int d = 1; d < G; d <<= 1) {
        Aff old[32];
        for (int l = 0; l < G; ++l) old[l] = T[l];
        for (int l = d; l < G; ++l) aff_compose(old[l - d], old[l], T[l]);
    }
    c.W(WS_ST + ST_T, 0, s) = 0.0;
    c.W(WS_ST + ST_B, 0, s) = 0.0;
    for (int l = 0; l < G; ++l) {
        int kLo, kHi;
        pit_chunk(N, G, l, kLo, kHi);
        double dx[3] = {0.0, 0.0, 0.0};
        if (l > 0) { dx[0] = T[l - 1].m[0]; dx[1] = T[l - 1].m[1]; dx[2] = T[l - 1].m[2]; }
        riccati_forward_range(c, s, N, kLo, kHi, mu, delta, ff, dx);
        if (l == G - 1) riccati_forward_range(c, s, N, N - 1, N, mu, delta, ff, dx);
    }
    return true;
}

// per-instance driver with the inertia-correction ladder, lane-parallel direction (host emulation)
template <class FetchB, class FetchF>
inline void inst_step_pit_emulated(const Ctx& c, int s, int G, FetchB& fb, FetchF& ff) {
    const Config& g = c.cfg;
    if (s >= g.nInst || c.I(SI_PHASE, s) != PH_FACTOR) return;
    const int N = c.I(SI_N_INT, s);
    const double mu = c.D(SD_MU, s);
    double delta = 0.0;
    const double dlast = c.D(SD_DELTA_LAST, s);
    bool ok = false;
    for (int tries = 0; tries < 40; ++tries) {
        count_cells(c, 2, N);
        if (pit_direction_emulated(c, s, N, G, mu, delta, fb, ff)) { ok = true; break; }
        c.I(SI_NREG, s) += 1;
        if (delta == 0.0) delta = (dlast == 0.0) ? 1e-4 : fmax(1e-20, dlast / 3.0);
        else delta *= (dlast == 0.0) ? 100.0 : 8.0;
        if (delta > 1e40) break;
    }
    if (!ok) { finish(c, s, ST_STEP_FAILED); return; }
    if (delta > 0.0) c.D(SD_DELTA_LAST, s) = delta;
    c.D(SD_DELTA, s) = delta;
    count_cells(c, 3, N);
    c.I(SI_PHASE, s) = PH_STEPPED;
}

}  // namespace mseetc
