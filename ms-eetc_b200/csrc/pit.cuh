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


// =====================================================================================================================
// Chunked parallel-in-time sweeps, second formulation (used for short and long horizons alike).
//
// A chunk of consecutive intervals [kLo, kHi) maps the value function (Pi, pi) at its end node to the value function at its
// start node by a linear-fractional transformation.  Its coefficients are what a Riccati recursion with ZERO terminal value
// produces anyway -- value function (Pbar, pbar) at the chunk start, closed-loop transition x_e = Phi x_a + phi under the
// zero-terminal gains -- plus the closed-loop controllability Gramian  W = sum_j Phi_{e<-j+1} B Rt_j^{-1} B' Phi_{e<-j+1}':
//     P_a = Pbar + Phi' T Phi,                 T = Pi (I + W Pi)^{-1} = (Pi^{-1} + W)^{-1}   (symmetric)
//     p_a = pbar + Phi' (pi + T (phi - W pi))
// (same element as in Sarkka & Garcia-Fernandez, but accumulated by the stable, structure-exploiting stage recursion of
// riccati.cuh instead of by repeated generic combinations, and only ever applied to a genuine value function.)
//   phase A   every chunk but the last: zero-terminal recursion -> element (27 numbers); the last chunk runs the ordinary
//             recursion from the terminal node (exact) and keeps its factors
//   chain     the value functions at the chunk ends follow one after the other, last chunk first: G - 2 applications of an
//             element to a value function (3x3 Cholesky factorisations, no ill-conditioned unsymmetric inverse)
//   phase C   ordinary recursion inside every chunk from its end value: gains, value functions, exact inertia test; closed-loop
//             transition of the chunk
//   chain     state step at the chunk starts (G affine maps applied one after the other)
//   phase F   forward sweep inside every chunk
struct ChunkElem {
    double P[6], p[3], Phi[9], phi[3], W[6];      // P, W symmetric: (00,01,02,11,12,22)
};

template <bool REG, class Fetch>
MS_HD bool chunk_element_t(const Ctx& c, int s, int kLo, int kHi, double mu, double delta, Fetch& fetch, const double* refP, const double* refp,
                           ChunkElem& E) {
    const double pn = c.cfg.withPn ? 1.0 : 0.0;
    double P[3][3], p[3];
    sym_to_full(refP, P);
    for (int i = 0; i < 3; ++i) p[i] = refp[i];
    for (int i = 0; i < 9; ++i) E.Phi[i] = 0.0;
    E.Phi[0] = E.Phi[4] = E.Phi[8] = 1.0;
    for (int i = 0; i < 3; ++i) E.phi[i] = 0.0;
    for (int i = 0; i < 6; ++i) E.W[i] = 0.0;
    bool ok = true;
    if (kHi > kLo) {
        fetch.start(c, s, kHi - 1, kLo, -1);
        for (int k = kHi - 1; k >= kLo; --k) {
            double v[BwdFields::NF], vs[6], K[3][3], kf[3], cb[4];
            fetch.get(c, k, s, v);
            if (REG) load_scol(c, k, s, vs);
            if (!stage_riccati_sparse(v, REG ? vs : nullptr, mu, delta, pn, P, p, K, kf, cb)) { ok = false; if (!Fetch::COLLECTIVE) return false; }
            const double tb = v[QP_TAU_B], tF = v[QP_TAU_F], pb = v[QP_PHI_B], pF = v[QP_PHI_F];
            // Gramian: W += Y Rt^{-1} Y',  Y = Phi_{e<-k+1} B,  B = [(tF, pF, 1)  pn (tF, pF, 0)]
            double yF[3], yQ[3], xF[3], xQ[3];
            for (int i = 0; i < 3; ++i) {
                const double g = E.Phi[3 * i] * tF + E.Phi[3 * i + 1] * pF;
                yF[i] = g + E.Phi[3 * i + 2]; yQ[i] = pn * g;
                sym2_solve(cb[0], cb[1], cb[2], cb[3], yF[i], yQ[i], xF[i], xQ[i]);
            }
            E.W[0] += yF[0] * xF[0] + yQ[0] * xQ[0]; E.W[1] += yF[0] * xF[1] + yQ[0] * xQ[1]; E.W[2] += yF[0] * xF[2] + yQ[0] * xQ[2];
            E.W[3] += yF[1] * xF[1] + yQ[1] * xQ[1]; E.W[4] += yF[1] * xF[2] + yQ[1] * xQ[2]; E.W[5] += yF[2] * xF[2] + yQ[2] * xQ[2];
            // closed loop of this interval under the zero-terminal gains, composed onto the transition of the intervals behind it
            const double kfw = kf[0] + pn * kf[1];
            double Mk[9], mk[3];
            mk[0] = v[QP_RT] + tF * kfw; mk[1] = v[QP_RB] + pF * kfw; mk[2] = kf[0];
            for (int j = 0; j < 3; ++j) {
                const double kw = K[0][j] + pn * K[1][j];
                Mk[j] = (j == 0 ? 1.0 : j == 1 ? tb : 0.0) + tF * kw;
                Mk[3 + j] = (j == 1 ? pb : 0.0) + pF * kw;
                Mk[6 + j] = K[0][j];
            }
            closed_loop_compose(E.Phi, E.phi, Mk, mk);
        }
    }
    full_to_sym(P, E.P);
    for (int i = 0; i < 3; ++i) E.p[i] = p[i];
    return ok;
}
template <class Fetch>
MS_HD bool chunk_element(const Ctx& c, int s, int kLo, int kHi, double mu, double delta, Fetch& fetch, const double* refP, const double* refp,
                         ChunkElem& E) {
    if (delta > 0.0) return chunk_element_t<true>(c, s, kLo, kHi, mu, delta, fetch, refP, refp, E);
    return chunk_element_t<false>(c, s, kLo, kHi, mu, delta, fetch, refP, refp, E);
}

// The same for a terminal value given as a difference to the reference value the element was accumulated with:
//     T = (I + dPi W)^{-1} dPi     (dPi symmetric, not necessarily definite; 3x3 elimination with row pivoting, then symmetrised)
// The closer the reference is to the value function (the chunk-end values of the previous interior-point iteration are used),
// the smaller the correction and the better conditioned the recursion that produced the element.
MS_HD bool chunk_apply_diff(const ChunkElem& E, const double dPi[3][3], const double dpi[3], double P[3][3], double p[3]) {
    double W[3][3];
    sym_to_full(E.W, W);
    // augmented rows [I + dPi W | dPi]
    double a0[6], a1[6], a2[6];
    {
        double* rows[3] = {a0, a1, a2};
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                rows[i][j] = (i == j ? 1.0 : 0.0) + dPi[i][0] * W[0][j] + dPi[i][1] * W[1][j] + dPi[i][2] * W[2][j];
                rows[i][3 + j] = dPi[i][j];
            }
    }
#define MS_ROWSWAP(x, y) { for (int q = 0; q < 6; ++q) { const double t_ = x[q]; x[q] = y[q]; y[q] = t_; } }
    // column 0
    if (fabs(a1[0]) > fabs(a0[0])) MS_ROWSWAP(a0, a1)
    if (fabs(a2[0]) > fabs(a0[0])) MS_ROWSWAP(a0, a2)
    if (!(fabs(a0[0]) > 1e-300) || !isfinite(a0[0])) return false;
    { const double r = rcp(a0[0]); const double f1 = a1[0] * r, f2 = a2[0] * r;
      for (int q = 1; q < 6; ++q) { a1[q] -= f1 * a0[q]; a2[q] -= f2 * a0[q]; } }
    // column 1
    if (fabs(a2[1]) > fabs(a1[1])) MS_ROWSWAP(a1, a2)
    if (!(fabs(a1[1]) > 1e-300) || !isfinite(a1[1])) return false;
    { const double r = rcp(a1[1]); const double f2 = a2[1] * r;
      for (int q = 2; q < 6; ++q) a2[q] -= f2 * a1[q]; }
    if (!(fabs(a2[2]) > 1e-300) || !isfinite(a2[2])) return false;
#undef MS_ROWSWAP
    // back substitution for the three right-hand sides
    double T[3][3];
    {
        const double r2 = rcp(a2[2]), r1 = rcp(a1[1]), r0 = rcp(a0[0]);
        for (int j = 0; j < 3; ++j) {
            const double x2 = a2[3 + j] * r2;
            const double x1 = (a1[3 + j] - a1[2] * x2) * r1;
            const double x0 = (a0[3 + j] - a0[1] * x1 - a0[2] * x2) * r0;
            T[0][j] = x0; T[1][j] = x1; T[2][j] = x2;
        }
    }
    for (int i = 0; i < 3; ++i)
        for (int j = i + 1; j < 3; ++j) { const double x = 0.5 * (T[i][j] + T[j][i]); T[i][j] = x; T[j][i] = x; }
    // P = Pbar + Phi' T Phi ;  p = pbar + Phi' (dpi + T (phi - W dpi))
    double TP[3][3], w[3], u[3];
    for (int i = 0; i < 3; ++i) {
        w[i] = E.phi[i] - (W[i][0] * dpi[0] + W[i][1] * dpi[1] + W[i][2] * dpi[2]);
        for (int j = 0; j < 3; ++j) TP[i][j] = T[i][0] * E.Phi[j] + T[i][1] * E.Phi[3 + j] + T[i][2] * E.Phi[6 + j];
    }
    for (int i = 0; i < 3; ++i) u[i] = dpi[i] + T[i][0] * w[0] + T[i][1] * w[1] + T[i][2] * w[2];
    double Pb[3][3];
    sym_to_full(E.P, Pb);
    for (int i = 0; i < 3; ++i) {
        p[i] = E.p[i] + (E.Phi[i] * u[0] + E.Phi[3 + i] * u[1] + E.Phi[6 + i] * u[2]);
        for (int j = i; j < 3; ++j) {
            const double x = Pb[i][j] + (E.Phi[i] * TP[0][j] + E.Phi[3 + i] * TP[1][j] + E.Phi[6 + i] * TP[2][j]);
            P[i][j] = x; P[j][i] = x;
        }
    }
    for (int i = 0; i < 3; ++i) { if (!isfinite(p[i])) return false; for (int j = 0; j < 3; ++j) if (!isfinite(P[i][j])) return false; }
    return true;
}

// Value function at the start of a chunk from the one at its end.  Pi = L L' (Cholesky, semi-definite tolerant: a pivot that is
// zero or negative within rounding gives a zero column), T = L (I + L' W L)^{-1} L' with a second Cholesky factorisation of
// the 3x3 matrix I + L' W L >= I.  Returns false when Pi is clearly not positive semi-definite (or not finite).
MS_HD bool chunk_apply(const ChunkElem& E, const double Pi[3][3], const double pi[3], double P[3][3], double p[3]) {
    const double big = fmax(fmax(fabs(Pi[0][0]), fabs(Pi[1][1])), fabs(Pi[2][2]));
    if (!isfinite(big)) return false;
    const double tiny = 1e-13 * big, neg = -1e-7 * big;
    double L[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    {
        const double d0 = Pi[0][0];
        if (d0 < neg) return false;
        if (d0 > tiny) { const double r = rcp(sqrt(d0)); L[0][0] = d0 * r; L[1][0] = Pi[1][0] * r; L[2][0] = Pi[2][0] * r; }
        const double d1 = Pi[1][1] - L[1][0] * L[1][0];
        if (d1 < neg) return false;
        if (d1 > tiny) { const double r = rcp(sqrt(d1)); L[1][1] = d1 * r; L[2][1] = (Pi[2][1] - L[2][0] * L[1][0]) * r; }
        const double d2 = Pi[2][2] - L[2][0] * L[2][0] - L[2][1] * L[2][1];
        if (d2 < neg) return false;
        if (d2 > tiny) L[2][2] = sqrt(d2);
    }
    double W[3][3];
    sym_to_full(E.W, W);
    // G = I + L' W L
    double WL[3][3], G[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) WL[i][j] = W[i][0] * L[0][j] + W[i][1] * L[1][j] + W[i][2] * L[2][j];
    for (int i = 0; i < 3; ++i)
        for (int j = i; j < 3; ++j) G[i][j] = (i == j ? 1.0 : 0.0) + L[0][i] * WL[0][j] + L[1][i] * WL[1][j] + L[2][i] * WL[2][j];
    // G = C C' (lower), reciprocals of the diagonal
    const double c00 = sqrt(G[0][0]), i00 = rcp(c00);
    const double c10 = G[0][1] * i00, c20 = G[0][2] * i00;
    const double e1 = G[1][1] - c10 * c10;
    if (!(e1 > 0.0)) return false;
    const double c11 = sqrt(e1), i11 = rcp(c11);
    const double c21 = (G[1][2] - c20 * c10) * i11;
    const double e2 = G[2][2] - c20 * c20 - c21 * c21;
    if (!(e2 > 0.0)) return false;
    const double i22 = rcp(sqrt(e2));
    // U = C^{-1} L'  (column j of L' is row j of L)
    double U[3][3];
    for (int j = 0; j < 3; ++j) {
        const double r0 = L[j][0], r1 = L[j][1], r2 = L[j][2];
        const double y0 = r0 * i00, y1 = (r1 - c10 * y0) * i11, y2 = (r2 - c20 * y0 - c21 * y1) * i22;
        U[0][j] = y0; U[1][j] = y1; U[2][j] = y2;
    }
    // V = U Phi ;  P = Pbar + V'V ;  p = pbar + Phi' pi + V' U (phi - W pi)
    double V[3][3], w[3], u[3];
    for (int i = 0; i < 3; ++i) {
        w[i] = E.phi[i] - (W[i][0] * pi[0] + W[i][1] * pi[1] + W[i][2] * pi[2]);
        for (int j = 0; j < 3; ++j) V[i][j] = U[i][0] * E.Phi[j] + U[i][1] * E.Phi[3 + j] + U[i][2] * E.Phi[6 + j];
    }
    for (int i = 0; i < 3; ++i) u[i] = U[i][0] * w[0] + U[i][1] * w[1] + U[i][2] * w[2];
    double Pb[3][3];
    sym_to_full(E.P, Pb);
    for (int i = 0; i < 3; ++i) {
        p[i] = E.p[i] + (E.Phi[i] * pi[0] + E.Phi[3 + i] * pi[1] + E.Phi[6 + i] * pi[2]) + (V[0][i] * u[0] + V[1][i] * u[1] + V[2][i] * u[2]);
        for (int j = i; j < 3; ++j) {
            const double x = Pb[i][j] + (V[0][i] * V[0][j] + V[1][i] * V[1][j] + V[2][i] * V[2][j]);
            P[i][j] = x; P[j][i] = x;
        }
    }
    return true;
}

// ---- host emulation of the chunked driver (tests/hostsim): the lanes of one instance in a loop ----------------------------
#if !defined(__CUDACC__)
struct PitDiag { double devP, devp; long applies, applyFails, elemFails; };
inline PitDiag& pit_diag() { static PitDiag d{0, 0, 0, 0, 0}; return d; }
#endif
template <class FetchB, class FetchF>
inline bool pit_chunks_direction_emulated(const Ctx& c, int s, int N, int G, double mu, double delta, FetchB& fb, FetchF& ff, bool& scanFailed) {
    scanFailed = false;
    ChunkElem E[64];
    double Pe[64][3][3], pe[64][3];
    Aff T[64];
    int kLo[64], kHi[64];
    for (int l = 0; l < G; ++l) pit_chunk(N, G, l, kLo[l], kHi[l]);
    // reference terminal values: what the previous factorisation left at the chunk-end nodes (zero before the first one)
    double refP[64][6], refp[64][3];
    const bool haveRef = c.I(SI_FACT, s) != 0;
    for (int l = 1; l < G - 1; ++l) {
        for (int i = 0; i < 6; ++i) refP[l][i] = haveRef ? c.W(WS_RIC + RIC_P + i, kHi[l], s) : 0.0;
        for (int i = 0; i < 3; ++i) refp[l][i] = haveRef ? c.W(WS_RIC + RIC_PV + i, kHi[l], s) : 0.0;
    }
    // phase A
    for (int l = 1; l < G - 1; ++l)
        if (!chunk_element(c, s, kLo[l], kHi[l], mu, delta, fb, refP[l], refp[l], E[l])) { scanFailed = true; return false; }
    {   // last chunk: ordinary recursion from the terminal node, factors kept
        double P[3][3], p[3];
        terminal_value(c, s, N, mu, delta, P, p);
        aff_identity(T[G - 1]);
        if (!riccati_backward_range(c, s, N, kLo[G - 1], N, mu, delta, fb, P, p, T[G - 1].M, T[G - 1].m)) return false;
        if (G >= 2) { for (int i = 0; i < 3; ++i) { pe[G - 2][i] = p[i]; for (int j = 0; j < 3; ++j) Pe[G - 2][i][j] = P[i][j]; } }
    }
    // chain of value functions at the chunk ends
    for (int l = G - 3; l >= 0; --l) {
        double dP[3][3], dp[3], R[3][3];
        sym_to_full(refP[l + 1], R);
        for (int i = 0; i < 3; ++i) { dp[i] = pe[l + 1][i] - refp[l + 1][i]; for (int j = 0; j < 3; ++j) dP[i][j] = Pe[l + 1][i][j] - R[i][j]; }
        if (!chunk_apply_diff(E[l + 1], dP, dp, Pe[l], pe[l])) { scanFailed = true; return false; }
    }
    // phase C
    for (int l = 0; l < G - 1; ++l) {
        aff_identity(T[l]);
        double P[3][3], p[3];
        for (int i = 0; i < 3; ++i) { p[i] = pe[l][i]; for (int j = 0; j < 3; ++j) P[i][j] = Pe[l][i][j]; }
        if (!riccati_backward_range(c, s, N, kLo[l], kHi[l], mu, delta, fb, P, p, T[l].M, T[l].m)) return false;
    }
#if !defined(__CUDACC__)
    if (getenv("HOSTSIM_PIT_DIAG") && atoi(getenv("HOSTSIM_PIT_DIAG")) >= 3) {
        for (int l = 0; l < G - 1; ++l) {
            const int k = kHi[l];
            printf("   chunk %2d end node %3d  P chain/rec:", l, k);
            const int ij[6][2] = {{0,0},{0,1},{0,2},{1,1},{1,2},{2,2}};
            for (int i = 0; i < 6; ++i) printf(" %.3e|%.1e", Pe[l][ij[i][0]][ij[i][1]], Pe[l][ij[i][0]][ij[i][1]] - c.W(WS_RIC + RIC_P + i, k, s));
            printf("   p:");
            for (int i = 0; i < 3; ++i) printf(" %.3e|%.1e", pe[l][i], pe[l][i] - c.W(WS_RIC + RIC_PV + i, k, s));
            printf("\n");
        }
    }
#endif
    // chain of chunk-start states, forward sweeps
    c.W(WS_ST + ST_T, 0, s) = 0.0;
    c.W(WS_ST + ST_B, 0, s) = 0.0;
    double dx[3] = {0.0, 0.0, 0.0};
    for (int l = 0; l < G; ++l) {
        double d[3] = {dx[0], dx[1], dx[2]};
        riccati_forward_range(c, s, N, kLo[l], (l == G - 1) ? N : kHi[l], mu, delta, ff, d);
        double nx[3];
        for (int i = 0; i < 3; ++i) nx[i] = T[l].m[i] + T[l].M[3 * i] * dx[0] + T[l].M[3 * i + 1] * dx[1] + T[l].M[3 * i + 2] * dx[2];
#if !defined(__CUDACC__)
        if (getenv("HOSTSIM_PIT_DIAG") && atoi(getenv("HOSTSIM_PIT_DIAG")) >= 3 && l < G - 1)
            printf("   fwd chunk %2d: chain (%.3e %.3e %.3e) in-chunk-minus-chain (%.1e %.1e %.1e)  m (%.2e %.2e %.2e)\n", l, nx[0], nx[1], nx[2], d[0]-nx[0], d[1]-nx[1], d[2]-nx[2], T[l].m[0], T[l].m[1], T[l].m[2]);
#endif
        for (int i = 0; i < 3; ++i) dx[i] = nx[i];
    }
    return true;
}

// per-instance driver with the inertia-correction ladder (host emulation).  A failure of the element algebra (zero-terminal
// recursion not positive definite, value function not positive semi-definite) sends the instance to the sequential sweeps.
template <class FetchB, class FetchF>
inline void inst_step_pit_chunks_emulated(const Ctx& c, int s, int G, FetchB& fb, FetchF& ff, long* fallbacks) {
    const Config& g = c.cfg;
    if (s >= g.nInst || c.I(SI_PHASE, s) != PH_FACTOR) return;
    const int N = c.I(SI_N_INT, s);
    const double mu = c.D(SD_MU, s);
    double delta = 0.0;
    const double dlast = c.D(SD_DELTA_LAST, s);
    bool ok = false;
    for (int tries = 0; tries < 40; ++tries) {
        count_cells(c, 2, N);
        bool scanFailed = false;
        if (pit_chunks_direction_emulated(c, s, N, G, mu, delta, fb, ff, scanFailed)) { ok = true; break; }
        if (scanFailed) { if (fallbacks) *fallbacks += 1; inst_step(c, s, fb, ff); return; }
        c.I(SI_NREG, s) += 1;
        if (delta == 0.0) delta = (dlast == 0.0) ? 1e-4 : fmax(1e-20, dlast / 3.0);
        else delta *= (dlast == 0.0) ? 100.0 : 8.0;
        if (delta > 1e40) break;
    }
    if (!ok) { finish(c, s, ST_STEP_FAILED); return; }
    if (delta > 0.0) c.D(SD_DELTA_LAST, s) = delta;
    c.D(SD_DELTA, s) = delta;
    count_cells(c, 3, N);
    c.I(SI_FACT, s) = 1;
    c.I(SI_PHASE, s) = PH_STEPPED;
}

}  // namespace mseetc
