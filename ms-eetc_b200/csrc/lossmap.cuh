// Dynamic loss map on the device: the two loss-epigraph rows  s - PLtr(Fel,vMid)/vMid,  s - PLrgb(Fel,vMid)/vMid
// (reference ocp.py:221-229) for train.powerLosses = efficiency.totalLossesFunction(...).
//
// Replaces, with exact first and second derivatives w.r.t. (Fel, b_k, b_{k+1}):
//   efficiency.py:7-12     forceToLoad (force-limited below the turning speed, power-limited above)
//   efficiency.py:23-51    createSpline: cubic not-a-knot tensor B-spline of the measured motor losses, speed clipped
//                          to the measured range (zero slope outside), value 0 outside the load range
//   efficiency.py:101-141  gear + motor + auxiliaries + transformer losses, zero where the motor map is zero
//   utils.py:197-220       splitLosses: the inactive half-plane is the tangent at f = +-1e-10, i.e.
//                          alpha(v) f + beta(v) with alpha = dPL/df(+-1e-10, v), beta = PL(0, v)
//   train.py:216           specific form PL(f*M, v)/M
// Jet variables: x0 = v (mid-point speed), x1 = specific force.  alpha(v), beta(v) are jets in v only; their second
// derivative needs the mixed third derivative d3S/dl dv2 of the spline, which the B-spline basis provides.
#pragma once
#include "jet.cuh"

namespace mseetc {

#define MS_TRAFO_R 10.0
#define MS_TRAFO_V 15000.0
#define MS_SPLIT_TOL 1e-10

struct LossMapDev {
    const double* tl;     // load knots  (nl + 4)
    const double* tv;     // speed knots (nv + 4)
    const double* coef;   // [nl][nv]
    int nl, nv;
};

// non-zero cubic B-spline basis functions at x and their first two derivatives (The NURBS Book, A2.3, p = 3)
MS_HD int bspline_ders(const double* t, int n, double x, double N[3][4]) {
    int i = 3;
    while (i < n - 1 && x >= t[i + 1]) ++i;          // span: t[i] <= x < t[i+1], last span closed
    double ndu[4][4], left[4], right[4];
    ndu[0][0] = 1.0;
    for (int j = 1; j <= 3; ++j) {
        left[j] = x - t[i + 1 - j];
        right[j] = t[i + j] - x;
        double saved = 0.0;
        for (int r = 0; r < j; ++r) {
            ndu[j][r] = right[r + 1] + left[j - r];
            const double tmp = ndu[r][j - 1] / ndu[j][r];
            ndu[r][j] = saved + right[r + 1] * tmp;
            saved = left[j - r] * tmp;
        }
        ndu[j][j] = saved;
    }
    for (int j = 0; j <= 3; ++j) N[0][j] = ndu[j][3];
    for (int r = 0; r <= 3; ++r) {
        double a[2][4];
        int s1 = 0, s2 = 1;
        a[0][0] = 1.0;
        for (int k = 1; k <= 2; ++k) {
            double d = 0.0;
            const int rk = r - k, pk = 3 - k;
            if (r >= k) { a[s2][0] = a[s1][0] / ndu[pk + 1][rk]; d = a[s2][0] * ndu[rk][pk]; }
            const int j1 = (rk >= -1) ? 1 : -rk;
            const int j2 = (r - 1 <= pk) ? k - 1 : 3 - r;
            for (int j = j1; j <= j2; ++j) {
                a[s2][j] = (a[s1][j] - a[s1][j - 1]) / ndu[pk + 1][rk + j];
                d += a[s2][j] * ndu[rk + j][pk];
            }
            if (r <= pk) { a[s2][k] = -a[s1][k - 1] / ndu[pk + 1][r]; d += a[s2][k] * ndu[r][pk]; }
            N[k][r] = d;
            const int tmp = s1; s1 = s2; s2 = tmp;
        }
    }
    for (int j = 0; j <= 3; ++j) { N[1][j] *= 3.0; N[2][j] *= 6.0; }
    return i;
}

// D[a][b] = d^(a+b) S / dl^a dv^b, a,b = 0..2.  Returns false (all zero) outside the grid.
MS_HD bool spline_eval(const LossMapDev& m, double l, double v, double D[3][3]) {
    for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) D[a][b] = 0.0;
    if (!(l >= m.tl[0] && l <= m.tl[m.nl + 3] && v >= m.tv[0] && v <= m.tv[m.nv + 3])) return false;
    double Nl[3][4], Nv[3][4];
    const int il = bspline_ders(m.tl, m.nl, l, Nl);
    const int iv = bspline_ders(m.tv, m.nv, v, Nv);
    for (int p = 0; p < 4; ++p) {
        double row[3] = {0.0, 0.0, 0.0};       // sum over the speed direction for value / d/dv / d2/dv2
        for (int q = 0; q < 4; ++q) {
            const double cf = m.coef[(il - 3 + p) * m.nv + (iv - 3 + q)];
            row[0] += cf * Nv[0][q]; row[1] += cf * Nv[1][q]; row[2] += cf * Nv[2][q];
        }
        for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) D[a][b] += Nl[a][p] * row[b];
    }
    return true;
}

// composition S(l, w) of jets l, w with the partials of S up to second order
MS_HD Jet2 jchain2(const Jet2& l, const Jet2& w, double S, double Sl, double Sw, double Sll, double Slw, double Sww) {
    Jet2 r;
    r.v = S;
    r.g0 = Sl * l.g0 + Sw * w.g0;
    r.g1 = Sl * l.g1 + Sw * w.g1;
    r.h00 = Sl * l.h00 + Sw * w.h00 + Sll * l.g0 * l.g0 + 2.0 * Slw * l.g0 * w.g0 + Sww * w.g0 * w.g0;
    r.h01 = Sl * l.h01 + Sw * w.h01 + Sll * l.g0 * l.g1 + Slw * (l.g0 * w.g1 + l.g1 * w.g0) + Sww * w.g0 * w.g1;
    r.h11 = Sl * l.h11 + Sw * w.h11 + Sll * l.g1 * l.g1 + 2.0 * Slw * l.g1 * w.g1 + Sww * w.g1 * w.g1;
    return r;
}

struct LossPar {       // per-instance numbers of the map
    double M;          // mass*rho: specific <-> absolute
    double aux, cgT, cgB;   // auxiliaries [W]; gear loss factors (1-eta)/eta (traction), (1-eta) (braking)
    double fMax, pMax; // of the measured drive (efficiency.py:64-65), absolute
    double scale;      // table scale (parameter studies); 1 for the reference map
};

// total losses / v as jets in (v, fs) on the side given by `traction`; full map (no tangent extension)
MS_HD Jet2 loss_full(const LossMapDev& m, const LossPar& p, const Jet2& v, const Jet2& fs, bool traction, bool clampLoad = false) {
    const double vMin = m.tv[0], vMax = m.tv[m.nv + 3];
    Jet2 vc = v;
    if (v.v < vMin) vc = jconst(vMin);
    else if (v.v > vMax) vc = jconst(vMax);
    const Jet2 fa = p.M * fs;                               // absolute force
    const Jet2 absf = traction ? fa : (-1.0) * fa;
    const double tp = p.pMax / p.fMax;
    Jet2 load = (vc.v <= tp) ? (100.0 / p.fMax) * absf : (100.0 / p.pMax) * (absf * vc);
    // clampLoad (time-domain loss integration, see power_loss_jets in core.cuh): a stage point that overshoots the upper load edge
    // of the grid -- the power hyperbola, where the NLP's power rows are active -- by the integration error is evaluated at the edge
    // (overshoots up to 0.1 % only: further out the map is zero as in the reference)
    if (clampLoad && load.v > m.tl[m.nl + 3] && load.v <= 1.001 * m.tl[m.nl + 3]) load = jconst(m.tl[m.nl + 3]);
    double D[3][3];
    const bool inside = spline_eval(m, load.v, vc.v, D);
    if (!inside || !(p.scale * D[0][0] > 0.0)) return jconst(0.0);      // efficiency.py:137
    const Jet2 mot = p.scale * jchain2(load, vc, D[0][0], D[1][0], D[0][1], D[2][0], D[1][1], D[0][2]);
    const Jet2 pw = traction ? fa * v : (-1.0) * (fa * v);  // power at the wheel (positive on both sides)
    const Jet2 gear = (traction ? p.cgT : p.cgB) * pw;
    Jet2 arg;                                               // V^2 -+ 4 R Pm
    if (traction) arg = (-4.0 * MS_TRAFO_R) * (pw + gear + mot + p.aux) + MS_TRAFO_V * MS_TRAFO_V;
    else arg = (4.0 * MS_TRAFO_R) * (pw - gear - mot - p.aux) + MS_TRAFO_V * MS_TRAFO_V;
    const Jet2 dv = (-1.0) * jsqrt(arg) + MS_TRAFO_V;       // V - sqrt(.)
    const Jet2 trafo = (1.0 / (4.0 * MS_TRAFO_R)) * (dv * dv);
    const Jet2 total = gear + mot + trafo + p.aux;
    return ((1.0 / p.M) * total) * jrecip(v);
}

// tangent extension (alpha(v) fs + beta(v)) / v of the side given by `traction`, used on the other half-plane
MS_HD Jet2 loss_tangent(const LossMapDev& m, const LossPar& p, const Jet2& v, const Jet2& fs, bool traction) {
    const double vMin = m.tv[0], vMax = m.tv[m.nv + 3];
    Jet2 vc = jvar0(v.v);
    Jet2 vj = jvar0(v.v);                                   // jets in v only
    if (v.v < vMin) vc = jconst(vMin);
    else if (v.v > vMax) vc = jconst(vMax);
    const double tp = p.pMax / p.fMax;
    const bool below = vc.v <= tp;
    const double f0 = MS_SPLIT_TOL * p.M;                   // |force| at which the slope is taken
    const double l0 = below ? 100.0 * f0 / p.fMax : 100.0 * f0 * vc.v / p.pMax;
    const Jet2 loadF = below ? jconst(100.0 / p.fMax) : (100.0 / p.pMax) * vc;     // d load / d |f|
    double D0[3][3], Dz[3][3];
    const bool in0 = spline_eval(m, l0, vc.v, D0);
    const bool inz = spline_eval(m, 0.0, vc.v, Dz);
    const double vcp = vc.g0;                               // 1 inside the measured speed range, 0 outside
    // ---- beta(v) = PL(0, v)/M   (f = 0 is on the traction side: f >= 0)
    Jet2 beta = jconst(0.0);
    if (inz && p.scale * Dz[0][0] > 0.0) {
        const Jet2 mz = p.scale * Jet2{Dz[0][0], Dz[0][1] * vcp, 0.0, Dz[0][2] * vcp * vcp, 0.0, 0.0};
        const Jet2 arg = (-4.0 * MS_TRAFO_R) * (mz + p.aux) + MS_TRAFO_V * MS_TRAFO_V;
        const Jet2 dvz = (-1.0) * jsqrt(arg) + MS_TRAFO_V;
        beta = (1.0 / p.M) * (mz + (1.0 / (4.0 * MS_TRAFO_R)) * (dvz * dvz) + p.aux);
    }
    // ---- alpha(v) = d PL / d f at f = +-f0  (absolute force; equals the slope w.r.t. the specific force of PL/M)
    Jet2 alpha = jconst(0.0);
    if (in0 && p.scale * D0[0][0] > 0.0) {
        const double sgn = traction ? 1.0 : -1.0;          // d|f|/df
        const Jet2 m0 = p.scale * Jet2{D0[0][0], D0[0][1] * vcp, 0.0, D0[0][2] * vcp * vcp, 0.0, 0.0};
        const Jet2 ml = p.scale * Jet2{D0[1][0], D0[1][1] * vcp, 0.0, D0[1][2] * vcp * vcp, 0.0, 0.0};   // dS/dl along v
        const Jet2 motF = sgn * (ml * loadF);
        Jet2 gearF, pmF, arg;
        if (traction) {
            gearF = p.cgT * vj;                              // d/df of cgT f v
            pmF = vj + gearF + motF;                         // d/df of (f v + gear + mot + aux)
            arg = (-4.0 * MS_TRAFO_R) * ((f0 * (1.0 + p.cgT)) * vj + m0 + p.aux) + MS_TRAFO_V * MS_TRAFO_V;
        } else {
            gearF = (-p.cgB) * vj;                           // d/df of cgB (-f v)
            pmF = (-1.0) * vj - gearF - motF;                // d/df of (-f v - gear - mot - aux)
            arg = (4.0 * MS_TRAFO_R) * ((f0 * (1.0 - p.cgB)) * vj - m0 - p.aux) + MS_TRAFO_V * MS_TRAFO_V;
        }
        const Jet2 sq = jsqrt(arg);
        const Jet2 ratio = ((-1.0) * sq + MS_TRAFO_V) * jrecip(sq);      // (V - sqrt D)/sqrt D
        const Jet2 trafoF = traction ? ratio * pmF : (-1.0) * (ratio * pmF);
        alpha = gearF + motF + trafoF;
    }
    return (alpha * fs + beta) * jrecip(v);
}

struct LossRow {       // G(Fel, b0, b1) with gradient and Hessian; the row is  s - G
    double v, gF, g0, g1, hFF, hF0, hF1, h00, h01, h11;
};

MS_HD void loss_rows_dynamic(const LossMapDev& m, const LossPar& p, double fel, double b0, double b1, LossRow& tr, LossRow& rg) {
    double s0, s1, is0, is1;
    sqrt_inv(b0, s0, is0);
    sqrt_inv(b1, s1, is1);
    const double vm = 0.5 * (s0 + s1);
    const double v_0 = 0.25 * is0, v_1 = 0.25 * is1;
    const double v_00 = -0.5 * v_0 * rcp(b0), v_11 = -0.5 * v_1 * rcp(b1);
    const Jet2 v = jvar0(vm), fs = jvar1(fel);
    const bool pos = fel >= 0.0;
    const Jet2 qt = pos ? loss_full(m, p, v, fs, true) : loss_tangent(m, p, v, fs, true);
    const Jet2 qr = pos ? loss_tangent(m, p, v, fs, false) : loss_full(m, p, v, fs, false);
    const Jet2* q[2] = {&qt, &qr};
    LossRow* o[2] = {&tr, &rg};
    for (int a = 0; a < 2; ++a) {
        const Jet2& j = *q[a];
        LossRow& r = *o[a];
        r.v = j.v; r.gF = j.g1; r.g0 = j.g0 * v_0; r.g1 = j.g0 * v_1;
        r.hFF = j.h11; r.hF0 = j.h01 * v_0; r.hF1 = j.h01 * v_1;
        r.h00 = j.h00 * v_0 * v_0 + j.g0 * v_00; r.h01 = j.h00 * v_0 * v_1; r.h11 = j.h00 * v_1 * v_1 + j.g0 * v_11;
    }
}

}  // namespace mseetc
