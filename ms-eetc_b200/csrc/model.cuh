// Shooting-interval device functions: train ODE in the position domain, RK4 ("ERK4+") step.
//
// Replaces, for the RK branch that simulations/config.json selects:
//   mseetc/train.py:251-259   TrainModel ODE  (rolling resistance, db/ds = 2a, dt/ds = 1/v, both * ds)
//   mseetc/train.py:298-301   ca.simpleRK(ode, numSteps, 4)  -> classic RK4, numSteps equal sub-steps
//   mseetc/train.py:324-344   time approximation with numApproxSteps sub-points (b evaluated from b0
//                             for every sub-point, t += 2*ds*dsigma/(v_i+v_{i+1}))
// Templated on the scalar so the same code yields values (double) or exact first+second
// sensitivities w.r.t. (b0, F) (Jet2).
#pragma once
#include "jet.cuh"

namespace mseetc {

struct IntervalCoef {
    double ds;   // interval length [m]                                   (ocp.py:125)
    double c0;   // g*grad/rho + curvRes(kappa)/rho  [m/s^2]              (train.py:252-254)
    double sr0, sr1, sr2;  // specific Davis coefficients                 (train.py:181-183)
};

MS_HD double msqrt(double x) { return sqrt(x); }
MS_HD Jet2 msqrt(const Jet2& x) { return jsqrt(x); }
MS_HD double mrecip(double x) { return rcp(x); }
MS_HD Jet2 mrecip(const Jet2& x) { return jrecip(x); }

// a(b, F) = F - (sr0 + sr1*sqrt(b) + sr2*b) - c0                     (train.py:251,254)
template <class T>
MS_HD T accel(const T& b, const T& F, const IntervalCoef& c) {
    return F - (c.sr1 * msqrt(b) + c.sr2 * b) - (c.sr0 + c.c0);
}

// numSteps classic RK4 steps of db/dsigma = 2*ds*a(b,F) over sigma in [0,h]
template <class T>
MS_HD T rk4_b(const T& b0, const T& F, double h, int numSteps, const IntervalCoef& c) {
    const double dt = (numSteps == 1) ? h : h / numSteps;          // the usual single step costs no division
    const double w = 2.0 * c.ds * dt;
    T b = b0;
    for (int s = 0; s < numSteps; ++s) {
        T k1 = w * accel(b, F, c);
        T k2 = w * accel(b + 0.5 * k1, F, c);
        T k3 = w * accel(b + 0.5 * k2, F, c);
        T k4 = w * accel(b + k3, F, c);
        b = b + (1.0 / 6.0) * (k1 + 2.0 * k2 + 2.0 * k3 + k4);
    }
    return b;
}

// One shooting interval: tau = t1 - t0, phib = b1.
template <class T>
MS_HD void shoot(const T& b0, const T& F, const IntervalCoef& c, int numSteps, int numApprox, T& tau, T& phib) {
    if (numApprox > 0) {
        T vprev = msqrt(b0);
        T acc_t = 0.0 * b0;
        T bf = b0;
        const double wq = (numApprox == 1) ? 2.0 * c.ds : 2.0 * c.ds / numApprox;
        for (int i = 1; i <= numApprox; ++i) {
            bf = rk4_b(b0, F, (i == numApprox) ? 1.0 : (double)i / numApprox, numSteps, c);
            T vnext = msqrt(bf);
            acc_t = acc_t + wq * mrecip(vprev + vnext);
            vprev = vnext;
        }
        tau = acc_t;
        phib = bf;
    } else {
        // RK4 on (t, b):  dt/dsigma = ds/sqrt(b),  db/dsigma = 2*ds*a          (train.py:255-259,298-299)
        const double dt = (numSteps == 1) ? 1.0 : 1.0 / numSteps;
        T b = b0;
        T t = 0.0 * b0;
        for (int s = 0; s < numSteps; ++s) {
            T k1b = (2.0 * c.ds * dt) * accel(b, F, c);
            T k1t = (c.ds * dt) * mrecip(msqrt(b));
            T b2 = b + 0.5 * k1b;
            T k2b = (2.0 * c.ds * dt) * accel(b2, F, c);
            T k2t = (c.ds * dt) * mrecip(msqrt(b2));
            T b3 = b + 0.5 * k2b;
            T k3b = (2.0 * c.ds * dt) * accel(b3, F, c);
            T k3t = (c.ds * dt) * mrecip(msqrt(b3));
            T b4 = b + k3b;
            T k4b = (2.0 * c.ds * dt) * accel(b4, F, c);
            T k4t = (c.ds * dt) * mrecip(msqrt(b4));
            b = b + (1.0 / 6.0) * (k1b + 2.0 * k2b + 2.0 * k3b + k4b);
            t = t + (1.0 / 6.0) * (k1t + 2.0 * k2t + 2.0 * k3t + k4t);
        }
        tau = t;
        phib = b;
    }
}

// ---- implicit Runge-Kutta (collocation) steps                              (train.py:303-310: ca.simpleIRK(ode, numSteps, order,
// collMethod, 'fast_newton')).  CasADi builds the collocation equations on `order` Radau / Legendre points per step and solves them
// with Newton's method from the constant guess; the derivatives it propagates are those of the implicitly defined solution.  A
// collocation method is the implicit Runge-Kutta method with A_ij = int_0^{c_i} l_j, w_j = int_0^1 l_j (l_j: Lagrange basis on the
// points), so the host passes (A, w) and the device solves  b_i = b + dt' * sum_j A_ij f(b_j)  for the stage values:
//   * values by Newton's method in double precision (the time equation is a quadrature once the b_i are known: dt/dsigma does not
//     depend on t), at most `maxNewton` iterations (reference default 10), stopped early at a relative update of 1e-15;
//   * first and second sensitivities w.r.t. (b0, F) by two further Newton passes in jet arithmetic with the converged Jacobian:
//     with the values converged the first pass makes the first-order part exact, the second pass the second-order part (the
//     error of a Newton step is quadratic in the previous error, and a jet whose value and first-order parts vanish squares to 0).
#define MS_IRK_MAXD 9         // OptionsIRK: 1 <= order <= 9 (train.py:503)
struct IrkTab {
    int d;                    // collocation points per step; 0 = explicit RK4 (the RK branch above)
    int maxNewton;
    double A[MS_IRK_MAXD * MS_IRK_MAXD], w[MS_IRK_MAXD];
};

// LU factorisation with partial pivoting of the d x d matrix J (row-major, stride MS_IRK_MAXD), in place
MS_HD void irk_lu(double* J, int* piv, int d) {
    for (int c = 0; c < d; ++c) {
        int p = c;
        for (int r = c + 1; r < d; ++r) if (fabs(J[r * MS_IRK_MAXD + c]) > fabs(J[p * MS_IRK_MAXD + c])) p = r;
        piv[c] = p;
        if (p != c) for (int j = 0; j < d; ++j) { const double t = J[c * MS_IRK_MAXD + j]; J[c * MS_IRK_MAXD + j] = J[p * MS_IRK_MAXD + j]; J[p * MS_IRK_MAXD + j] = t; }
        const double ip = 1.0 / J[c * MS_IRK_MAXD + c];
        for (int r = c + 1; r < d; ++r) {
            const double l = J[r * MS_IRK_MAXD + c] * ip;
            J[r * MS_IRK_MAXD + c] = l;
            for (int j = c + 1; j < d; ++j) J[r * MS_IRK_MAXD + j] -= l * J[c * MS_IRK_MAXD + j];
        }
    }
}
// x <- J^{-1} x with the factors of irk_lu
MS_HD void irk_lu_solve(const double* J, const int* piv, int d, double* x) {
    for (int c = 0; c < d; ++c) {
        if (piv[c] != c) { const double t = x[c]; x[c] = x[piv[c]]; x[piv[c]] = t; }
        for (int r = c + 1; r < d; ++r) x[r] -= J[r * MS_IRK_MAXD + c] * x[c];
    }
    for (int c = d - 1; c >= 0; --c) {
        for (int j = c + 1; j < d; ++j) x[c] -= J[c * MS_IRK_MAXD + j] * x[j];
        x[c] /= J[c * MS_IRK_MAXD + c];
    }
}

// numSteps collocation steps of db/dsigma = 2*ds*a(b,F) over sigma in [0,h]; tq (optional) += int ds/sqrt(b) dsigma by the
// quadrature of the same method
MS_HD Jet2 irk_b(const Jet2& b0, const Jet2& F, double h, int numSteps, const IntervalCoef& c, const IrkTab& K, Jet2* tq) {
    const int d = K.d;
    const double dt = h / numSteps, wb = 2.0 * c.ds * dt, wt = c.ds * dt;
    Jet2 b = b0;
    for (int step = 0; step < numSteps; ++step) {
        double bv[MS_IRK_MAXD], J[MS_IRK_MAXD * MS_IRK_MAXD], rhs[MS_IRK_MAXD];
        int piv[MS_IRK_MAXD];
        for (int i = 0; i < d; ++i) bv[i] = b.v;
        auto jacobian = [&]() {
            for (int i = 0; i < d; ++i)
                for (int j = 0; j < d; ++j) {
                    const double fp = -(0.5 * c.sr1 / sqrt(bv[j]) + c.sr2);              // d a / d b
                    J[i * MS_IRK_MAXD + j] = (i == j ? 1.0 : 0.0) - wb * K.A[i * d + j] * fp;
                }
            irk_lu(J, piv, d);
        };
        for (int it = 0; it < K.maxNewton; ++it) {
            double f[MS_IRK_MAXD];
            for (int j = 0; j < d; ++j) f[j] = accel(bv[j], F.v, c);
            for (int i = 0; i < d; ++i) {
                double acc = 0.0;
                for (int j = 0; j < d; ++j) acc += K.A[i * d + j] * f[j];
                rhs[i] = bv[i] - b.v - wb * acc;
            }
            jacobian();
            irk_lu_solve(J, piv, d, rhs);
            double big = 0.0;
            for (int i = 0; i < d; ++i) { bv[i] -= rhs[i]; big = fmax(big, fabs(rhs[i])); }
            if (big <= 1e-15 * fmax(1.0, fabs(b.v))) break;
        }
        jacobian();
        Jet2 bs[MS_IRK_MAXD], f[MS_IRK_MAXD];
        for (int i = 0; i < d; ++i) bs[i] = jconst(bv[i]);
        for (int pass = 0; pass < 2; ++pass) {
            for (int j = 0; j < d; ++j) f[j] = accel(bs[j], F, c);
            Jet2 G[MS_IRK_MAXD];
            for (int i = 0; i < d; ++i) {
                Jet2 acc = jconst(0.0);
                for (int j = 0; j < d; ++j) acc = acc + K.A[i * d + j] * f[j];
                G[i] = bs[i] - b - wb * acc;
            }
            for (int q = 0; q < 6; ++q) {            // the same linear solve for each of the six components of the jets
                for (int i = 0; i < d; ++i) rhs[i] = (&G[i].v)[q];
                irk_lu_solve(J, piv, d, rhs);
                for (int i = 0; i < d; ++i) (&bs[i].v)[q] -= rhs[i];
            }
        }
        Jet2 incr = jconst(0.0);
        for (int j = 0; j < d; ++j) incr = incr + K.w[j] * accel(bs[j], F, c);
        if (tq) {
            Jet2 q = jconst(0.0);
            for (int j = 0; j < d; ++j) q = q + K.w[j] * mrecip(msqrt(bs[j]));
            *tq = *tq + wt * q;
        }
        b = b + wb * incr;
    }
    return b;
}

// One shooting interval with the collocation integrator: tau = t1 - t0, phib = b1 (same composition as `shoot`).
MS_HD void shoot_irk(const Jet2& b0, const Jet2& F, const IntervalCoef& c, int numSteps, int numApprox, const IrkTab& K, Jet2& tau, Jet2& phib) {
    if (numApprox > 0) {
        Jet2 vprev = msqrt(b0), acc_t = jconst(0.0), bf = b0;
        for (int i = 1; i <= numApprox; ++i) {
            bf = irk_b(b0, F, (double)i / numApprox, numSteps, c, K, nullptr);
            const Jet2 vnext = msqrt(bf);
            acc_t = acc_t + (2.0 * c.ds / numApprox) * mrecip(vprev + vnext);
            vprev = vnext;
        }
        tau = acc_t;
        phib = bf;
    } else {
        tau = jconst(0.0);
        phib = irk_b(b0, F, 1.0, numSteps, c, K, &tau);
    }
}

}  // namespace mseetc
