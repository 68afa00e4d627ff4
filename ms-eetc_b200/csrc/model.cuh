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

}  // namespace mseetc
