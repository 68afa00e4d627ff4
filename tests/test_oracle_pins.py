"""Pins of the CPU oracle against every known answer the reference holds for this path."""
import os

import numpy as np
import pytest

from common import FLAT_JSON, SWISS_JSON, fig5_train, fig10_train, oracle_nlp, oracle_solve, virm6

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def test_min_time_pin_figure5():
    """reference simulations/figure5.py:96  minimumTime = 272.4726 s (8.5 km crop, v0 = 1, vN = 100 km/h)."""
    from oracle.problem import load_track
    train = fig5_train()
    train.losses = ('none',)
    track = load_track(FLAT_JSON).crop(positionEnd=8500)
    nlp = oracle_nlp(train, track, 300, energy=False)
    r = oracle_solve(nlp, 354.0, v0=1.0, vN=100 / 3.6)
    assert r.success and r.kkt <= 1e-8
    tN = r.x[nlp.iT[-1]]
    assert abs(tN - 272.4726) < 1e-4          # 7-digit constant; we get 272.472544 (2e-7 relative)
    # independent of the upper bound on the trip time
    r2 = oracle_solve(nlp, 600.0, v0=1.0, vN=100 / 3.6)
    assert r2.success and abs(r2.x[nlp.iT[-1]] - tN) < 1e-6


def test_gpops_energy_pin_figure10():
    """gpops/00_var_speed_limit_100_GPOPS{I,II}.csv: 440.1414723 / 440.1406149 kWh (continuous-time optimum of
    the figure-10 problem).  The N-interval multiple-shooting optimum converges to it at O(1/N)."""
    import csv
    from oracle.problem import load_track
    energies = []
    for name in ('00_var_speed_limit_100_GPOPSI.csv', '00_var_speed_limit_100_GPOPSII.csv'):
        with open(os.path.join(GOLDEN, name)) as fh:
            rows = list(csv.reader(fh))
        energies.append(float(rows[1][6]))
    assert energies == [440.1414723, 440.1406149]
    track = load_track(FLAT_JSON)
    gaps = []
    for N in (150, 300, 600):
        r = oracle_solve(oracle_nlp(fig10_train(), track, N), 1541.0)
        assert r.success
        gaps.append(r.f - energies[1])
    assert 0 < gaps[2] < gaps[1] < gaps[0]
    assert gaps[1] / energies[1] < 3e-3        # 0.21 % at N = 300
    assert gaps[2] < 0.6 * gaps[1]            # first-order convergence towards the GPOPS optimum
    # velocity profile against the GPOPS-II trajectory
    with open(os.path.join(GOLDEN, '00_var_speed_limit_100_GPOPSII.csv')) as fh:
        rows = list(csv.reader(fh))[1:]
    gp = np.array([[float(x) for x in row[:3]] for row in rows])
    nlp = oracle_nlp(fig10_train(), track, 600)
    r = oracle_solve(nlp, 1541.0)
    v = np.sqrt(np.interp(gp[:, 1], nlp.pos, r.x[nlp.iB]))   # v^2 is close to piecewise linear in position
    assert np.abs(v - gp[:, 2]).max() < 0.8   # m/s, discretisation-level agreement
    assert np.abs(v - gp[:, 2]).mean() < 0.1


def test_ode_constants_figure4():
    """reference simulations/figure4.py:22-23: braking at -0.5 N/kg over 100 m from 36.61894 / 37.95880 km/h ends
    at 1 / 10 km/h (pins the ODE restatement, train.py:251-259)."""
    from scipy.integrate import solve_ivp
    t = virm6()
    M = t.mass * t.rho
    sr = (t.r0 / M, t.r1 / M, t.r2 / M)
    for v0_kmh, vend_kmh in ((36.61894, 1.0), (37.95880, 10.0)):
        rhs = lambda s, b: 2 * (-0.5 - (sr[0] + sr[1] * np.sqrt(b) + sr[2] * b))
        sol = solve_ivp(rhs, (0, 100), [(v0_kmh / 3.6) ** 2], rtol=1e-12, atol=1e-14)
        assert abs(np.sqrt(sol.y[0, -1]) * 3.6 - vend_kmh) < 2e-3


def test_erk4_time_rule_matches_sympy_and_fine_integration():
    "ERK4+ (train.py:324-344): b by one RK4 step, t by the average-speed rule; compare with a fine reference."
    from scipy.integrate import solve_ivp
    from oracle.nlp import _stage_functions
    t = virm6()
    M = t.mass * t.rho
    sr = (t.r0 / M, t.r1 / M, t.r2 / M)
    names, fn = _stage_functions(True, 1, 1, 'static')
    b0, F, ds, c0 = (40 / 3.6) ** 2, 0.3, 150.0, -0.015 * 9.81 / 1.06
    out = fn(b0, F, 0.0, 200.0, ds, c0, *sr, 0.1, 0.3)
    per = 15
    tau = -out[names.index('ct') * per]
    phib = 200.0 - out[names.index('cb') * per]
    sol = solve_ivp(lambda s, y: [1 / np.sqrt(y[1]), 2 * (F - (sr[0] + sr[1] * np.sqrt(y[1]) + sr[2] * y[1]) - c0)],
                    (0, ds), [0.0, b0], rtol=1e-12, atol=1e-14)
    assert abs(phib - sol.y[1, -1]) < 1e-3      # one RK4 step over 150 m
    assert abs(tau - sol.y[0, -1]) < 5e-3      # the trapezoid-in-1/v rule is a low-order approximation


def test_unit_test_properties_curvature():
    """reference unitTests/curvatureResistance/curvatureResistance.py:94-201 re-expressed on the oracle."""
    from oracle.problem import load_track
    g, rho, K = 9.81, 1.06, 1 / 300
    fcurv = g * 0.5 * K / ((1 - 30 * K) * rho)
    # minimum-time: shifting the force limits by the curve resistance reproduces the straight-track speed profile
    tr = virm6(forceMinPn=0, powerMax=None, powerMin=None)
    tr.losses = ('none',)
    straight = load_track(FLAT_JSON).crop(positionEnd=3475)
    curved = load_track(FLAT_JSON, constant_curvature=K).crop(positionEnd=3475)
    nlp0 = oracle_nlp(tr, straight, 300, energy=False)
    r0 = oracle_solve(nlp0, 180.0)
    tr2 = tr.copy()
    tr2.forceMax = tr.forceMax + fcurv * tr.mass * tr.rho
    tr2.forceMin = tr.forceMin + fcurv * tr.mass * tr.rho
    nlp1 = oracle_nlp(tr2, curved, 300, energy=False)
    r1 = oracle_solve(nlp1, 180.0)
    assert r0.success and r1.success
    v0, v1 = np.sqrt(r0.x[nlp0.iB]), np.sqrt(r1.x[nlp1.iB])
    assert np.all(np.abs((v0 - v1) / v0) <= 1e-3)
    # minimum-energy: extra mechanical energy on the curved track = curve-resistance work (within 5 %)
    tr = virm6(forceMinPn=0)
    work = fcurv * tr.rho * tr.mass * 3475 / 3.6e6
    for losses in (('none',), ('static', 0.73, 0.73)):
        tr.losses = losses
        e = []
        for track in (straight, curved):
            nlp = oracle_nlp(tr, track, 300)
            r = oracle_solve(nlp, 200.0)
            assert r.success
            u = nlp.unpack(r.x)
            e.append(np.sum(nlp.ds * u['Fel']) * nlp.M / 3.6e6)    # mechanical energy at the wheel [kWh]
        assert abs((e[1] - e[0]) - work) / work <= 5e-2


def test_loss_map_pins_figure3():
    """reference simulations/figure3.py:113-115: max static(eta=0.73) loss / max dynamic loss on the plotted grid must lie in
    [0.99, 1.01].  Checked for the oracle's torch restatement of efficiency.py AND for the product's efficiency.py."""
    import torch
    from oracle.lossmap import DynamicLossMap
    from mseetc.train import Train
    from mseetc.efficiency import totalLossesFunction, loadToForce, motorLossesFunction
    tr = Train(config={'id': 'NL_Intercity_VIRM6'})
    fun = totalLossesFunction(tr, auxiliaries=27000, etaGear=0.96)
    assert abs(tr.powerMax - 213900 * 52.6666666666666 / 3.6) < 1e-3 and tr.powerMin == -tr.powerMax      # efficiency.py:64-71 side effects
    assert tr.forceMin == -213900 and abs(tr.velocityMax - 160 / 3.6) < 1e-12
    eta = 0.73
    L, V = np.meshgrid(np.linspace(-100, 100, 200), np.linspace(1, 170, 170) / 3.6, indexing='ij')
    F = loadToForce(L, V, tr.forceMax, tr.powerMax)
    stat = F * V * (F > 0) * (1 - eta) / eta - (1 - eta) * F * V * (F < 0)
    lm = DynamicLossMap(tr.forceMax, 27000.0, 0.96)
    dyn_oracle = lm.total(torch.tensor(F), torch.tensor(V)).numpy()
    dyn_product = fun(F, V)
    for dyn in (dyn_oracle, dyn_product):
        assert 0.99 <= stat.max() / dyn.max() <= 1.01
    assert np.max(np.abs(dyn_oracle - dyn_product)) < 1e-6 * dyn_oracle.max()      # two independent spline constructions
    # the interpolant reproduces the measured table (min of the two converter configurations, 4 motors)
    from mseetc.data import dataLosses
    A, B = dataLosses()
    best = np.minimum(np.array(A['losses']), np.array(B['losses'])) * 4
    lut = motorLossesFunction(Train(config={'id': 'NL_Intercity_VIRM6'})).lut
    loads = np.array(A['loads'], float); loads[-1] += 1e-4
    speeds = (((np.array(A['frequencies']) - 20) / 150) * 140 + 20) / 3.6
    LL, SS = np.meshgrid(loads, speeds, indexing='ij')
    assert np.max(np.abs(lut(LL, SS) - best)) < 1e-8
    assert lut(150.0, 20.0) == 0.0 and fun(3e5, 20.0) == 0.0                      # zero outside the measured box


def test_integrated_loss_energies_match_an_adaptive_integrator():
    """oracle/intlosses.py (TrainIntegrator.calcLosses, reference train.py:367-413, integrated there by CVODES with relTol 1e-6):
    the fixed-step RK4 restatement against scipy's DOP853 at 1e-12 on the same right-hand side -- constant efficiencies (smooth) and
    the spline loss map incl. intervals whose speed crosses its kinks (32 steps there)."""
    import torch
    from scipy.integrate import solve_ivp
    from oracle.intlosses import energy_fn, static_power_fns, dynamic_power_fns
    from oracle.lossmap import DynamicLossMap
    tr = fig5_train()
    M = tr.mass * tr.rho
    sr = (tr.r0 / M, tr.r1 / M, tr.r2 / M)
    lm = DynamicLossMap(tr.forceMax, 27000.0, 0.96, 1.0)
    kinks = (lm.box[2], lm.powerMax / lm.forceMax, lm.box[3])
    # (b0, Fel, dt, c0): acceleration from 1 m/s through both kinks, cruising, braking, a gradient
    pts = np.array([[1.0, 0.25, 60.0, 0.0], [400.0, 0.05, 20.0, 0.002], [900.0, -0.2, 15.0, -0.01], [36.0, 0.28, 10.0, 0.0]])
    b0, Fel, dt, c0 = pts.T
    zero = np.zeros(len(b0))

    def reference(power_fns):
        out = []
        for i in range(len(b0)):
            def rhs(t, x):
                v = torch.tensor([x[0]], dtype=torch.float64)
                ptr, prg = power_fns(torch.tensor([Fel[i]], dtype=torch.float64), v)
                return [Fel[i] - (sr[0] + sr[1] * x[0] + sr[2] * x[0] ** 2) - c0[i], float(ptr.detach()[0]), float(prg.detach()[0])]
            sol = solve_ivp(rhs, [0.0, dt[i]], [np.sqrt(b0[i]), 0.0, 0.0], method='DOP853', rtol=1e-12, atol=1e-12, max_step=dt[i] / 60)
            out.append(sol.y[:, -1])
        return np.array(out)
    for power_fns, kk, tol in ((static_power_fns(0.12, 0.15), (), 1e-8), (dynamic_power_fns(lm, M), kinks, 1e-5)):      # 6e-6 on the start from 1 m/s through both kinks
        ref = reference(power_fns)
        fn = energy_fn(sr, None, power_fns, False, steps=8, kinks=kk)
        v_end = ref[:, 0]
        (etr, _, _), (erg, _, _) = fn(b0, Fel, zero, dt, c0, v_end ** 2, derivs=False)
        scale = np.maximum(np.abs(ref[:, 1]), np.abs(ref[:, 2]))
        assert np.max(np.abs(etr - ref[:, 1]) / scale) < tol, (np.abs(etr - ref[:, 1]) / scale)
        assert np.max(np.abs(erg - ref[:, 2]) / scale) < tol, (np.abs(erg - ref[:, 2]) / scale)
