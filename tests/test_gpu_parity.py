"""Parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle (-m gpu)."""
import ctypes

import numpy as np
import pytest

from common import FLAT_JSON, SWISS_JSON, active_set, fig5_train, fig10_train, oracle_nlp, oracle_solve, virm6

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def cabi(built_lib):
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from mseetc import _cabi
    _cabi.lib()
    return _cabi


def device_solve(cabi, nlps, Ts, v0=1.0, vN=1.0, max_iter=500, want_lam=True, initial_guess=0, lanes=1):
    "Raw C-ABI call with arrays packed from oracle NLP objects (same packing as the CPU emulation harness)."
    import torch
    import harness
    ref = nlps[0]
    n = len(nlps)
    Nmax = max(x.N for x in nlps)
    packs = [harness.pack_instance(x, T, 0.0, v0, vN) for x, T in zip(nlps, Ts)]
    params = np.ascontiguousarray(np.stack([p[0] for p in packs], axis=1))
    nint = np.array([x.N for x in nlps], np.int32)
    trk_off = np.concatenate([[0], np.cumsum(nint)]).astype(np.int32)
    cu = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to('cuda', dtype=dt)
    h = cabi.Handle(Nmax, ref.withPn, ref.withPower, ref.energy, {'none': 0, 'static': 1}[ref.lossKind],
                    ref.opts['numSteps'], ref.opts['numApproxSteps'], max_iter, initial_guess=initial_guess)
    h.set_sweep_lanes(lanes)
    if ref.opts.get('irk'):                      # collocation integrator: the tableau the product computes
        h.set_integrator(*harness.product_tableau(*ref.opts['irk']), 10)
    if ref.opts.get('integrateLosses'):
        h.set_integrate_losses(True)
    out = h.solve_device(cu(params, torch.float64), cu(nint, torch.int32), cu(np.arange(n, dtype=np.int32), torch.int32),
                         cu(trk_off, torch.int32), cu(np.concatenate([p[1] for p in packs]), torch.float64),
                         cu(np.concatenate([p[2] for p in packs]), torch.float64),
                         cu(np.concatenate([p[3] for p in packs]), torch.float64), want_z=True, want_lam=want_lam)
    return {k: (v.cpu().numpy() if hasattr(v, 'cpu') else v) for k, v in out.items() if v is not None}


def test_interval_kernel_matches_sympy(cabi):
    from oracle.nlp import _stage_functions
    rng = np.random.default_rng(5)
    n = 4096
    b0 = rng.uniform(1.0, 1600.0, n); F = rng.uniform(-0.6, 0.5, n); ds = rng.uniform(0.05, 400.0, n)
    F = np.maximum(F, (2.0 - b0) / (2 * ds) + 0.3)
    c0 = rng.uniform(-0.2, 0.2, n)
    sr = (1.41244e-2, 1.78932e-4, 3.12696e-5)
    inp = np.stack([b0, F, ds, c0, np.full(n, sr[0]), np.full(n, sr[1]), np.full(n, sr[2])])
    for numSteps, numApprox in ((1, 1), (1, 0)):
        out = cabi.eval_interval(inp, numSteps, numApprox)
        names, fn = _stage_functions(False, numSteps, numApprox, 'none')
        res = fn(b0, F, np.zeros(n), np.ones(n), ds, c0, *sr, 0.0, 0.0)
        ct, cb = names.index('ct') * 15, names.index('cb') * 15
        for a, base in ((0, ct), (1, cb)):
            for i, j in enumerate((0, 1, 2, 5, 6, 9)):
                ref = -np.asarray(res[base + j]) * np.ones(n) + (1.0 if (a == 1 and i == 0) else 0.0)
                assert np.max(np.abs(out[6 * a + i] - ref) / np.maximum(1e-12, np.abs(ref))) < 1e-9


def test_collocation_interval_kernel_matches_oracle(cabi):
    "mseetc_eval_interval_irk (TrainIntegrator 'IRK', train.py:303-310, 347-364) vs the oracle's collocation restatement."
    import harness
    from oracle import irk
    rng = np.random.default_rng(5)
    n = 256
    b0 = rng.uniform(4.0, 1600.0, n); F = rng.uniform(-0.6, 0.5, n); ds = rng.uniform(0.05, 400.0, n)
    F = np.maximum(F, (6.0 - b0) / (2 * ds) + 0.3)
    c0 = rng.uniform(-0.2, 0.2, n)
    sr = (1.41244e-2, 1.78932e-4, 3.12696e-5)
    inp = np.stack([b0, F, ds, c0, np.full(n, sr[0]), np.full(n, sr[1]), np.full(n, sr[2])])
    for order, scheme, numSteps, numApprox in ((2, 'radau', 1, 0), (3, 'legendre', 1, 1), (4, 'legendre', 4, 0), (9, 'radau', 1, 2)):
        A, w = harness.product_tableau(order, scheme)
        out = cabi.eval_interval(inp, numSteps, numApprox, dict(A=A, w=w, maxIter=12))
        (tau, gt, ht), (phi, gp, hp) = irk.shoot(b0, F, ds, c0, sr, numSteps, numApprox, order, scheme)
        ref = [tau] + gt + ht + [phi] + gp + hp
        for i in range(12):
            assert np.max(np.abs(out[i] - ref[i]) / np.maximum(1e-10, np.abs(ref[i]))) < 2e-8, (order, scheme, numSteps, numApprox, i)


def test_public_api_integration_methods(cabi):
    """casadiSolver with integrationMethod 'IRK' and 'CVODES' (reference ocp.py:116, train.py:303-322): optimum against the oracle
    with the same collocation scheme; ERK4+ with the average-speed rule for time and the CVODES-equivalent scheme agree on the
    energy (a two-point Radau step on (t, b) does not: its quadrature of dt = ds/v over the first interval, where the train starts
    at 1 m/s, is coarse -- the reference's IRK branch has the same property, so that case is only compared with the oracle)."""
    import pandas as pd      # noqa: F401
    from mseetc.ocp import casadiSolver
    from mseetc.train import Train, TrainIntegrator, CVODES_EQUIVALENT as Q
    from mseetc.track import Track
    from oracle.problem import load_track
    train = Train(config={'id': 'NL_Intercity_VIRM6'})
    track = Track(config={'id': '00_var_speed_limit_100'})
    costs = {}
    for method, io, irk_key, ns, na in (('RK', {'numApproxSteps': 1}, None, 1, 1),
                                       ('IRK', {'order': 2, 'collMethod': 'radau', 'numApproxSteps': 0}, (2, 'radau'), 1, 0),
                                       ('CVODES', {}, (Q['order'], Q['collMethod']), Q['numSteps'], Q['numApproxSteps'])):
        solver = casadiSolver(train, track, {'numIntervals': 200, 'integrationMethod': method, 'integrationOptions': io})
        df, stats = solver.solve(1541.0)
        assert df is not None and stats['Cost'] > 0, (method, stats)
        costs[method] = stats['Cost']
        if irk_key:
            nlp = oracle_nlp(virm6(), load_track(FLAT_JSON), 200, energy=True, numSteps=ns, numApproxSteps=na, irk=irk_key)
            ref = oracle_solve(nlp, 1541.0)
            assert ref.success
            assert abs(stats['Cost'] - nlp.cost(ref.f)) <= 1e-6 * nlp.cost(ref.f), method
            assert np.max(np.abs(df['Velocity [m/s]'].values ** 2 - ref.x[nlp.iB])) <= 1e-4 * nlp.limit.max() ** 2
    assert abs(costs['RK'] - costs['CVODES']) < 5e-3 * costs['CVODES'], costs
    # TrainIntegrator.solve (train.py:347-364) with the three methods on one interval
    model = train.exportModel()
    ends = [TrainIntegrator(model, m, o).solve(0.0, 400.0, 250.0, traction=0.3) for m, o in (('RK', {'numSteps': 4}), ('IRK', {'order': 3}), ('CVODES', {}))]
    for e in ends[:2]:
        assert abs(e['time'] - ends[2]['time']) < 1e-5 * ends[2]['time'] and abs(e['velSquared'] - ends[2]['velSquared']) < 1e-5 * ends[2]['velSquared']


def test_spline_loss_map_with_the_other_integrators(cabi):
    "efficiency.py loss map + integrationMethod 'IRK' / 'CVODES' through the public API (kernels k_cell_*_dyn_irk)"
    from mseetc.ocp import casadiSolver
    from mseetc.train import Train
    from mseetc.track import Track
    from mseetc.efficiency import totalLossesFunction
    train = Train(config={'id': 'NL_Intercity_VIRM6'}); train.forceMinPn = 0
    train.powerLosses = totalLossesFunction(train, auxiliaries=27000, etaGear=0.96)
    track = Track(config={'id': 'CH_StGallen_Wil'})
    cost = {}
    for method, io in (('RK', {'numApproxSteps': 1}), ('CVODES', {}), ('IRK', {'order': 2, 'collMethod': 'radau'})):
        df, stats = casadiSolver(train, track, {'numIntervals': 300, 'integrationMethod': method, 'integrationOptions': io}).solve(1242.0)
        assert df is not None, (method, stats)
        cost[method] = stats['Cost']
    assert abs(cost['RK'] - cost['CVODES']) < 2e-4 * cost['CVODES']
    assert 2e-4 * cost['CVODES'] < abs(cost['IRK'] - cost['CVODES']) < 5e-3 * cost['CVODES']
    assert abs(cost['RK'] - 47.3401426) < 1e-5 and abs(cost['CVODES'] - 47.3379331) < 1e-5      # the CPU emulation's optima


@pytest.mark.parametrize('lanes', [1, 16], ids=['seq', 'pit16'])
def test_integrated_losses_match_oracle_reference_formulation(cabi, lanes):
    """integrateLosses = True (reference ocp.py:231-241, train.py:367-413) through the C ABI: the oracle keeps the reference's rows
    s_i - E(sqrt(b_i), t_{i+1} - t_i, Fel_i, Fpb_i); the kernels take the duration from the shooting function of the interval.  Same
    optimum: objective 1e-6, trajectories 1e-4, the oracle's rows satisfied by the device's solution."""
    from oracle.problem import load_track
    T = 1541.0
    nlp = oracle_nlp(virm6(), load_track(FLAT_JSON), 100, energy=True, integrateLosses=True)
    ref = oracle_solve(nlp, T)
    assert ref.success
    for guess in (0, 1):
        out = device_solve(cabi, [nlp], [T], initial_guess=guess, lanes=lanes)
        assert out['status'][0] == 0 and out['kkt'][0] <= 1e-8
        assert abs(out['obj'][0] - ref.f) <= 1e-6 * abs(ref.f)
        z = out['z'][0]
        for idx, scale in ((nlp.iB, nlp.limit.max() ** 2), (nlp.iT, T), (nlp.iFel, nlp.forceMax), (nlp.iS, np.max(ref.x[nlp.iS]))):
            assert np.max(np.abs(z[idx] - ref.x[idx])) <= 1e-4 * scale
        lbz, ubz, lbg, ubg = nlp.bounds(T)
        g = nlp.g(z)
        assert np.max(np.maximum(lbg - g, g - ubg)) < 1e-7 * max(1.0, np.max(ref.x[nlp.iS]))      # energies up to ~100 J/kg; 4 vs 8 RK4 steps
        # full KKT conditions of the REFERENCE formulation (rows on t_{i+1} - t_i; oracle with the device's 4 RK4 steps) at the device's
        # solution with the device's multipliers, the time-row multipliers mapped (io.cuh: cell_fix_time_multiplier_intl)
        same = oracle_nlp(virm6(), load_track(FLAT_JSON), 100, energy=True, integrateLosses=True, oracleLossSteps=4)
        lam = out['lam'][0]
        free = lbz != ubz
        r = (same.grad_f(z) + same.jac(z).T @ lam)[free]
        sl, su = (z - lbz)[free], (ubz - z)[free]
        with np.errstate(invalid='ignore'):
            comp = np.where(r > 0, r * sl, -r * np.where(np.isfinite(su), su, 1.0))
        assert np.max(comp) < 1e-7


def test_integrated_losses_trip_time_sweep_converges(cabi):
    """A sweep of 1024 trip times with integrateLosses = True on the bench problem: every instance reaches the full tolerance and
    the energy decreases with the trip time (a build whose loss-integration routine was not inlined passed the single-instance
    parity cases and stalled on 8 % of such a sweep)."""
    from mseetc.ocp import casadiSolver
    from mseetc.train import Train
    from mseetc.track import Track
    solver = casadiSolver(Train(config={'id': 'NL_Intercity_VIRM6'}), Track(config={'id': 'CH_StGallen_Wil'}),
                          {'numIntervals': 300, 'maxIterations': 500, 'integrateLosses': True, 'integrationMethod': 'RK',
                           'integrationOptions': {'order': 4, 'numSteps': 1, 'numApproxSteps': 1}})
    tmin = float(np.atleast_1d(solver.minimum_time()[0])[0])
    T = tmin * np.linspace(1.001, 1.2, 1024)
    res = solver.solve_batch(T, screen=False)
    assert np.all(np.asarray(res['status']) == 0), np.unique(np.asarray(res['status']), return_counts=True)
    assert np.all(np.asarray(res['kkt']) <= 1e-8) and int(np.max(res['iters'])) < 80
    assert np.all(np.diff(np.asarray(res['cost'])) < 0)


def test_public_api_integrate_losses(cabi):
    """casadiSolver(..., {'integrateLosses': True}) with constant efficiencies (against the oracle) and with the spline loss map of
    simulations/table3.py (against the oracle fixture); a batch of trip times gives the same optima as single solves."""
    from mseetc.ocp import casadiSolver
    from mseetc.train import Train
    from mseetc.track import Track
    from mseetc.efficiency import totalLossesFunction
    from oracle.problem import load_track
    track = Track(config={'id': '00_var_speed_limit_100'})
    opts = {'numIntervals': 100, 'integrateLosses': True, 'integrationMethod': 'RK', 'integrationOptions': {'numApproxSteps': 1}}
    solver = casadiSolver(Train(config={'id': 'NL_Intercity_VIRM6'}), track, opts)
    df, stats = solver.solve(1541.0)
    nlp = oracle_nlp(virm6(), load_track(FLAT_JSON), 100, energy=True, integrateLosses=True)
    ref = oracle_solve(nlp, 1541.0)
    assert df is not None and abs(stats['Cost'] - nlp.cost(ref.f)) <= 1e-6 * nlp.cost(ref.f)
    assert np.max(np.abs(df['Velocity [m/s]'].values ** 2 - ref.x[nlp.iB])) <= 1e-4 * nlp.limit.max() ** 2
    plain = casadiSolver(Train(config={'id': 'NL_Intercity_VIRM6'}), track, dict(opts, integrateLosses=False)).solve(1541.0)[1]['Cost']
    assert 1e-5 < abs(plain - stats['Cost']) / plain < 1e-3            # another formulation, not the mid-point rows
    res = solver.solve_batch(np.array([1541.0, 1600.0, 1700.0]), screen=False)
    assert np.all(res['status'] == 0) and abs(res['cost'][0] - stats['Cost']) <= 1e-9 * stats['Cost'] and np.all(np.diff(res['cost']) < 0)
    # spline loss map
    gold = _golden('intlosses_dynamic_flat_N60.json')
    train = Train(config={'id': 'NL_Intercity_VIRM6'}); train.forceMinPn = 0
    train.powerLosses = totalLossesFunction(train, auxiliaries=gold['auxiliaries'], etaGear=gold['etaGear'])
    dyn = casadiSolver(train, track, dict(opts, numIntervals=gold['N']))
    df, stats = dyn.solve(gold['T'])
    assert df is not None and abs(stats['Cost'] - gold['cost_kwh']) <= 1e-6 * gold['cost_kwh']
    assert np.max(np.abs(df['Velocity [m/s]'].values ** 2 - np.array(gold['b']))) <= 1e-4 * 1975.0
    assert np.max(np.abs(df.index.values - np.array(gold['t']))) <= 1e-4 * gold['T']
    # together with the collocation integrator (three Gauss points, time from the average-speed rule): against the oracle
    irk = casadiSolver(Train(config={'id': 'NL_Intercity_VIRM6'}), track, dict(opts, integrationMethod='IRK',
                       integrationOptions={'order': 3, 'collMethod': 'legendre', 'numApproxSteps': 1}))
    df, stats = irk.solve(1541.0)
    nlp3 = oracle_nlp(virm6(), load_track(FLAT_JSON), 100, energy=True, integrateLosses=True, numSteps=1, numApproxSteps=1, irk=(3, 'legendre'))
    ref3 = oracle_solve(nlp3, 1541.0)
    assert df is not None and ref3.success and abs(stats['Cost'] - nlp3.cost(ref3.f)) <= 1e-6 * nlp3.cost(ref3.f)
    assert np.max(np.abs(df['Velocity [m/s]'].values ** 2 - ref3.x[nlp3.iB])) <= 1e-4 * nlp3.limit.max() ** 2


def test_interval_kernel_bitwise_equals_host_compilation(cabi):
    "Same source compiled by nvcc (device) and g++ (tests/hostsim): values agree to the last few ulps (FMA contraction differs)."
    import harness
    lib = harness.build()
    rng = np.random.default_rng(9)
    n = 1000
    inp = np.ascontiguousarray(np.stack([rng.uniform(1, 1500, n), rng.uniform(0.0, 0.5, n), rng.uniform(1, 300, n), rng.uniform(-0.1, 0.1, n),
                                         np.full(n, 1.4e-2), np.full(n, 1.8e-4), np.full(n, 3.1e-5)]))
    dev = cabi.eval_interval(inp, 1, 1)
    host = np.zeros((12, n))
    lib.hostsim_eval_interval(n, 1, 1, inp.ctypes.data, host.ctypes.data)
    assert np.max(np.abs(dev - host) / np.maximum(1e-300, np.abs(host))) < 1e-11


CASES = [
    ('config1 flat energy', lambda: virm6(), FLAT_JSON, None, 300, 1541.0, True, 1.0, 1.0, dict()),
    ('swiss energy', lambda: virm6(), SWISS_JSON, None, 300, 1242.0, True, 1.0, 1.0, dict()),
    ('swiss time-optimal', lambda: virm6(), SWISS_JSON, None, 300, 2000.0, False, 1.0, 1.0, dict()),
    ('figure10 no pn brake', fig10_train, FLAT_JSON, None, 300, 1541.0, True, 1.0, 1.0, dict()),
    ('figure5 time-optimal crop', lambda: (lambda t: (setattr(t, 'losses', ('none',)), t)[1])(fig5_train()), FLAT_JSON, 8500, 300,
     354.0, False, 1.0, 100 / 3.6, dict()),
    ('unit test energy no power rows', lambda: virm6(forceMinPn=0, powerMax=None, powerMin=None, losses=('none',)), FLAT_JSON, 3475,
     300, 200.0, True, 1.0, 1.0, dict()),
    ('rk4 on both states', lambda: virm6(), FLAT_JSON, None, 300, 1541.0, True, 1.0, 1.0, dict(numApproxSteps=0)),
    ('two time sub-points', lambda: virm6(), FLAT_JSON, None, 200, 1541.0, True, 1.0, 1.0, dict(numSteps=1, numApproxSteps=2)),
    # integrationMethod 'IRK' (reference train.py:303-310)
    ('irk radau 2 on both states', lambda: virm6(), FLAT_JSON, None, 200, 1541.0, True, 1.0, 1.0, dict(numApproxSteps=0, irk=(2, 'radau'))),
    ('irk legendre 3, two steps, time approximation', lambda: virm6(), SWISS_JSON, None, 300, 1242.0, True, 1.0, 1.0,
     dict(numSteps=2, numApproxSteps=1, irk=(3, 'legendre'))),
]


_ORACLE = {}


def oracle_case(case):
    "Oracle NLP and optimum of one parity case (solved once per session)."
    from oracle.problem import load_track
    name, mk, path, crop, N, T, energy, v0, vN, rk = case
    if name not in _ORACLE:
        track = load_track(path)
        if crop:
            track.crop(positionEnd=crop)
        nlp = oracle_nlp(mk(), track, N, energy=energy, **rk)
        _ORACLE[name] = (nlp, oracle_solve(nlp, T, v0=v0, vN=vN))
    return _ORACLE[name]


# sweeps: 1 = sequential Riccati sweeps, 8 / 16 / 32 = parallel-in-time sweeps with that many chunk lanes per instance
@pytest.mark.parametrize('lanes', [1, 8, 16, 32], ids=['seq', 'pit8', 'pit16', 'pit32'])
@pytest.mark.parametrize('guess', [0, 1], ids=['reference-guess', 'profile-guess'])
@pytest.mark.parametrize('case', CASES, ids=[c[0] for c in CASES])
def test_cuda_solver_matches_oracle(cabi, case, guess, lanes):
    """north_star bar: optimal energy 1e-6 relative, trajectories 1e-4 relative, same active set, KKT <= 1e-8 -- with the
    sequential and with the parallel-in-time sweeps, each against the oracle."""
    name, mk, path, crop, N, T, energy, v0, vN, rk = case
    nlp, ref = oracle_case(case)
    assert ref.success, ref.status
    out = device_solve(cabi, [nlp], [T], v0=v0, vN=vN, initial_guess=guess, lanes=lanes)
    assert out['status'][0] == 0
    assert out['kkt'][0] <= 1e-8
    if guess == 0:
        assert abs(int(out['iters'][0]) - ref.iters) <= max(3, ref.iters // 4)     # same algorithm, same start: IPOPT-like counts
    assert abs(out['obj'][0] - ref.f) <= 1e-6 * abs(ref.f)
    z, zr = out['z'][0], ref.x
    # with zero losses (unit-test case) the force profile has a flat direction: the optimum is unique only to ~1e-3 there
    ftol = 1e-3 if nlp.lossKind == 'none' and energy else 1e-4
    for idx, scale, tol in ((nlp.iB, nlp.limit.max() ** 2, 1e-4), (nlp.iT, T, 1e-4), (nlp.iFel, nlp.forceMax, ftol)):
        assert np.max(np.abs(z[idx] - zr[idx])) <= tol * scale
    if name.startswith('figure5'):
        assert abs(z[nlp.iT[-1]] - 272.4726) < 1e-4          # reference simulations/figure5.py:96
    lam = out['lam'][0]
    lbz, ubz, lbg, ubg = nlp.bounds(T, 0.0, v0, vN)
    free = lbz != ubz
    r = (nlp.grad_f(z) + nlp.jac(z).T @ lam)[free]
    sl, su = (z - lbz)[free], (ubz - z)[free]
    with np.errstate(invalid='ignore'):
        comp = np.where(r > 0, r * sl, -r * np.where(np.isfinite(su), su, 1.0))
    # (collocation cases: the oracle's Jacobian comes from autograd through unrolled Newton iterations and agrees with the device's
    # to ~1e-8 relative only, which shows here through the multipliers)
    assert np.max(comp) < (1e-6 if rk.get('irk') else 1e-7)
    g = nlp.g(z)
    assert np.max(np.maximum(lbg - g, g - ubg)) < 1e-7
    a0, amb0 = active_set(nlp, z, T, 0.0, v0, vN)
    a1, amb1 = active_set(nlp, zr, T, 0.0, v0, vN)
    sure = ~(amb0 | amb1)
    assert np.array_equal(a0[sure], a1[sure])


def test_public_api_single_solve_table(cabi):
    "casadiSolver(train, track, opts).solve(T) -> (DataFrame, stats) as the reference returns (ocp.py:310-409)."
    import json
    import os
    from mseetc.ocp import casadiSolver
    from mseetc.train import Train
    from mseetc.track import Track
    from common import PKG
    with open(os.path.join(PKG, 'simulations', 'config.json')) as fh:
        opts = json.load(fh)
    train = Train(config={'id': 'NL_Intercity_VIRM6'})
    solver = casadiSolver(train, Track(config={'id': '00_var_speed_limit_100'}), opts)
    df, stats = solver.solve(1541)
    assert df is not None and stats['Solver status'] == 'Solve_Succeeded'
    assert set(stats) == {'Solver status', 'IP iterations', 'CPU time [s]', 'Cost'}
    nlp = oracle_nlp(virm6(), __import__('oracle.problem', fromlist=['load_track']).load_track(FLAT_JSON), 300)
    ref = oracle_solve(nlp, 1541.0)
    assert abs(stats['Cost'] - nlp.cost(ref.f)) <= 1e-6 * ref.f
    assert len(df) == 301 and df.index.name == 'Time [s]'
    assert np.max(np.abs(df['Velocity [m/s]'].values - np.sqrt(ref.x[nlp.iB]))) < 1e-4 * 38.9
    assert abs(df.index.values[-1] - 1541) < 1e-3
    # infeasible trip time: (None, stats), no exception (reference ocp.py:364-370)
    df2, stats2 = solver.solve(1000)
    assert df2 is None and stats2['Solver status'] != 'Solve_Succeeded'
    # same solver object, repeated solve: identical iteration count (reference table3.py:60-62)
    df3, stats3 = solver.solve(1541)
    assert stats3['IP iterations'] == stats['IP iterations'] and np.array_equal(df3.values[:, :5], df.values[:, :5], equal_nan=True)


def test_trip_time_sweep_batch_properties(cabi):
    """BASELINE config 2 at full size: 4096 VIRM6 instances on CH_StGallen_Wil, T in Tmin*[0.8, 1.2].
    Size-independent properties + oracle spot checks."""
    from mseetc.ocp import casadiSolver
    from mseetc.train import Train
    from mseetc.track import Track
    from oracle.problem import load_track
    opts = {'numIntervals': 300, 'maxIterations': 500, 'integrationMethod': 'RK',
            'integrationOptions': {'order': 4, 'numSteps': 1, 'numApproxSteps': 1}}
    train = Train(config={'id': 'NL_Intercity_VIRM6'})
    solver = casadiSolver(train, Track(config={'id': 'CH_StGallen_Wil'}), opts)
    tsolver = casadiSolver(train, Track(config={'id': 'CH_StGallen_Wil'}), dict(opts, energyOptimal=False))
    tmin = tsolver.solve_batch(1500.0)
    assert tmin['status'][0] == 0
    Tmin = tmin['z'][0][-2]
    assert abs(Tmin - 1035.5536) < 2e-3
    n = 4096
    T = Tmin * (0.8 + 0.4 * np.arange(n) / (n - 1))
    res = solver.solve_batch(T)
    feas = T >= Tmin * (1 + 1e-6)
    assert np.all(res['status'][feas] == 0)                       # every feasible instance converges
    assert np.all(res['status'][T < Tmin * (1 - 1e-6)] == 4)      # infeasible instances are flagged, never "solved"
    assert np.all(res['kkt'][feas] <= 1e-8)
    cost = res['cost'][feas]
    assert np.all(np.diff(cost) < 0)                              # energy strictly decreases with trip time
    tN = res['z'][feas, -2]
    assert np.all(tN <= T[feas] * (1 + 2e-8))
    b = res['z'][feas][:, 4::5][:, :300]
    assert b.min() >= 1 - 1e-6 and np.sqrt(b.max()) <= 125 / 3.6 * (1 + 1e-6)
    nlp = oracle_nlp(virm6(), load_track(SWISS_JSON), 300)
    idx = np.where(feas)[0]
    for i in (idx[3], idx[len(idx) // 2], idx[-1]):
        ref = oracle_solve(nlp, float(T[i]))
        assert ref.success
        assert abs(res['obj'][i] - ref.f) <= 1e-6 * ref.f
        assert np.max(np.abs(res['z'][i][nlp.iB] - ref.x[nlp.iB])) <= 1e-4 * 1206.0
    # bitwise determinism for a fixed batch composition
    res2 = solver.solve_batch(T)
    assert np.array_equal(res['z'], res2['z']) and np.array_equal(res['iters'][feas], res2['iters'][feas])
    assert np.array_equal(res['status'], res2['status'])


def test_parameter_monte_carlo_batch(cabi):
    "BASELINE config 3 (static-efficiency half), reduced to 512 instances + oracle spot checks."
    from mseetc.ocp import casadiSolver
    from mseetc.train import Train
    from mseetc.track import Track
    from oracle.problem import load_track
    opts = {'numIntervals': 300, 'maxIterations': 500, 'integrationMethod': 'RK',
            'integrationOptions': {'order': 4, 'numSteps': 1, 'numApproxSteps': 1}}
    train = Train(config={'id': 'NL_Intercity_VIRM6'})
    solver = casadiSolver(train, Track(config={'id': '00_var_speed_limit_100'}), opts)
    rng = np.random.default_rng(20260101)
    n = 512
    ov = dict(mass=391000 * rng.uniform(0.85, 1.15, n), r0=train.r0 * rng.uniform(0.8, 1.2, n), r1=train.r1 * rng.uniform(0.8, 1.2, n),
              r2=train.r2 * rng.uniform(0.8, 1.2, n), etaTraction=rng.uniform(0.80, 0.92, n), etaRgBrake=rng.uniform(0.55, 0.85, n))
    res = solver.solve_batch(1541.0, overrides=ov)
    assert np.all(res['status'] == 0) and np.all(res['kkt'] <= 1e-8)
    for i in (0, 17, 511):
        t = virm6(mass=ov['mass'][i], r0=ov['r0'][i], r1=ov['r1'][i], r2=ov['r2'][i])
        t.losses = ('static', ov['etaTraction'][i], ov['etaRgBrake'][i])
        nlp = oracle_nlp(t, load_track(FLAT_JSON), 300)
        ref = oracle_solve(nlp, 1541.0)
        assert ref.success
        assert abs(res['cost'][i] - nlp.cost(ref.f)) <= 1e-6 * nlp.cost(ref.f)


def test_cabi_usage_errors(cabi):
    lib = cabi.lib()
    h = ctypes.c_void_p(0)
    bad = cabi.Problem(1, 1, 1, 1, 1, 1, 1, 100, 1e-8, 0.1, 0, 0)
    assert lib.mseetc_create(ctypes.byref(bad), ctypes.byref(h)) < 0
    assert b'n_intervals_max' in lib.mseetc_last_error()
    good = cabi.Problem(50, 1, 1, 1, 1, 1, 1, 100, 1e-8, 0.1, 1, 0)
    assert lib.mseetc_create(ctypes.byref(good), ctypes.byref(h)) == 0
    assert lib.mseetc_workspace_bytes(h, 64) > 0
    null = ctypes.c_void_p(0)
    rc = lib.mseetc_solve_batch(h, 4, *([null] * 14), null, 0, null)
    assert rc < 0 and b'null' in lib.mseetc_last_error()
    assert lib.mseetc_destroy(h) == 0


def test_dynamic_loss_rows_kernel_matches_oracle(cabi):
    from test_hostsim_parity import _dynamic_params, loss_row_points, check_loss_rows
    import harness
    tr = fig5_train()
    M = tr.mass * tr.rho
    Fel, b0, b1 = loss_row_points(4096, seed=3)
    tl, tv, cf = harness.loss_map_arrays()
    h = cabi.Handle(300, 0, 1, 1, 2, 1, 1, 500)
    h.set_loss_map(tl, tv, cf)
    out = h.eval_loss_rows(Fel, b0, b1, _dynamic_params(len(Fel), M, tr.forceMax))
    check_loss_rows(out, Fel, b0, b1, M, tr.forceMax)


def test_dynamic_loss_map_public_api_matches_golden(cabi):
    """reference simulations/table3.py:14-31 through the drop-in API: Train, totalLossesFunction (with its side effects on
    the train), casadiSolver(...).solve(1541) -- against the oracle fixture tests/golden/table3_dynamic_flat_N300.json."""
    import json, os
    from mseetc.ocp import casadiSolver
    from mseetc.train import Train
    from mseetc.track import Track
    from mseetc.efficiency import totalLossesFunction
    here = os.path.dirname(os.path.abspath(__file__))
    with open(os.path.join(here, '..', 'ms-eetc_b200', 'simulations', 'config.json')) as fh:
        opts = json.load(fh)
    opts['minimumVelocity'] = 1
    for name, track_id in (('table3_dynamic_flat_N300', '00_var_speed_limit_100'), ('dynamic_swiss_N300', 'CH_StGallen_Wil')):
        gold = json.load(open(os.path.join(here, 'golden', name + '.json')))
        train = Train(config={'id': 'NL_Intercity_VIRM6'})
        train.forceMinPn = 0
        train.powerLosses = totalLossesFunction(train, auxiliaries=27000, etaGear=0.96)
        solver = casadiSolver(train, Track(config={'id': track_id}), opts)
        solver.initialGuess = 'reference'          # same starting point as the oracle -> same iteration count
        df, stats = solver.solve(gold['T'], terminalVelocity=1, initialVelocity=1)
        assert df is not None
        assert abs(stats['Cost'] - gold['cost_kwh']) <= 1e-6 * abs(gold['cost_kwh'])
        assert stats['IP iterations'] == gold['iterations']
        fast = casadiSolver(train, Track(config={'id': track_id}), opts)      # default: speed-envelope starting profile
        df2, stats2 = fast.solve(gold['T'], terminalVelocity=1, initialVelocity=1)
        assert abs(stats2['Cost'] - gold['cost_kwh']) <= 1e-6 * abs(gold['cost_kwh'])
        assert stats2['IP iterations'] < gold['iterations']
        assert np.max(np.abs(df2['Velocity [m/s]'].values - np.sqrt(np.array(gold['b'])))) <= 1e-4 * 44.5
        assert np.max(np.abs(df['Velocity [m/s]'].values - np.sqrt(np.array(gold['b'])))) <= 1e-4 * 44.5
        # the table's energy column: sum = J - smoothing penalty (the epigraph rows are active at the optimum)
        Fel = np.array(gold['Fel'])
        pen = 1e-3 * np.sum(np.diff(Fel) ** 2) * (1e-6 * solver.totalMass / 3.6)
        assert abs(df['Energy [kWh]'].sum() - (gold['cost_kwh'] - pen)) <= 2e-6 * gold['cost_kwh']
    # parameter study on the loss map (BASELINE config 3, dynamic half): per-instance auxiliaries and table scale
    rng = np.random.default_rng(5)
    n = 64
    res = solver.solve_batch(1242.0, overrides=dict(auxiliaries=rng.uniform(20e3, 35e3, n), tableScale=rng.uniform(0.9, 1.1, n)))
    assert np.all(res['status'] == 0) and np.all(res['kkt'] <= 1e-8)
    assert res['cost'].std() > 0.05


def test_stream_pool_is_bitwise_identical_to_single_stream(cabi):
    "Concurrent sub-batches (StreamPool) change the schedule, not the arithmetic."
    from mseetc.ocp import casadiSolver
    from mseetc.train import Train
    from mseetc.track import Track
    opts = {'numIntervals': 300, 'maxIterations': 500, 'integrationMethod': 'RK', 'integrationOptions': {'order': 4, 'numSteps': 1, 'numApproxSteps': 1}}
    train = Train(config={'id': 'NL_Intercity_VIRM6'})
    T = np.linspace(1040.0, 1400.0, 1500)
    one = casadiSolver(train, Track(config={'id': 'CH_StGallen_Wil'}), opts)
    one.streams = 1
    three = casadiSolver(train, Track(config={'id': 'CH_StGallen_Wil'}), opts)
    three.streams = 3
    a = one.solve_batch(T, screen=False)
    b = three.solve_batch(T, screen=False)
    assert np.all(a['status'] == 0)
    for key in ('z', 'obj', 'kkt', 'iters', 'status'):
        assert np.array_equal(a[key], b[key]), key


def test_pooled_batch_with_overrides_and_screening_matches_single_stream(cabi):
    """Per-instance parameters, per-instance track tables (rho) and screening through the tile-interleaved two-stream path give
    the same results, in the caller's order, as one stream; trips below their own minimum time are reported infeasible."""
    from mseetc.ocp import casadiSolver
    from mseetc.train import Train
    from mseetc.track import Track
    opts = {'numIntervals': 200, 'maxIterations': 500, 'integrationMethod': 'RK', 'integrationOptions': {'order': 4, 'numSteps': 1, 'numApproxSteps': 1}}
    train = Train(config={'id': 'NL_Intercity_VIRM6'})
    rng = np.random.default_rng(5)
    n = 1100
    ov = dict(mass=391000 * rng.uniform(0.85, 1.15, n), r0=train.r0 * rng.uniform(0.8, 1.2, n), rho=rng.uniform(1.04, 1.08, n),
              etaTraction=rng.uniform(0.80, 0.92, n))
    T = rng.uniform(1300.0, 1700.0, n)            # the minimum time on this track is about 1470 s: a third is infeasible
    res = {}
    for streams in (1, 2):
        solver = casadiSolver(train, Track(config={'id': '00_var_speed_limit_100'}), opts)
        solver.streams = streams
        res[streams] = solver.solve_batch(T, overrides=ov)
    a, b = res[1], res[2]
    assert np.array_equal(a['status'], b['status']) and np.array_equal(a['tmin'], b['tmin'])
    ok = a['status'] == 0
    assert ok.sum() > 500 and (a['status'] == 4).sum() > 200
    for key in ('z', 'obj', 'kkt', 'iters'):
        assert np.array_equal(a[key][ok], b[key][ok]), key
    assert np.all(a['kkt'][ok] <= 1e-8)
    feasible = T >= a['tmin'] * (1 - 1e-9)
    assert np.all(ok[feasible]) and np.all(a['status'][~feasible] == 4)


def test_early_screening_flags_are_checked_against_the_certificate(cabi, monkeypatch):
    """The envelope screening of the device (before the first iteration) is only trusted when the exact minimum time confirms
    it: with a (faked) certificate that calls those trips feasible, the flagged instances are solved again without
    screening -- they then end as the plain iteration ends on an infeasible problem, not as 'Infeasible_Problem_Detected'."""
    from mseetc import ocp
    from mseetc.train import Train
    from mseetc.track import Track
    opts = {'numIntervals': 300, 'maxIterations': 120, 'integrationMethod': 'RK', 'integrationOptions': {'order': 4, 'numSteps': 1, 'numApproxSteps': 1}}
    solver = ocp.casadiSolver(Train(config={'id': 'NL_Intercity_VIRM6'}), Track(config={'id': 'CH_StGallen_Wil'}), opts)
    T = np.array([900.0, 950.0, 1100.0, 1242.0])
    plain = solver.solve_batch(T)
    assert list(plain['status']) == [4, 4, 0, 0] and list(plain['iters'][:2]) == [0, 0]
    real_join = ocp._Presolve.join
    monkeypatch.setattr(ocp._Presolve, 'join', lambda self: 0.5 * real_join(self))
    faked = solver.solve_batch(T)
    assert list(faked['status'][2:]) == [0, 0] and np.array_equal(faked['z'][2:], plain['z'][2:])
    assert all(st not in (0, 4) for st in faked['status'][:2]) and np.all(faked['iters'][:2] > 0)


def test_concurrent_handles_of_different_batch_sizes(cabi):
    """Regression: the sweep kernel's shared-memory request depends on the batch size of a handle, and handles of different
    sizes run concurrently on several host threads (two sub-batches of 4096 and a presolve batch of 8192 distinct problems
    here) -- the kernel attribute that admits the request is process-wide and must not be lowered by another handle."""
    from mseetc.ocp import casadiSolver
    from mseetc.train import Train
    from mseetc.track import Track
    opts = {'numIntervals': 300, 'maxIterations': 500, 'integrationMethod': 'RK', 'integrationOptions': {'order': 4, 'numSteps': 1, 'numApproxSteps': 1}}
    train = Train(config={'id': 'NL_Intercity_VIRM6'})
    solver = casadiSolver(train, Track(config={'id': '00_var_speed_limit_100'}), opts)
    solver.streams = 2
    rng = np.random.default_rng(3)
    n = 8192
    ov = dict(mass=391000 * rng.uniform(0.85, 1.15, n), r0=train.r0 * rng.uniform(0.8, 1.2, n), etaTraction=rng.uniform(0.80, 0.92, n))
    res = solver.solve_batch(1541.0, overrides=ov)
    assert np.all(res['status'] == 0) and np.all(res['kkt'] <= 1e-8)


def test_solve_instances_mixed_tracks_and_interval_counts(cabi):
    "BASELINE config 5 in miniature: random tracks, mixed numIntervals, one device call; oracle spot check."
    from mseetc.ocp import casadiSolver, solve_instances
    from mseetc.train import Train
    from mseetc.synthetic import random_track
    from oracle.problem import TrackData
    rng = np.random.default_rng(11)
    train = Train(config={'id': 'NL_Intercity_VIRM6'})
    solvers = []
    while len(solvers) < 48:
        N = int(rng.choice([100, 200, 300]))
        track = random_track(rng)
        try:
            solvers.append(casadiSolver(train, track, {'numIntervals': N, 'maxIterations': 300, 'integrationOptions': {'numApproxSteps': 1}}))
            solvers[-1]._track = track
        except ValueError:
            continue
    lim = [np.minimum(s.points['Speed limit [m/s]'].values[:-1], s._base['velocityMax']) for s in solvers]
    T = 1.3 * np.array([float(np.sum(s.steps / l)) for s, l in zip(solvers, lim)]) + 60.0
    res = solve_instances(solvers, T)
    ok = res['status'] == 0
    assert ok.sum() >= 40 and np.all(res['kkt'][ok] <= 1e-8)
    assert np.all((res['status'] == 0) | (res['status'] == 4) | (res['status'] == 1) | (res['status'] == 2))
    i = int(np.where(ok)[0][0])
    s = solvers[i]
    tk = s._track
    td = TrackData(tk.length, (tk.speedLimits.index.values, tk.speedLimits.iloc[:, 0].values), (tk.gradients.index.values, tk.gradients.iloc[:, 0].values),
                   (tk.curvatures.index.values, tk.curvatures.iloc[:, 0].values))
    nlp = oracle_nlp(virm6(), td, s.numIntervals)
    ref = oracle_solve(nlp, float(T[i]))
    assert ref.success
    assert abs(res['obj'][i] - ref.f) <= 1e-6 * abs(ref.f)
    z = res['z'][i][:nlp.nz]
    assert np.max(np.abs(z[nlp.iB] - ref.x[nlp.iB])) <= 1e-4 * nlp.limit.max() ** 2


def test_parallel_in_time_sweeps_long_horizon(cabi):
    """BASELINE config 4 in miniature: one instance, synthetic 200 km track, N = 2000.  The parallel-in-time sweeps
    (32 lanes per instance, default for N >= 2048, forced here) must reproduce the sequential sweeps."""
    from mseetc.ocp import casadiSolver
    from mseetc.train import Train
    from mseetc.synthetic import random_track
    train = Train(config={'id': 'NL_Intercity_VIRM6'})
    track = random_track(np.random.default_rng(7), length=200e3)
    o = {'numIntervals': 2000, 'maxIterations': 1000, 'integrationMethod': 'RK', 'integrationOptions': {'order': 4, 'numSteps': 1, 'numApproxSteps': 1}}
    out = {}
    for lanes in (1, 32):
        s = casadiSolver(train, track, o)
        s.sweepLanes = lanes
        out[lanes] = s.solve_batch(8240.0, screen=False)
        assert out[lanes]['status'][0] == 0 and out[lanes]['kkt'][0] <= 1e-8
    assert abs(out[32]['obj'][0] - out[1]['obj'][0]) <= 1e-9 * abs(out[1]['obj'][0])
    assert np.max(np.abs(out[32]['z'] - out[1]['z'])) < 1e-5
    auto = casadiSolver(train, track, dict(o, numIntervals=2048))
    assert auto._make_handle is not None and auto.sweepLanes == 'auto'


# ---- BASELINE configs 3, 4, 5 against committed oracle fixtures (tests/golden/make_golden_configs.py) ----------------------
def _golden(name):
    import json
    import os
    return json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', name)))


def _dynamic_train():
    "reference simulations/table3.py:14-22: pn brake off, totalLossesFunction(train, 27000, 0.96)"
    from mseetc.train import Train
    from mseetc.efficiency import totalLossesFunction
    train = Train(config={'id': 'NL_Intercity_VIRM6'})
    train.forceMinPn = 0
    train.powerLosses = totalLossesFunction(train, auxiliaries=27000, etaGear=0.96)
    return train


def _check_instance(z, gold, N, stp, fmax_specific, T):
    "trajectories of one solution vector (reference variable order, ocp.py:166-181,248-249) against a fixture entry: 1e-4 of scale"
    body = z[:N * stp].reshape(N, stp)
    t = np.append(body[:, stp - 2], z[N * stp])
    b = np.append(body[:, stp - 1], z[N * stp + 1])
    gb = np.array(gold['b'])
    assert np.max(np.abs(t - np.array(gold['t']))) <= 1e-4 * T
    assert np.max(np.abs(b - gb)) <= 1e-4 * gb.max()
    assert np.max(np.abs(body[:, 0] - np.array(gold['Fel']))) <= 1e-4 * fmax_specific


def test_config5_mixed_tracks_dynamic_map_matches_oracle_fixture(cabi):
    """BASELINE configs[4] in miniature, as specified: 256 random tracks (rng 11), mixed numIntervals, the spline loss map of
    efficiency.py active, pn brake off, T = 1.15 Tmin_i.  Minimum times and optima of 10 evenly spaced instances against the
    oracle; every instance the device does not converge on must be one the oracle does not converge on either."""
    from common import config5_instances
    from mseetc.ocp import casadiSolver, solve_instances
    gold = _golden('config5_mixed_dynamic.json')
    inst = config5_instances(gold['n'])
    train = _dynamic_train()
    mk = lambda N, track, energy: casadiSolver(train, track, {'numIntervals': N, 'maxIterations': 500, 'integrationMethod': 'RK', 'energyOptimal': energy,
                                                              'integrationOptions': {'order': 4, 'numSteps': 1, 'numApproxSteps': 1}})
    solvers = [mk(N, tr, True) for N, tr in inst]
    tsolvers = [mk(N, tr, False) for N, tr in inst]
    lim = [np.minimum(s.points['Speed limit [m/s]'].values[:-1], s._base['velocityMax']) for s in solvers]
    horizon = 1.5 * np.array([float(np.sum(s.steps / l)) for s, l in zip(solvers, lim)])
    tres = solve_instances(tsolvers, horizon, screen=False)
    nint = np.array([s.numIntervals for s in solvers])
    stp = 4
    tmin = tres['z'][np.arange(len(inst)), nint * stp]
    by_index = {g['index']: g for g in gold['instances']}
    T = 1.15 * tmin
    for i in gold['spot']:
        g = by_index[i]
        assert g['N'] == inst[i][0] and tres['status'][i] == 0
        assert abs(tmin[i] - g['tmin']) <= 1e-6 * g['tmin']
        T[i] = g['T']                                    # exactly the oracle's terminal time for the spot checks
    feasible = tres['status'] == 0
    res = solve_instances([s for s, f in zip(solvers, feasible) if f], T[feasible], screen=False)
    where = np.flatnonzero(feasible)
    status = dict(zip(where.tolist(), res['status'].tolist()))
    assert feasible.sum() >= 250 and np.all(res['kkt'][res['status'] == 0] <= 1e-8)
    for i in gold['spot']:
        g, j = by_index[i], int(np.searchsorted(where, i))
        assert g['success'] and status[i] == 0
        assert abs(res['cost'][j] - g['cost_kwh']) <= 1e-6 * abs(g['cost_kwh'])
        _check_instance(res['z'][j], g, g['N'], stp, train.forceMax / (train.mass * train.rho), g['T'])
    failed = [i for i, st in status.items() if st != 0]
    assert len(failed) <= 0.05 * len(status)
    for i in failed:                                     # a device failure must be an oracle failure (fixture: CONFIG5_ALSO)
        assert i in by_index and not by_index[i].get('success', False), (i, status[i])


def test_config3_dynamic_monte_carlo_matches_oracle_fixture(cabi):
    """Spline-loss-map half of BASELINE configs[2]: the first 512 instances of the recipe (per-instance mass, Davis coefficients,
    auxiliaries, loss-table scale) in one call; all converge; 8 evenly spaced instances against the oracle."""
    from common import mc_dynamic_overrides
    from mseetc.ocp import casadiSolver
    from mseetc.track import Track
    gold = _golden('config3_dynamic_mc.json')
    train = _dynamic_train()
    solver = casadiSolver(train, Track(config={'id': '00_var_speed_limit_100'}),
                          {'numIntervals': 300, 'maxIterations': 500, 'integrationMethod': 'RK', 'integrationOptions': {'order': 4, 'numSteps': 1, 'numApproxSteps': 1}})
    ov = mc_dynamic_overrides(gold['n'])
    res = solver.solve_batch(1541.0, overrides=ov, screen=np.zeros(gold['n'], dtype=bool))
    assert np.all(res['status'] == 0) and np.all(res['kkt'] <= 1e-8)
    for g in gold['instances']:
        i = g['index']
        assert g['success']
        assert abs(res['cost'][i] - g['cost_kwh']) <= 1e-6 * abs(g['cost_kwh'])
        _check_instance(res['z'][i], g, 300, 4, train.forceMax / (ov['mass'][i] * train.rho), 1541.0)


def test_config4_long_horizon_parallel_in_time_matches_oracle_fixture(cabi):
    """BASELINE configs[3] at N = 2000 (synthetic 200 km track, rng 7): the parallel-in-time sweeps (32 chunk lanes, what the
    library picks from 1024 intervals on) against the ORACLE's optimum, which comes from sparse LU solves of the full KKT matrix."""
    from mseetc.ocp import casadiSolver
    from mseetc.train import Train
    from mseetc.synthetic import random_track
    gold = _golden('config4_long_N2000.json')
    train = Train(config={'id': 'NL_Intercity_VIRM6'})
    track = random_track(np.random.default_rng(7), length=200e3)
    solver = casadiSolver(train, track, {'numIntervals': 2000, 'maxIterations': 1000, 'integrationMethod': 'RK',
                                         'integrationOptions': {'order': 4, 'numSteps': 1, 'numApproxSteps': 1}})
    res = solver.solve_batch(gold['T'], screen=False)
    assert solver._ensure_handle().sweep_lanes() == 32
    assert res['status'][0] == 0 and res['kkt'][0] <= 1e-8 and gold['success']
    assert abs(res['cost'][0] - gold['cost_kwh']) <= 1e-6 * abs(gold['cost_kwh'])
    _check_instance(res['z'][0], gold, 2000, 5, train.forceMax / (train.mass * train.rho), gold['T'])


@pytest.mark.parametrize('kind', ['static', 'dynamic'])
def test_batched_tables_match_post_process_data_frame(cabi, kind):
    """solve_batch(..., tables=True): the device-side tables of a batch (csrc/table.cuh) against the host restatement of
    utils.postProcessDataFrame applied to each instance (reference ocp.py:407, utils.py:223-336), incl. the re-simulation columns."""
    from mseetc.ocp import casadiSolver
    from mseetc.train import Train
    from mseetc.track import Track
    from mseetc.utils import postProcessDataFrame
    opts = {'numIntervals': 300, 'maxIterations': 500, 'integrationMethod': 'RK', 'integrationOptions': {'order': 4, 'numSteps': 1, 'numApproxSteps': 1}}
    train = _dynamic_train() if kind == 'dynamic' else Train(config={'id': 'NL_Intercity_VIRM6'})
    solver = casadiSolver(train, Track(config={'id': 'CH_StGallen_Wil'}), opts)
    T = np.array([1000.0, 1100.0, 1242.0, 1300.0])                       # the first one is infeasible
    rng = np.random.default_rng(5)
    ov = dict(mass=train.mass * rng.uniform(0.9, 1.1, 4))
    res = solver.solve_batch(T, overrides=ov, tables=True)
    assert list(res['status']) == [4, 0, 0, 0]
    assert np.all(np.isnan(res['table'][0]))
    for i in (1, 2, 3):
        import copy
        tr = copy.copy(train)
        tr.mass = float(ov['mass'][i])
        one = copy.copy(solver)
        one.totalMass = tr.mass * tr.rho
        ref = postProcessDataFrame(one.table_from_z(res['z'][i]), solver.points, tr)
        got = solver.table(res, i)
        assert list(got.columns) == list(ref.columns) and got.index.name == ref.index.name
        assert np.allclose(got.index.values, ref.index.values, rtol=0, atol=0)
        for col in ref.columns:
            a, b = got[col].values, ref[col].values
            assert np.array_equal(np.isnan(a), np.isnan(b)), col
            scale = max(1.0, np.nanmax(np.abs(b)))
            tol = 1e-7 if 'cvodes' in col or 'Error' in col else 1e-11
            assert np.nanmax(np.abs(a - b)) <= tol * scale, (col, np.nanmax(np.abs(a - b)))


def test_solve_tracks_matches_solve_instances(cabi):
    """Batch preprocessing path (mseetc.trackbatch: native grids + vectorised tables, no casadiSolver per track) against the
    per-track path (one casadiSolver per track, solve_instances): the same tables, hence the same results bit for bit."""
    from mseetc.ocp import casadiSolver, solve_instances
    from mseetc.trackbatch import TrackBatch, solve_tracks
    from mseetc.synthetic import random_track
    rng = np.random.default_rng(21)
    train = _dynamic_train()
    tracks, N = [], []
    while len(tracks) < 48:
        tr, n = random_track(rng), int(rng.choice([100, 200, 300, 400]))
        try:
            casadiSolver(train, tr, {'numIntervals': n})
        except ValueError:
            continue
        tracks.append(tr); N.append(n)
    opts = {'maxIterations': 500, 'integrationMethod': 'RK', 'integrationOptions': {'order': 4, 'numSteps': 1, 'numApproxSteps': 1}}
    batch = TrackBatch.from_tracks(tracks)
    a = solve_tracks(train, batch, np.array(N), opts, timeFactor=1.15)
    assert np.all(a['grid_error'] == 0) and (a['status'] == 0).sum() >= 44
    solvers = [casadiSolver(train, tr, dict(opts, numIntervals=n)) for tr, n in zip(tracks, N)]
    b = solve_instances(solvers, a['terminalTime'], screen=False)
    assert np.array_equal(a['status'], b['status']) and np.array_equal(a['iters'], b['iters'])
    ok = a['status'] == 0
    assert np.array_equal(a['z'][ok], b['z'][ok]) and np.array_equal(a['cost'][ok], b['cost'][ok])


def test_energy_optimum_converges_to_the_gpops_energy_of_the_reference(cabi):
    """The only energy-optimal numbers the reference repository holds are the GPOPS results of the figure-10 problem
    (gpops/00_var_speed_limit_100_GPOPSI.csv / ...GPOPSII.csv: 440.1414723 / 440.1406149 kWh, simulations/figure10.py:14-36),
    solutions of the CONTINUOUS problem.  The multiple-shooting optimum carries an O(1/N) discretisation error (0.2 % at N = 300);
    solved on the device at N = 1200, 2400, 4800 and extrapolated to N -> infinity (Aitken) it must meet them to 1e-4."""
    import csv
    import os
    from mseetc.ocp import casadiSolver
    from mseetc.train import Train
    from mseetc.track import Track
    here = os.path.dirname(os.path.abspath(__file__))
    gpops = [float(list(csv.reader(open(os.path.join(here, 'golden', '00_var_speed_limit_100_GPOPS%s.csv' % v))))[1][6]) for v in ('I', 'II')]
    assert abs(gpops[0] - 440.1414723) < 1e-6 and abs(gpops[1] - 440.1406149) < 1e-6
    train = Train(config={'id': 'NL_Intercity_VIRM6'})                    # figure10.py:14-22
    train.forceMinPn = 0
    train.forceMin = -train.forceMax
    train.powerMax = 3129277
    train.powerMin = -train.powerMax
    train.etaTraction = train.etaRgBrake = 0.73
    J = []
    for N in (1200, 2400, 4800):
        solver = casadiSolver(train, Track(config={'id': '00_var_speed_limit_100'}),
                              {'numIntervals': N, 'maxIterations': 1000, 'integrationMethod': 'RK', 'integrationOptions': {'order': 4, 'numSteps': 1, 'numApproxSteps': 1}})
        res = solver.solve_batch(1541.0, screen=False)
        assert res['status'][0] == 0 and res['kkt'][0] <= 1e-8
        J.append(float(res['cost'][0]))
    assert J[0] > J[1] > J[2] > max(gpops)                                 # monotone from above
    d1, d2 = J[1] - J[0], J[2] - J[1]
    limit = J[2] - d2 * d2 / (d2 - d1)
    assert abs(limit - gpops[1]) <= 1e-4 * gpops[1] and abs(limit - gpops[0]) <= 1e-4 * gpops[0], (J, limit)


def test_compaction_of_the_running_batch_does_not_change_results(cabi):
    """csrc/compact.cuh: running instances are moved into the slots of finished ones so that they fill whole warps.  An instance
    computes the same numbers in any slot: every output is bitwise identical with and without compaction."""
    from mseetc.ocp import casadiSolver
    from mseetc.train import Train
    from mseetc.track import Track
    opts = {'numIntervals': 300, 'maxIterations': 500, 'integrationMethod': 'RK', 'integrationOptions': {'order': 4, 'numSteps': 1, 'numApproxSteps': 1}}
    train = Train(config={'id': 'NL_Intercity_VIRM6'})
    rng = np.random.default_rng(8)
    T = rng.permutation(np.linspace(1037.0, 1400.0, 3000))               # iteration counts scattered over the slots
    out = {}
    for on in (True, False):
        solver = casadiSolver(train, Track(config={'id': 'CH_StGallen_Wil'}), opts)
        h = solver._ensure_handle()
        h.set_compaction(on)
        res = solver.solve_batch(T, screen=False, want_multipliers=True)
        out[on] = {k: np.array(v) for k, v in res.items() if isinstance(v, np.ndarray)}
        passes, moved = h.last_compactions()
        assert (passes > 0) == on and (moved > 100) == on               # a batch in random order does get compacted
    assert np.all(out[True]['status'] == 0)
    for key in ('z', 'lam', 'obj', 'kkt', 'iters', 'status'):
        assert np.array_equal(out[True][key], out[False][key]), key
    # a larger parameter Monte Carlo (sequential sweeps beyond 4096 instances): some instances are in the middle of their line
    # search when they are moved -- their step planes have to travel with them
    n = 16384
    ov = dict(mass=391000 * rng.uniform(0.85, 1.15, n), r0=train.r0 * rng.uniform(0.8, 1.2, n), r1=train.r1 * rng.uniform(0.8, 1.2, n),
              r2=train.r2 * rng.uniform(0.8, 1.2, n), etaTraction=rng.uniform(0.80, 0.92, n), etaRgBrake=rng.uniform(0.55, 0.85, n))
    out = {}
    for on in (True, False):
        solver = casadiSolver(train, Track(config={'id': '00_var_speed_limit_100'}), opts)
        solver._ensure_handle().set_compaction(on)
        res = solver.solve_batch(1541.0, overrides=ov, screen=np.zeros(n, dtype=bool), restart=False)
        out[on] = {k: np.array(v) for k, v in res.items() if isinstance(v, np.ndarray)}
        assert (solver._ensure_handle().last_compactions()[1] > 500) == on
    assert np.all(out[True]['status'] == 0)
    for key in ('z', 'obj', 'kkt', 'iters', 'status'):
        assert np.array_equal(out[True][key], out[False][key]), key
