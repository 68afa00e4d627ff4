"""Host layer (Train / Track / options / table) -- CPU only.  Includes the pure-Python assertions of the
reference's testClothoidApproximation (unitTests/curvatureResistance/curvatureResistance.py:204-286)."""
import copy
import ctypes
import os
import re

import numpy as np
import pytest

from common import FLAT_JSON, SWISS_JSON, TRAIN_JSON, ROOT


def test_train_fields_match_oracle_restatement():
    from mseetc.train import Train
    from oracle.problem import load_train
    a = Train(config={'id': 'NL_Intercity_VIRM6'})
    b = load_train(TRAIN_JSON)
    for k in ('mass', 'rho', 'velocityMax', 'forceMax', 'forceMin', 'forceMinPn', 'powerMax', 'powerMin', 'accMax', 'accMin',
              'r0', 'r1', 'r2', 'etaTraction', 'etaRgBrake', 'g'):
        assert getattr(a, k) == getattr(b, k), k
    M = a.mass * a.rho
    assert abs(M - 414460) < 1e-6 and abs(a.forceMax / M - 0.516093) < 1e-6      # SURVEY 8(a2) probed values
    assert abs(a.r0 / M - 1.41244e-2) < 1e-7 and abs(a.r2 / M - 3.12696e-5) < 1e-10


def test_train_overrides_and_errors():
    from mseetc.train import Train
    cfg = {'id': 'NL_Intercity_VIRM6', 'max deceleration': None, 'max acceleration': {'unit': 'm/s^2', 'value': 0.45}}
    t = Train(config=cfg)
    assert t.accMin is None and t.accMax == 0.45
    assert 'id' not in cfg                      # consumed, like the reference (train.py:42)
    with pytest.raises(ValueError):
        Train(config={'id': 'NL_Intercity_VIRM6', 'nonsense': {'unit': 'm', 'value': 1}})
    with pytest.raises(ValueError):
        Train(config={'id': 'NL_Intercity_VIRM6', 'mass': 5})
    with pytest.raises(ValueError):
        Train(config='NL_Intercity_VIRM6')
    t = Train(config={'id': 'NL_Intercity_VIRM6'})
    t.forceMin = 0
    t.forceMinPn = 0
    with pytest.raises(ValueError):
        t.checkFields()


def test_grid_matches_oracle_and_survey_facts():
    from mseetc.track import Track, computeDiscretizationPoints
    from oracle.problem import load_track, discretization_points
    for name, path in (('00_var_speed_limit_100', FLAT_JSON), ('CH_StGallen_Wil', SWISS_JSON)):
        tk = Track(config={'id': name})
        pts = computeDiscretizationPoints(tk, 300)
        pos, g, v, c = discretization_points(load_track(path), 300)
        assert len(pts) == 301
        assert np.array_equal(pts.index.values, pos)
        assert np.array_equal(pts['Gradient [permil]'].values, g)
        assert np.array_equal(pts['Speed limit [m/s]'].values, v)
        assert np.array_equal(pts['Curvature [1/m]'].values, c)
    ds = np.diff(pts.index.values)
    assert abs(ds.min() - 0.0713) < 1e-3 and abs(ds.max() - 217.32) < 1e-2
    assert len(tk.mergeDataFrames()) == 165
    with pytest.raises(ValueError):
        computeDiscretizationPoints(tk, 100)     # fewer intervals than track sections


def test_crop_and_reverse():
    from mseetc.track import Track
    tk = Track(config={'id': '00_var_speed_limit_100'})
    tk.updateLimits(positionEnd=8500)
    assert tk.length == 8500 and list(tk.speedLimits.index) == [0.0]
    tk = Track(config={'id': 'CH_StGallen_Wil'})
    L = tk.length
    g0 = tk.gradients.copy()
    tk.reverse().reverse()
    assert np.allclose(tk.gradients.index.values, g0.index.values) and np.allclose(tk.gradients.values, g0.values)
    assert abs(tk.length - L) < 1e-12
    with pytest.raises(ValueError):
        tk.updateLimits(positionStart=-1)


def test_clothoid_approximation_reference_assertions():
    from mseetc.track import Track
    track = Track(config={'id': '00_var_speed_limit_100'})
    t = copy.deepcopy(track)
    r0, rf = 1000, 500
    k0, kf = 1 / r0, 1 / rf
    t.importCurvatureTuples(tuples=[[0.0, r0, rf]])
    assert t.curvatures['Curvature [1/m]'].to_dict() == {0.0: (k0 + kf) / 2}
    t.importCurvatureTuples(tuples=[[0.0, r0, rf]], clothoidSamplingInterval=t.length + 1)
    assert t.curvatures['Curvature [1/m]'].to_dict() == {0.0: (k0 + kf) / 2}
    ds = t.length / 4
    t.importCurvatureTuples(tuples=[[0.0, r0, rf]], clothoidSamplingInterval=ds)
    alpha = t.length / (kf - k0)
    k1 = (k0 + (k0 + ds * 1 / alpha)) / 2
    k2 = ((k0 + ds * 1 / alpha) + (k0 + ds * 2 / alpha)) / 2
    k3 = ((k0 + ds * 2 / alpha) + (k0 + ds * 3 / alpha)) / 2
    k4 = ((k0 + ds * 3 / alpha) + kf) / 2
    assert t.curvatures['Curvature [1/m]'].to_dict() == {0.0: k1, ds: k2, 2 * ds: k3, 3 * ds: k4}
    ds = t.length / 4 + 1
    t.importCurvatureTuples(tuples=[[0.0, r0, rf]], clothoidSamplingInterval=ds)
    k1 = (k0 + (k0 + ds * 1 / alpha)) / 2
    k2 = ((k0 + ds * 1 / alpha) + (k0 + ds * 2 / alpha)) / 2
    k3 = ((k0 + ds * 2 / alpha) + kf) / 2
    assert t.curvatures['Curvature [1/m]'].to_dict() == {0.0: k1, ds: k2, 2 * ds: k3}
    t.importCurvatureTuples(tuples=[[0.0, r0, "infinity"]])
    assert t.curvatures['Curvature [1/m]'].to_dict() == {0.0: k0 / 2}
    with pytest.raises(ValueError):
        t.importCurvatureTuples(tuples=[[0.0, r0, rf]], clothoidSamplingInterval=-1)
    with pytest.raises(ValueError):
        t.importCurvatureTuples(tuples=[[0.0, 0.0, rf]])
    with pytest.raises(ValueError):
        t.importCurvatureTuples(tuples=[[500, r0, rf], [500, rf, 1 + rf]])
    with pytest.raises(ValueError):
        t.importCurvatureTuples(tuples=[[-1, r0, rf]])


def test_options_validation():
    from mseetc.ocp import OptionsCasadiSolver
    import json
    with open(os.path.join(ROOT, 'ms-eetc_b200', 'simulations', 'config.json')) as fh:
        cfg = json.load(fh)
    o = OptionsCasadiSolver(cfg)
    assert o.numIntervals == 300 and o.maxIterations == 500 and o.energyOptimal is True and o.minimumVelocity == 1
    assert o.integrationOptions.numSteps == 1 and o.integrationOptions.numApproxSteps == 1 and o.integrationOptions.order == 4
    d = o.toDict()
    assert d['integrationOptions'] == {'numApproxSteps': 1, 'numSteps': 1, 'order': 4}
    for bad in ({'nonexistent': 1}, {'numIntervals': 0}, {'numIntervals': 2.5}, {'energyOptimal': 1}, {'minimumVelocity': 0},
                {'integrationMethod': 'LINEAR'}, {'integrateLosses': 'yes'}, {'integrationOptions': {'order': 3}},
                {'integrationOptions': {'numSteps': 0}}, {'integrationOptions': {'bogus': 0}}):
        with pytest.raises(ValueError):
            OptionsCasadiSolver(bad)


def test_solver_construction_and_argument_errors_without_gpu():
    "Construction (grid, scalars) needs no device; bad boundary times raise before the ABI is crossed."
    from mseetc.ocp import casadiSolver
    from mseetc.train import Train
    from mseetc.track import Track
    s = casadiSolver(Train(config={'id': 'NL_Intercity_VIRM6'}), Track(config={'id': '00_var_speed_limit_100'}),
                     {'numIntervals': 300, 'integrationOptions': {'numApproxSteps': 1}})
    assert len(s.points) == 301 and len(s.steps) == 300 and s.withPnBrake and s.withRgBrake and s.numIntervals == 300
    assert abs(s.totalMass - 414460) < 1e-6 and abs(s.scalingFactorObjective - 3.6 / (1e-6 * 414460)) < 1e-9
    with pytest.raises(ValueError):
        s.solve(-5)
    with pytest.raises(ValueError):
        s.solve(100, initialTime=-1)
    with pytest.raises(ValueError):
        s.solve('abc')
    with pytest.raises(ValueError):
        casadiSolver(Train(config={'id': 'NL_Intercity_VIRM6'}), Track(config={'id': 'CH_StGallen_Wil'}), {'numIntervals': 100})


def test_parameter_planes_match_oracle():
    "The planes handed to the C ABI equal the oracle's independently derived specific-unit scalars."
    from mseetc.ocp import casadiSolver
    from mseetc.train import Train
    from mseetc.track import Track
    from mseetc import _cabi
    from common import virm6, oracle_nlp
    from oracle.problem import load_track
    import harness
    s = casadiSolver(Train(config={'id': 'NL_Intercity_VIRM6'}), Track(config={'id': 'CH_StGallen_Wil'}),
                     {'numIntervals': 300, 'integrationOptions': {'numApproxSteps': 1}})
    P, M = s._planes(1, np.array([1242.0]), np.array([0.0]), np.array([1.0]), np.array([1.0]), {}, 0.125 / 0.875, 0.3)
    nlp = oracle_nlp(virm6(), load_track(SWISS_JSON), 300)
    p, ds, c0, bmax = harness.pack_instance(nlp, 1242.0)
    assert np.allclose(P[:, 0], p, rtol=1e-14, atol=0)
    d2, c2, b2 = s._tables(s._base['rho'], s._base['g'], s._base['velocityMax'])
    assert np.array_equal(d2, ds) and np.allclose(c2, c0, rtol=1e-15) and np.array_equal(b2, bmax)


def test_cabi_library_exports_every_declared_symbol(built_lib):
    "No compute calls without a GPU: only that the library loads and exports what include/mseetc_b200.h declares."
    lib = ctypes.CDLL(built_lib)
    header = open(os.path.join(ROOT, 'include', 'mseetc_b200.h')).read()
    names = set(re.findall(r'\b(mseetc_[a-z_]+)\s*\(', header))
    assert {'mseetc_create', 'mseetc_destroy', 'mseetc_solve_batch', 'mseetc_workspace_bytes', 'mseetc_eval_interval',
            'mseetc_last_error', 'mseetc_version'} <= names
    for n in names:
        assert hasattr(lib, n), n
    lib.mseetc_version.restype = ctypes.c_int
    assert lib.mseetc_version() == 100


def test_product_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from mseetc.ocp import casadiSolver
    from mseetc.train import Train
    from mseetc.track import Track
    s = casadiSolver(Train(config={'id': 'NL_Intercity_VIRM6'}), Track(config={'id': '00_var_speed_limit_100'}),
                     {'numIntervals': 50, 'integrationOptions': {'numApproxSteps': 1}})
    with pytest.raises(RuntimeError):
        s.solve(1541)


def test_postprocess_table_columns():
    "Column set / order of the returned table (reference ocp.py:401-405, utils.py:230-334)."
    from mseetc.ocp import casadiSolver
    from mseetc.train import Train
    from mseetc.track import Track
    from mseetc.utils import postProcessDataFrame
    from common import virm6, oracle_nlp, oracle_solve
    from oracle.problem import load_track
    N = 60
    train = Train(config={'id': 'NL_Intercity_VIRM6'})
    s = casadiSolver(train, Track(config={'id': '00_var_speed_limit_100'}), {'numIntervals': N, 'integrationOptions': {'numApproxSteps': 1}})
    nlp = oracle_nlp(virm6(), load_track(FLAT_JSON), N)
    r = oracle_solve(nlp, 1541.0)
    assert r.success
    df = postProcessDataFrame(s.table_from_z(r.x), s.points, train)
    assert df.index.name == 'Time [s]'
    assert list(df.columns) == ['Position [m]', 'Velocity [m/s]', 'Force (el) [N]', 'Force (pnb) [N]', 'Slacks', 'Speed limit [m/s]',
                                'Gradient [permil]', 'Curvature [1/m]', 'Force (acc) [N]', 'Force (rgb) [N]', 'Force [N]',
                                'Max. Power [kW]', 'Min. Power [kW]', 'Losses [kWh]', 'Energy [kWh]', 'Energy (pnb) [kWh]',
                                'Energy (kin) [kWh]', 'Acceleration [m/s^2]', 'Position - cvodes [m]', 'Velocity - cvodes [m/s]',
                                'Error position [m]', 'Error velocity [m/s]']
    assert np.isnan(df['Force (el) [N]'].values[-1]) and np.isnan(df['Slacks'].values[-1])
    # at the optimum the epigraph rows are active: sum(Energy) = J - smoothing penalty
    u = nlp.unpack(r.x)
    pen = 1e-3 * np.sum(np.diff(u['Fel']) ** 2) / nlp.scale
    assert abs(df['Energy [kWh]'].sum() - (r.f - pen)) / r.f < 1e-6
    # the re-simulated trajectory stays close to the RK4 multiple-shooting one
    assert df['Error velocity [m/s]'].max() < 0.5


def test_postprocess_integrated_losses_and_rolling_resistance():
    """postProcessDataFrame(integrateLosses=True, integrateRollingResistance=True) (reference utils.py:261-320): the columns are
    checked against an independent adaptive integration (scipy DOP853, rtol 1e-12) of the same ODEs (train.py:367-413, :415-456)."""
    from scipy.integrate import solve_ivp
    from mseetc.ocp import casadiSolver
    from mseetc.train import Train
    from mseetc.track import Track
    from mseetc.efficiency import totalLossesFunction
    from mseetc.utils import postProcessDataFrame
    from common import virm6, oracle_nlp, oracle_solve
    from oracle.problem import load_track
    N = 60
    nlp = oracle_nlp(virm6(), load_track(FLAT_JSON), N)
    r = oracle_solve(nlp, 1541.0)
    assert r.success
    for dynamic in (False, True):
        train = Train(config={'id': 'NL_Intercity_VIRM6'})
        s = casadiSolver(train, Track(config={'id': '00_var_speed_limit_100'}), {'numIntervals': N, 'integrationOptions': {'numApproxSteps': 1}})
        if dynamic:
            train.powerLosses = totalLossesFunction(train, auxiliaries=27000, etaGear=0.96)
        raw = s.table_from_z(r.x)
        mid = postProcessDataFrame(raw, s.points, train, CVODES=False)
        df = postProcessDataFrame(raw, s.points, train, CVODES=False, integrateLosses=True, integrateRollingResistance=True)
        assert list(df.columns) == list(mid.columns[:-1]) + ['Rolling resistance [kWh]', 'Acceleration [m/s^2]']
        assert np.isnan(df['Losses [kWh]'].values[-1]) and np.isnan(df['Rolling resistance [kWh]'].values[-1])
        M = train.mass * train.rho
        kWh = 1e-6 / 3.6
        loss = train.powerLossesFuns(split=False)
        t = df.index.values
        roll = lambda v: (train.r0 + train.r1 * v + train.r2 * v * v) / M
        for j in (0, 7, 23, 41, N - 1):
            f = df['Force (el) [N]'].values[j] / M
            p = df['Force (pnb) [N]'].values[j] / M
            gd = train.g * df['Gradient [permil]'].values[j] / 1e3 / train.rho
            v0 = df['Velocity [m/s]'].values[j]
            rhs = lambda _, y: [f + p - roll(y[0]) - gd, float(loss(f, y[0]))]
            ref = solve_ivp(rhs, (0.0, t[j + 1] - t[j]), [v0, 0.0], method='DOP853', rtol=1e-12, atol=1e-14).y[1, -1] * M * kWh
            assert abs(df['Losses [kWh]'].values[j] - ref) <= 1e-6 * abs(ref) + 1e-9      # the reference asks CVODES for reltol 1e-6
            fa = df['Force (acc) [N]'].values[j] / M
            rhs = lambda _, y: [2.0 * (fa + p - roll(np.sqrt(y[0])) - gd), roll(np.sqrt(y[0]))]
            ds = df['Position [m]'].values[j + 1] - df['Position [m]'].values[j]
            ref = solve_ivp(rhs, (0.0, ds), [v0 * v0, 0.0], method='DOP853', rtol=1e-12, atol=1e-14).y[1, -1] * M * kWh
            assert abs(df['Rolling resistance [kWh]'].values[j] - ref) <= 1e-6 * abs(ref) + 1e-9
        if not dynamic:
            # constant efficiencies: losses are proportional to force x distance, so the two estimates differ only by the
            # distance the time-domain re-simulation covers versus the grid step
            rel = np.nanmax(np.abs(df['Losses [kWh]'] - mid['Losses [kWh]'])) / np.nanmax(mid['Losses [kWh]'])
            assert rel < 2e-2
        assert abs(np.nansum(df['Energy [kWh]']) - np.nansum(mid['Energy [kWh]'])) / np.nansum(mid['Energy [kWh]']) < 2e-2


def test_stream_pool_interleave_is_a_tile_aligned_permutation():
    "Sub-batches of the two-stream solve: every instance once, tiles of 32 dealt round-robin, every sub-batch starts on a tile."
    from mseetc._cabi import StreamPool
    for n, k in ((4096, 2), (4100, 2), (1030, 3), (64, 2), (33, 2), (5000, 4)):
        perm, parts = StreamPool.interleave(n, k)
        assert sorted(perm.tolist()) == list(range(n))
        assert parts[0][0] == 0 and parts[-1][1] == n and all(a % 32 == 0 for a, _ in parts)
        assert all(b == a2 for (_, b), (a2, _) in zip(parts[:-1], parts[1:]))
        first = perm[parts[0][0]:parts[0][1]]
        assert first[:32].tolist() == list(range(32))                      # tile 0 -> sub-batch 0
        if len(parts) > 1 and n >= 64:
            second = perm[parts[1][0]:parts[1][1]]
            assert second[0] == 32                                          # tile 1 -> sub-batch 1
        sizes = [b - a for a, b in parts]
        assert max(sizes) - min(sizes) <= 64                                # balanced up to the ragged tile


def test_track_batch_grids_match_compute_discretization_points(built_lib):
    """Native batch preprocessing (mseetc_discretize_tracks) against computeDiscretizationPoints track by track (reference
    track.py:91-107, :377-383): the two shipped tracks, random tracks with mixed interval counts, and the ValueError case."""
    from mseetc.track import Track, computeDiscretizationPoints
    from mseetc.trackbatch import TrackBatch
    from mseetc.synthetic import random_track
    rng = np.random.default_rng(4)
    tracks = [Track(config={'id': '00_var_speed_limit_100'}), Track(config={'id': 'CH_StGallen_Wil'})] + [random_track(rng) for _ in range(20)]
    N = np.array([300, 300] + list(rng.choice([100, 200, 300, 400], 20)), dtype=np.int32)
    N[5] = 3                                               # more sections than intervals: the reference raises (track.py:103-105)
    grid = TrackBatch.from_tracks(tracks).discretize(N)
    for i, (tr, n) in enumerate(zip(tracks, N)):
        try:
            ref = computeDiscretizationPoints(tr, int(n))
        except ValueError:
            assert grid['error'][i] == 1
            continue
        assert grid['error'][i] == 0
        a = grid['off'][i] + i
        assert np.array_equal(grid['pos'][a:a + n + 1], ref.index.values)
        assert np.array_equal(grid['limit'][a:a + n + 1], ref['Speed limit [m/s]'].values)
        assert np.array_equal(grid['grad'][a:a + n + 1], ref['Gradient [permil]'].values)
        assert np.array_equal(grid['curv'][a:a + n + 1], ref['Curvature [1/m]'].values)
    assert grid['error'][5] == 1


def test_random_track_batch_has_the_recipe_statistics(built_lib):
    import time
    from mseetc.trackbatch import TrackBatch
    t0 = time.perf_counter()
    batch = TrackBatch.random(np.random.default_rng(11), 16384)
    grid = batch.discretize(np.random.default_rng(1).choice([100, 200, 300, 400], 16384))
    assert time.perf_counter() - t0 < 5.0                  # 16 384 tracks: well under the 2 s target on an idle box
    assert batch.n == 16384 and 5e3 <= batch.length.min() and batch.length.max() <= 50e3
    (lo, lp, lv), (go, gp, gv), (co, cp, cv) = batch.tables
    assert set(np.round(lv * 3.6).astype(int)) <= {80, 100, 120, 140} and np.abs(gv).max() <= 25.0 and np.abs(cv).max() <= 1 / 300 + 1e-12
    assert np.all(lp[lo[:-1]] == 0) and np.all(gp[go[:-1]] == 0)
    assert 0.6 < (cv == 0).mean() < 0.8 and (grid['error'] == 0).mean() > 0.9


def test_collocation_tableau_is_the_known_radau_and_gauss_scheme():
    "mseetc.train.collocationTableau (host side of integrationMethod 'IRK', reference train.py:310) against the textbook coefficients."
    from mseetc.train import collocationTableau, integratorSetup, OptionsIRK, OptionsCVODES, OptionsRK
    A, w, c = collocationTableau(2, 'radau')                      # Radau IIA, order 3
    assert np.allclose(c, [1 / 3, 1]) and np.allclose(w, [3 / 4, 1 / 4]) and np.allclose(A, [[5 / 12, -1 / 12], [3 / 4, 1 / 4]], atol=1e-14)
    A, w, c = collocationTableau(2, 'legendre')                   # Gauss, order 4
    r = np.sqrt(3) / 6
    assert np.allclose(c, [0.5 - r, 0.5 + r]) and np.allclose(w, [0.5, 0.5]) and np.allclose(A, [[0.25, 0.25 - r], [0.25 + r, 0.25]], atol=1e-14)
    A, w, c = collocationTableau(1, 'radau')                      # implicit Euler
    assert np.allclose(A, [[1.0]]) and np.allclose(w, [1.0]) and np.allclose(c, [1.0])
    for d in range(1, 10):
        for m in ('radau', 'legendre'):
            A, w, c = collocationTableau(d, m)
            assert abs(w.sum() - 1) < 1e-12 and np.allclose(A.sum(axis=1), c, atol=1e-11)          # consistency conditions
            order = 2 * d - 1 if m == 'radau' else 2 * d                                          # quadrature order of the points
            for q in range(1, order):
                assert abs(w @ c ** q - 1 / (q + 1)) < 1e-10, (d, m, q)
    with pytest.raises(ValueError, match='Unknown collocation method'):
        collocationTableau(2, 'lobatto')
    assert integratorSetup('RK', OptionsRK({'numSteps': 3, 'numApproxSteps': 2})) == (3, 2, None)
    ns, na, tab = integratorSetup('IRK', OptionsIRK({'order': 3, 'maxIter': 7}))
    assert (ns, na, tab['maxIter'], tab['A'].shape) == (1, 0, 7, (3, 3))
    ns, na, tab = integratorSetup('CVODES', OptionsCVODES({}))
    assert na == 0 and tab['A'].shape == (4, 4)                   # reference train.py:315: numApproxSteps = 0 with CVODES
