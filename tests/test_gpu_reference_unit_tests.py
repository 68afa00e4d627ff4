"""The reference's own unit tests (unitTests/curvatureResistance/curvatureResistance.py:94-201) run through the public API
of the drop-in (-m gpu): same trains, tracks, options, calls and assertions; only the imports and the JSON paths differ."""
import copy

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CONSTANT_K = 1 / 300            # [1/m]                     curvatureResistance.py:22
FINAL_POSITION = 3475           # [m]                       curvatureResistance.py:24
THRESHOLD_VELOCITY = 1e-3       #                           curvatureResistance.py:30
THRESHOLD_ENERGY = 5e-2         #                           curvatureResistance.py:32


@pytest.fixture(scope='module')
def tracks(built_lib):
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from mseetc.track import Track
    straight = Track(config={'id': '00_var_speed_limit_100'})
    curved = copy.deepcopy(straight)
    curved.importCurvatureTuples(tuples=[[0.0, str(1 / CONSTANT_K), str(1 / CONSTANT_K)]])       # curvatureResistance.py:44
    return straight, curved


def specific_curvature_resistance_force(g, rho):                                                # curvatureResistance.py:47-56
    k = abs(CONSTANT_K)
    return g * 0.5 * k / ((1 - 30 * k) * rho) * (k <= 1 / 300) + g * 0.65 * k / ((1 - 55 * k) * rho) * (k > 1 / 300)


def solve_ocp(track, energyOptimal, lossFunction, terminalTime, train, finalPosition):          # curvatureResistance.py:59-91
    from mseetc.ocp import casadiSolver
    v0 = vN = 1
    track.updateLimits(positionEnd=finalPosition)
    opts = {"maxIterations": 500, "numIntervals": 300, "integrationMethod": "RK",
            "integrationOptions": {"order": 4, "numSteps": 1, "numApproxSteps": 1},
            "energyOptimal": energyOptimal, "minimumVelocity": min(v0, vN)}
    train.powerLosses = lossFunction
    ocp0 = casadiSolver(train, track, opts)
    df0, _ = ocp0.solve(terminalTime, terminalVelocity=vN, initialVelocity=v0)
    return df0


def test_minimum_time_problem(tracks):                                                          # curvatureResistance.py:94-140
    from mseetc.train import Train
    straight, curved = tracks
    lossFun = lambda f, v: 0
    train = Train(config={'id': 'NL_Intercity_VIRM6'})
    train.forceMinPn = 0
    train.powerMax = None
    train.powerMin = None
    r0 = solve_ocp(straight, False, lossFun, 180, train, FINAL_POSITION)
    shift = specific_curvature_resistance_force(train.g, train.rho) * train.mass * train.rho
    train.forceMax = train.forceMax + shift
    train.forceMin = train.forceMin + shift
    r1 = solve_ocp(curved, False, lossFun, 180, train, FINAL_POSITION)
    assert r0 is not None and r1 is not None
    r0, r1 = r0.reset_index(), r1.reset_index()
    assert all(np.abs((r0['Velocity [m/s]'] - r1['Velocity [m/s]']) / r0['Velocity [m/s]']) <= THRESHOLD_VELOCITY)


def test_minimum_energy_problem(tracks):                                                        # curvatureResistance.py:143-201
    from mseetc.train import Train
    from mseetc.efficiency import totalLossesFunction
    straight, curved = tracks
    train = Train(config={'id': 'NL_Intercity_VIRM6'})
    train.forceMinPn = 0
    etaMax = 0.73
    noLosses = lambda f, v: 0
    idealLosses = lambda f, v: f * v * (f > 0) * (1 - etaMax) / etaMax - (1 - etaMax) * f * v * (f < 0)
    realLosses = totalLossesFunction(train, auxiliaries=27000, etaGear=0.96)
    work = specific_curvature_resistance_force(train.g, train.rho) * train.rho * train.mass * FINAL_POSITION / (3600 * 1000)
    results = []
    for lossFunction in (noLosses, idealLosses, realLosses):
        r0 = solve_ocp(straight, True, lossFunction, 200, train, FINAL_POSITION)
        r1 = solve_ocp(curved, True, lossFunction, 200, train, FINAL_POSITION)
        assert r0 is not None and r1 is not None
        e0, e1 = round(r0['Energy [kWh]'].sum(), 1), round(r1['Energy [kWh]'].sum(), 1)
        l0, l1 = round(r0['Losses [kWh]'].sum(), 1), round(r1['Losses [kWh]'].sum(), 1)
        results.append(abs(work - ((e1 - l1) - (e0 - l0))) / work)
    assert all(r <= THRESHOLD_ENERGY for r in results), results
