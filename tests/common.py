"""Shared problem builders for the tests (oracle side and product side from the same JSON fixtures)."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, 'ms-eetc_b200')
TRAIN_JSON = os.path.join(PKG, 'trains', 'NL_Intercity_VIRM6.json')
FLAT_JSON = os.path.join(PKG, 'tracks', '00_var_speed_limit_100.json')
SWISS_JSON = os.path.join(PKG, 'tracks', 'CH_StGallen_Wil.json')
RK = dict(numSteps=1, numApproxSteps=1)


def oracle_nlp(train, track, N, energy=True, vmin=1.0, **rk):
    from oracle.problem import discretization_points
    from oracle.nlp import ReferenceNLP
    pos, g, v, c = discretization_points(track, N)
    o = dict(RK); o.update(rk); o.update(energyOptimal=energy, minimumVelocity=vmin)
    interval_fn = None
    if o.get('irk'):                       # (order, collMethod): the reference's 'IRK' integrator (train.py:303-310)
        from oracle.irk import interval_rows
        interval_fn = interval_rows(o['irk'][0], o['irk'][1], o['numSteps'], o['numApproxSteps'])
    energy_fn = None
    if o.get('integrateLosses') and energy:    # ocp.py:231-241 with the constant-efficiency model of the train
        from oracle.intlosses import energy_fn as mk, static_power_fns
        M = train.mass * train.rho
        losses = train.losses if train.losses is not None else ('none',)
        sr = (train.r0 / M, train.r1 / M, train.r2 / M)
        if losses[0] == 'dynamic':             # efficiency.totalLossesFunction(train, auxiliaries, etaGear)
            from oracle.intlosses import dynamic_power_fns
            from oracle.lossmap import DynamicLossMap
            lm = DynamicLossMap(train.forceMax, losses[1], losses[2], losses[3])
            energy_fn = mk(sr, None, dynamic_power_fns(lm, M), train.forceMinPn != 0, steps=o.get('oracleLossSteps', 8),
                           kinks=(lm.box[2], lm.powerMax / lm.forceMax, lm.box[3]))
        else:
            cT, cR = ((1 - losses[1]) / losses[1], 1 - losses[2]) if losses[0] == 'static' else (0.0, 0.0)
            energy_fn = mk(sr, None, static_power_fns(cT, cR), train.forceMinPn != 0, steps=o.get('oracleLossSteps', 8))
    return ReferenceNLP(train, pos, g, v, c, track.length, o, interval_fn=interval_fn, energy_fn=energy_fn)


def oracle_solve(nlp, T, t0=0.0, v0=1.0, vN=1.0, **kw):
    from oracle import ipm
    lbz, ubz, lbg, ubg = nlp.bounds(T, t0, v0, vN)
    return ipm.solve(nlp, nlp.x0(T, t0), lbz, ubz, lbg, ubg, **kw)


def virm6(**mut):
    from oracle.problem import load_train
    t = load_train(TRAIN_JSON)
    for k, v in mut.items():
        setattr(t, k, v)
    return t


def fig5_train():
    "VIRM6 after the side effects of totalLossesFunction (reference efficiency.py:64-71), pn brake off."
    t = virm6(forceMinPn=0)
    t.powerMax = t.forceMax * (((55 - 20) / 150) * 140 + 20) / 3.6
    t.powerMin = -t.powerMax
    t.forceMin = -t.forceMax
    t.velocityMax = 160 / 3.6
    return t


def fig10_train():
    "reference simulations/figure10.py:14-22"
    t = virm6(forceMinPn=0)
    t.forceMin = -t.forceMax
    t.powerMax = 3129277
    t.powerMin = -t.powerMax
    t.losses = ('static', 0.73, 0.73)
    return t


def active_set(nlp, z, T, t0=0.0, v0=1.0, vN=1.0, rtol=1e-6):
    """Active variable bounds and rows at z.  Returns (active, ambiguous) boolean vectors over
    [lower z, upper z, lower g, upper g]; 'ambiguous' marks slacks within a decade of the threshold."""
    lbz, ubz, lbg, ubg = nlp.bounds(T, t0, v0, vN)
    g = nlp.g(z)
    slack = np.concatenate([z - lbz, ubz - z, g - lbg, ubg - g])
    bound = np.concatenate([lbz, ubz, lbg, ubg])
    scale = np.maximum(1.0, np.abs(np.where(np.isfinite(bound), bound, 1.0)))
    thr = rtol * scale
    with np.errstate(invalid='ignore'):
        active = slack <= thr
        ambiguous = (slack > thr / 10) & (slack < thr * 10)
    return active, ambiguous


def to_track_data(track):
    "Product Track -> oracle TrackData (same step functions)."
    from oracle.problem import TrackData
    return TrackData(track.length, (track.speedLimits.index.values, track.speedLimits.iloc[:, 0].values),
                     (track.gradients.index.values, track.gradients.iloc[:, 0].values),
                     (track.curvatures.index.values, track.curvatures.iloc[:, 0].values))


def config5_instances(n, seed=11):
    """BASELINE configs[4] recipe: n random tracks with numIntervals drawn from {100, 200, 300, 400}; a track
    with more sections than intervals (reference track.py:103-105) is drawn again.  Returns [(numIntervals, Track)]."""
    import sys
    if PKG not in sys.path:
        sys.path.insert(0, PKG)
    from mseetc.synthetic import random_track
    from mseetc.track import computeDiscretizationPoints
    rng = np.random.default_rng(seed)
    out = []
    while len(out) < n:
        N = int(rng.choice([100, 200, 300, 400]))
        track = random_track(rng)
        try:
            computeDiscretizationPoints(track, N)
        except ValueError:
            continue
        out.append((N, track))
    return out


def mc_dynamic_overrides(n, seed=20260101):
    "Spline-loss-map half of BASELINE configs[2] (SURVEY.md 8d recipe): the draws that follow the 32768 constant-efficiency instances."
    rng = np.random.default_rng(seed)
    half = 32768
    for _ in range(6):
        rng.uniform(0, 1, half)                      # mass, r0, r1, r2, etaTraction, etaRgBrake of the first half
    nominal = virm6()
    ov = dict(mass=391000 * rng.uniform(0.85, 1.15, half), r0=nominal.r0 * rng.uniform(0.8, 1.2, half), r1=nominal.r1 * rng.uniform(0.8, 1.2, half),
              r2=nominal.r2 * rng.uniform(0.8, 1.2, half), auxiliaries=rng.uniform(20e3, 35e3, half), tableScale=rng.uniform(0.9, 1.1, half))
    return {k: v[:n] for k, v in ov.items()}
