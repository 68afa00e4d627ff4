"""Generates the oracle fixtures of the BASELINE-config parity tests (run here on the CPU, committed with its output):

  config5_mixed_dynamic.json   BASELINE configs[4] in miniature: 256 random tracks (rng 11), mixed numIntervals, spline loss map,
                               pn brake off, T = 1.15 Tmin_i -- oracle minimum times and optima of 10 evenly spaced instances,
                               plus the oracle outcome of every instance listed in CONFIG5_ALSO (the ones the device does not
                               converge on: the test asserts that the oracle does not converge on them either)
  config3_dynamic_mc.json      BASELINE configs[2], spline-loss-map half: first 512 instances of the recipe (rng 20260101) --
                               oracle optima of 8 evenly spaced instances
  config4_long_N2000.json      BASELINE configs[3] at N = 2000: synthetic 200 km track (rng 7), T = 8240 s -- oracle optimum
                               (full sparse KKT solves, no Riccati recursion, no parallel-in-time sweeps)

The reference (CasADi+IPOPT) cannot run in this image, so these are ORACLE outputs, not reference outputs."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'ms-eetc_b200')]
HERE = os.path.dirname(os.path.abspath(__file__))

CONFIG5_N = 256
CONFIG5_SPOT = list(np.linspace(0, CONFIG5_N - 1, 10).astype(int))
CONFIG5_ALSO = [int(x) for x in os.environ.get('CONFIG5_ALSO', '55,81,114').split(',') if x]      # the device's failures (profiles/probe_config5_failures.py)


def _oracle_imports():
    from common import FLAT_JSON, fig5_train, virm6, config5_instances, to_track_data, mc_dynamic_overrides      # noqa: F401
    from oracle.problem import load_track, discretization_points
    from oracle.nlp import ReferenceNLP
    from oracle.lossmap import DynamicLossMap
    from oracle import ipm
    return locals()


def dynamic_nlp(m, tr, track, N, aux, etag, scale, energy=True):
    pos, g, v, c = m['discretization_points'](track, N)
    opts = dict(numSteps=1, numApproxSteps=1, energyOptimal=energy, minimumVelocity=1)
    if not energy:
        tr = tr.copy(); tr.losses = ('none',)
        return m['ReferenceNLP'](tr, pos, g, v, c, track.length, opts)
    tr = tr.copy(); tr.losses = ('dynamic', aux, etag, scale)
    lm = m['DynamicLossMap'](tr.forceMax, aux, etag, scale)
    return m['ReferenceNLP'](tr, pos, g, v, c, track.length, opts, loss_rows=lm.rows(tr.mass * tr.rho))


def solve(m, nlp, T, max_iter=500):
    lbz, ubz, lbg, ubg = nlp.bounds(T)
    r = m['ipm'].solve(nlp, nlp.x0(T), lbz, ubz, lbg, ubg, max_iter=max_iter)
    u = nlp.unpack(r.x)
    return r, u


def config5_one(i):
    m = _oracle_imports()
    inst = m['config5_instances'](CONFIG5_N)
    N, track = inst[i]
    td = m['to_track_data'](track)
    tr = m['fig5_train']()
    t0 = time.time()
    tn = dynamic_nlp(m, tr, td, N, 27000.0, 0.96, 1.0, energy=False)
    horizon = 1.5 * float(np.sum(tn.ds / np.minimum(tn.limit[:-1], tr.velocityMax)))
    rt, ut = solve(m, tn, horizon)
    out = dict(index=int(i), N=int(N), tmin_status=rt.status, tmin=float(ut['t'][-1]) if rt.success else None)
    if rt.success:
        T = 1.15 * out['tmin']
        en = dynamic_nlp(m, tr, td, N, 27000.0, 0.96, 1.0)
        r, u = solve(m, en, T)
        out.update(T=T, status=r.status, success=bool(r.success), iterations=int(r.iters), kkt=float(r.kkt), cost_kwh=float(en.cost(r.f)),
                   t=u['t'].tolist(), b=u['b'].tolist(), Fel=u['Fel'].tolist())
    print('config5', i, N, out.get('status'), out.get('cost_kwh'), '%.0fs' % (time.time() - t0), flush=True)
    return out


def config3_one(i):
    m = _oracle_imports()
    ov = m['mc_dynamic_overrides'](512)
    tr = m['fig5_train']()
    tr.mass = float(ov['mass'][i]); tr.r0, tr.r1, tr.r2 = float(ov['r0'][i]), float(ov['r1'][i]), float(ov['r2'][i])
    track = m['load_track'](m['FLAT_JSON'])
    t0 = time.time()
    nlp = dynamic_nlp(m, tr, track, 300, float(ov['auxiliaries'][i]), 0.96, float(ov['tableScale'][i]))
    r, u = solve(m, nlp, 1541.0)
    print('config3', i, r.status, nlp.cost(r.f), '%.0fs' % (time.time() - t0), flush=True)
    return dict(index=int(i), status=r.status, success=bool(r.success), iterations=int(r.iters), kkt=float(r.kkt), cost_kwh=float(nlp.cost(r.f)),
                t=u['t'].tolist(), b=u['b'].tolist(), Fel=u['Fel'].tolist())


def config4():
    m = _oracle_imports()
    from mseetc.synthetic import random_track
    from common import oracle_nlp
    track = m['to_track_data'](random_track(np.random.default_rng(7), length=200e3))
    nlp = oracle_nlp(m['virm6'](), track, 2000)
    t0 = time.time()
    r, u = solve(m, nlp, 8240.0, max_iter=1000)
    print('config4', r.status, nlp.cost(r.f), r.iters, '%.0fs' % (time.time() - t0), flush=True)
    return dict(N=2000, T=8240.0, status=r.status, success=bool(r.success), iterations=int(r.iters), kkt=float(r.kkt), cost_kwh=float(nlp.cost(r.f)),
                t=u['t'].tolist(), b=u['b'].tolist(), Fel=u['Fel'].tolist())


if __name__ == '__main__':
    import multiprocessing as mp
    which = sys.argv[1:] or ['5', '3', '4']
    ctx = mp.get_context('spawn')
    with ctx.Pool(min(8, os.cpu_count() or 1)) as pool:
        jobs = {}
        if '4' in which:
            jobs['4'] = pool.apply_async(config4)
        if '5' in which:
            jobs['5'] = pool.map_async(config5_one, sorted(set(CONFIG5_SPOT + CONFIG5_ALSO)))
        if '3' in which:
            jobs['3'] = pool.map_async(config3_one, list(np.linspace(0, 511, 8).astype(int)))
        if '5' in jobs:
            json.dump(dict(n=CONFIG5_N, spot=[int(i) for i in CONFIG5_SPOT], also=CONFIG5_ALSO, instances=jobs['5'].get()),
                      open(os.path.join(HERE, 'config5_mixed_dynamic.json'), 'w'))
        if '3' in jobs:
            json.dump(dict(n=512, instances=jobs['3'].get()), open(os.path.join(HERE, 'config3_dynamic_mc.json'), 'w'))
        if '4' in jobs:
            json.dump(jobs['4'].get(), open(os.path.join(HERE, 'config4_long_N2000.json'), 'w'))
