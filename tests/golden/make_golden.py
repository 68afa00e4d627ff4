"""Generates tests/golden/*.json from the CPU oracle (run here, committed with its output).

The reference (CasADi+IPOPT) cannot run in this image, so these are ORACLE outputs, not reference outputs; they
make the slow oracle solves (torch-autograd loss map) available to the GPU tests as fixtures."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests')]

from common import FLAT_JSON, SWISS_JSON, fig5_train, virm6, oracle_nlp, oracle_solve   # noqa: E402
from oracle.problem import load_track, discretization_points                              # noqa: E402
from oracle.nlp import ReferenceNLP                                                        # noqa: E402
from oracle.lossmap import DynamicLossMap                                                  # noqa: E402
from oracle import ipm                                                                     # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def dynamic_case(name, track_json, N, T, aux=27000.0, eta_gear=0.96):
    "reference simulations/table3.py:14-31 (pn brake off, totalLossesFunction(train, 27000, 0.96))"
    tr = fig5_train()
    tr.losses = ('dynamic', aux, eta_gear, 1.0)
    track = load_track(track_json)
    pos, g, v, c = discretization_points(track, N)
    lm = DynamicLossMap(tr.forceMax, aux, eta_gear)
    nlp = ReferenceNLP(tr, pos, g, v, c, track.length, dict(numSteps=1, numApproxSteps=1, energyOptimal=True, minimumVelocity=1),
                       loss_rows=lm.rows(tr.mass * tr.rho))
    lbz, ubz, lbg, ubg = nlp.bounds(T)
    r = ipm.solve(nlp, nlp.x0(T), lbz, ubz, lbg, ubg)
    assert r.success
    u = nlp.unpack(r.x)
    out = dict(name=name, N=N, T=T, auxiliaries=aux, etaGear=eta_gear, cost_kwh=nlp.cost(r.f), iterations=r.iters, kkt=r.kkt,
               t=u['t'].tolist(), b=u['b'].tolist(), Fel=u['Fel'].tolist(), s=u['s'].tolist())
    json.dump(out, open(os.path.join(HERE, name + '.json'), 'w'))
    print(name, out['cost_kwh'], r.iters)


if __name__ == '__main__':
    dynamic_case('table3_dynamic_flat_N300', FLAT_JSON, 300, 1541.0)
    dynamic_case('dynamic_swiss_N300', SWISS_JSON, 300, 1242.0)
