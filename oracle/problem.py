"""Oracle: problem data (train scalars, track grid) as plain numpy — test infrastructure only.

Restates, without pandas, what the reference does in
  * ``mseetc/utils.py:367-438``  convertUnit
  * ``mseetc/train.py:69-111``   Train field extraction (SI units, rho<1 -> rho+1)
  * ``mseetc/track.py:137-167``  Track JSON import, ``:377-383`` mergeDataFrames,
    ``:420-450`` updateLimits (crop), ``:91-107`` computeDiscretizationPoints
  * ``mseetc/ocp.py:96-116``     specific-unit scalar bounds
"""
import json
from dataclasses import dataclass, field
from pathlib import Path

import numpy as np

_UNIT = {  # utils.py:367-438
    'km': 1e-3, 'km/h': 1 / 3.6, 't': 1e3, '%': 1e-2, 'kW': 1e3, 'MW': 1e6, 'kN': 1e3,
    'kN/(m/s)': 1e3, 'kN/(km/h)': 3.6e3, 'N/(km/h)': 3.6, 'kN/(m/s)^2': 1e3,
    'kN/(km/h)^2': 1e3 * 3.6 ** 2, 'N/(km/h)^2': 3.6 ** 2, 't/m': 1e3,
}
_UNIT_ID = {'m', 'm/s', 'permil', 'kg', 'W', 'N', 'm/s^2', '-', 'N/(m/s)', 'N/(m/s)^2', 'kg/m'}


def convert_unit(value, unit):
    if unit in _UNIT_ID:
        return value
    if unit == 'km':
        return value / 1e3
    if unit == 'km/h':
        return value / 3.6
    if unit == '%':
        return value / 100
    if unit == 'kN/(km/h)':
        return value * 1e3 * 3.6
    if unit == 'kN/(km/h)^2':
        return value * 1e3 * 3.6 ** 2
    if unit in _UNIT:
        return value * _UNIT[unit]
    raise ValueError("Unknown unit: {}!".format(unit))


@dataclass
class TrainData:
    """Attributes of ``mseetc.train.Train`` that the NLP reads (train.py:69-111)."""
    mass: float
    rho: float
    velocityMax: float
    forceMax: float = None
    forceMin: float = None
    forceMinPn: float = None
    powerMax: float = None
    powerMin: float = None
    accMax: float = None
    accMin: float = None
    r0: float = 0.0
    r1: float = 0.0
    r2: float = 0.0
    etaTraction: float = None
    etaRgBrake: float = None
    g: float = 9.81
    # loss model: ('static', etaT, etaR) | ('none',) | ('dynamic', auxiliaries, etaGear, tableScale)
    losses: tuple = None

    def copy(self):
        import copy
        return copy.copy(self)


def load_train(path):
    with open(path) as fh:
        d = json.load(fh)

    def get(key, neg=False):
        if key not in d:
            return None
        v = d[key]['value']
        return convert_unit(-abs(v) if neg else v, d[key]['unit'])

    rho = get('rho')
    if rho < 1:
        rho += 1  # train.py:73-75
    t = TrainData(mass=get('mass'), rho=rho, velocityMax=get('max speed'),
                  forceMax=get('max traction force'), forceMin=get('max reg braking force', True),
                  forceMinPn=get('max pn braking force', True), powerMax=get('max traction power'),
                  powerMin=get('max reg braking power', True), accMax=get('max acceleration'),
                  accMin=get('max deceleration', True), r0=get('rolling resistance r0'),
                  r1=get('rolling resistance r1'), r2=get('rolling resistance r2'),
                  etaTraction=get('efficiency traction'), etaRgBrake=get('efficiency reg brake'))
    if t.etaTraction is not None:
        t.losses = ('static', t.etaTraction, t.etaRgBrake)
    return t


def _step_union(*tables):
    """Outer join + forward fill of step functions given as (pos[], val[]) (track.py:377-383)."""
    pos = np.unique(np.concatenate([p for p, _ in tables]))
    cols = []
    for p, v in tables:
        idx = np.searchsorted(p, pos, side='right') - 1
        col = np.where(idx >= 0, np.asarray(v, float)[np.clip(idx, 0, None)], np.nan)
        cols.append(col)
    return pos, cols


@dataclass
class TrackData:
    length: float
    speedLimits: tuple  # (pos[], v[m/s])
    gradients: tuple    # (pos[], permil)
    curvatures: tuple   # (pos[], 1/m)
    title: str = ''

    def crop(self, positionStart=None, positionEnd=None):
        """track.py:420-450 updateLimits."""
        a = 0.0 if positionStart is None else float(positionStart)
        b = self.length if positionEnd is None else float(positionEnd)
        if (not 0 <= a < self.length) or (not 0 < b <= self.length):
            raise ValueError("Given positions must be between limits of track!")

        def crop1(tab):
            p, v = np.asarray(tab[0], float), np.asarray(tab[1], float)
            pos = np.unique(np.concatenate([p, [a]]))
            idx = np.searchsorted(p, pos, side='right') - 1
            val = v[idx]
            keep = (pos >= a) & (pos <= b)
            return pos[keep] - pos[keep][0], val[keep]

        self.length -= a + (self.length - b)
        self.speedLimits = crop1(self.speedLimits)
        self.gradients = crop1(self.gradients)
        self.curvatures = crop1(self.curvatures)
        return self

    def merged(self):
        """track.py:377-383: rows = union of section starts, values forward-filled."""
        pos, (c, g, v) = _step_union(self.curvatures, self.gradients, self.speedLimits)
        return pos, g, v, c


def load_track(path, constant_curvature=None):
    with open(path) as fh:
        d = json.load(fh)
    length = convert_unit(d['stops']['values'][-1], d['stops']['unit'])
    vu = d['speed limits']['units']['velocity']
    sl = d['speed limits']['values']
    speed = (np.array([p for p, _ in sl], float), np.array([convert_unit(v, vu) for _, v in sl], float))
    gr = d['gradients']['values'] if 'gradients' in d else [(0.0, 0.0)]
    grad = (np.array([p for p, _ in gr], float), np.array([v for _, v in gr], float))
    if 'curvatures' in d:
        # only the constant-radius case is restated here; clothoids are covered by the product's own test
        cu = d['curvatures']['values']
        curv = (np.array([p for p, _, _ in cu], float),
                np.array([(1 / float(r0) + 1 / float(r1)) / 2 for _, r0, r1 in cu], float))
    else:
        curv = (np.array([0.0]), np.array([0.0]))
    if constant_curvature is not None:
        curv = (np.array([0.0]), np.array([float(constant_curvature)]))
    return TrackData(length, speed, grad, curv, d['metadata']['id'])


def discretization_points(track, numIntervals):
    """track.py:91-107: linspace(0, L, N+1-(M-1)) U merged breakpoints, forward filled."""
    mpos, g, v, c = track.merged()
    lin = np.linspace(0, track.length, numIntervals + 1 - (len(mpos) - 1))
    pos = np.unique(np.concatenate([lin, mpos]))
    if len(pos) != numIntervals + 1:
        raise ValueError("Wrong number of computed discretization intervals!")
    idx = np.searchsorted(mpos, pos, side='right') - 1
    return pos, g[idx], v[idx], c[idx]


def curvature_resistance(kappa, g=9.81):
    """train.py:252-253 (Roeckl-type formula, two regimes)."""
    k = np.abs(kappa)
    return np.where(k <= 1 / 300, g * 0.5 * k / (1 - 30 * k), g * 0.65 * k / (1 - 55 * k))
