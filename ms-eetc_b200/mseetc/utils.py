"""Host-side helpers of the B200 solver: option containers, unit conversion, loss splitting and the
post-processing that defines the returned trajectory table.

Mirrors the call surface of the reference's ``mseetc/utils.py`` (``Options`` :45-107,
``splitLosses`` :197-220, ``postProcessDataFrame`` :223-336, ``simulateCVODES`` :164-194,
``checkTTOBenchVersion`` :339-364, ``convertUnit`` :367-438) without CasADi: everything is numpy,
vectorised over the rows of the table.
"""
import re
from types import MethodType

import numpy as np

# factor tables instead of an if-chain; same units as utils.py:372-432 of the reference
_IDENTITY_UNITS = frozenset(['m', 'm/s', 'permil', 'kg', 'W', 'N', 'm/s^2', '-', 'N/(m/s)', 'N/(m/s)^2', 'kg/m'])
_SCALE_UNITS = {'t': 1e3, 'kW': 1e3, 'MW': 1e6, 'kN': 1e3, 'kN/(m/s)': 1e3, 'kN/(m/s)^2': 1e3, 't/m': 1e3}


def convertUnit(value, unit):
    "Convert a value given in a TTOBench unit to the SI unit used internally."
    if unit in _IDENTITY_UNITS:
        return value
    if unit in _SCALE_UNITS:
        return value * _SCALE_UNITS[unit]
    if unit == 'km':
        return value / 1e3
    if unit == 'km/h':
        return value / 3.6
    if unit == '%':
        return value / 100
    if unit == 'kN/(km/h)':
        return value * 1e3 * 3.6
    if unit == 'N/(km/h)':
        return value * 3.6
    if unit == 'kN/(km/h)^2':
        return value * 1e3 * 3.6 ** 2
    if unit == 'N/(km/h)^2':
        return value * 3.6 ** 2
    raise ValueError("Unknown unit: {}!".format(unit))


def checkTTOBenchVersion(jsonDict, supportedVersions):
    if not isinstance(supportedVersions, list) or not all(isinstance(x, str) for x in supportedVersions):
        raise TypeError("'supportedVersions' must be specified a list of strings!")
    meta = jsonDict.get('metadata', {}) if isinstance(jsonDict, dict) else {}
    if 'library version' not in meta:
        raise ValueError("Library version not found in json file!")
    found = re.search(r'v([\d.]+)', meta['library version'])
    if not found:
        raise ValueError("Unexpected format of 'library version' in json file!")
    if found.group(1) not in supportedVersions:
        raise ValueError("Import function works only for library versions {}!".format(','.join(supportedVersions)))


def vecToList(x):
    return np.asarray(x, dtype=float).flatten().tolist()


def vecToNum(x):
    lst = vecToList(x)
    if len(lst) > 1:
        raise ValueError("Vector should have only one element!")
    return lst[0]


class Options():
    """Defaults as attributes, user dictionary on top, unknown keys rejected (reference utils.py:45-107)."""

    def __init__(self, paramsDict):
        self.overwriteDefaults(paramsDict)
        self.checkValues()

    def checkValues(self):
        pass

    def checkPositiveInteger(self, num, fieldName, allowZero=True):
        bad = int(num) != num or (num < 0 if allowZero else num <= 0)
        if bad:
            raise ValueError("{} must be a {} positive integer!".format(fieldName, 'strictly' if not allowZero else ''))

    def checkBounds(self, num, fieldName, lowerBound, upperBound):
        if not lowerBound <= num <= upperBound:
            raise ValueError("{} must be between {} and {}!".format(fieldName, lowerBound, upperBound))

    def overwriteDefaults(self, paramsDict):
        for key, val in paramsDict.items():
            if not hasattr(self, key):
                raise ValueError("Specified option ({}) does not exist!".format(key))
            cur = getattr(self, key)
            if isinstance(cur, Options):
                if not isinstance(val, dict):
                    raise ValueError("Nested options must be specified as a dictionary!")
                cur.overwriteDefaults(val)
            else:
                setattr(self, key, val)

    def toDict(self):
        out = {}
        for name in dir(self):
            if name.startswith('__') or name == 'ignoreFields':
                continue
            val = getattr(self, name)
            if type(val) == MethodType:
                continue
            out[name] = val.toDict() if isinstance(val, Options) else val
        return out


# ---------------------------------------------------------------------------------------------
# loss functions
# ---------------------------------------------------------------------------------------------

def _dlosses_df(fun, f, v, h=1e-6):
    "df-derivative of a loss callable; analytic when the callable offers it, central difference otherwise."
    if hasattr(fun, 'dforce'):
        return fun.dforce(f, v)
    return (fun(f + h, v) - fun(f - h, v)) / (2 * h)


def splitLosses(fun):
    """Split a loss map into a traction and a regenerative-brake piece, each smooth at f = 0
    (reference utils.py:197-220): the inactive half-plane is replaced by the tangent at f = +-1e-10."""
    tol = 1e-10

    def funTr(f, v):
        f = np.asarray(f, dtype=float)
        lin = _dlosses_df(fun, tol, v) * f + fun(0.0, v)
        return np.where(f >= 0, fun(f, v), lin)

    def funRgb(f, v):
        f = np.asarray(f, dtype=float)
        lin = _dlosses_df(fun, -tol, v) * f + fun(0.0, v)
        return np.where(f < 0, fun(f, v), lin)

    return funTr, funRgb


def _rolling(train, v, totalMass):
    return (train.r0 + train.r1 * v + train.r2 * v ** 2) / totalMass


def _curve_resistance(train, kappa):
    k = np.abs(kappa)
    return np.where(k <= 1 / 300, train.g * 0.5 * k / ((1 - 30 * k) * train.rho), train.g * 0.65 * k / ((1 - 55 * k) * train.rho))


def _resimulate(time, pos0, vel0, force, grad, curv, train, totalMass, substeps=24):
    """Time-domain re-simulation of the optimal controls (reference: utils.py:110-194, CVODES at
    1e-12/1e-14).  Here: RK4 on (s, v) with Richardson extrapolation (substeps and 2*substeps), errors
    accumulated from interval to interval exactly as simulateCVODES(accumulatedErrors=True)."""
    n = len(time) - 1
    s_out = np.empty(n + 1)
    v_out = np.empty(n + 1)
    s_out[0], v_out[0] = pos0, vel0
    cres = _curve_resistance(train, curv)

    def rk(s, v, f, gd, cr, dt, m):
        h = dt / m
        acc = lambda vv: f - _rolling(train, vv, totalMass) - train.g * gd / train.rho - cr
        for _ in range(m):
            k1s, k1v = v, acc(v)
            k2s, k2v = v + 0.5 * h * k1v, acc(v + 0.5 * h * k1v)
            k3s, k3v = v + 0.5 * h * k2v, acc(v + 0.5 * h * k2v)
            k4s, k4v = v + h * k3v, acc(v + h * k3v)
            s = s + h / 6 * (k1s + 2 * k2s + 2 * k3s + k4s)
            v = v + h / 6 * (k1v + 2 * k2v + 2 * k3v + k4v)
        return s, v

    for i in range(n):
        dt = time[i + 1] - time[i]
        a = rk(s_out[i], v_out[i], force[i], grad[i], cres[i], dt, substeps)
        b = rk(s_out[i], v_out[i], force[i], grad[i], cres[i], dt, 2 * substeps)
        s_out[i + 1] = b[0] + (b[0] - a[0]) / 15
        v_out[i + 1] = b[1] + (b[1] - a[1]) / 15
    return s_out, v_out


def simulateCVODES(dfIn, model, totalMass, accumulatedErrors=True):
    "Re-simulation columns of the table (reference utils.py:164-194); ``model`` is the Train or its exportModel()."
    train = getattr(model, 'train', model)
    t = dfIn.index.values.astype(float)
    pos = dfIn['Position [m]'].values.astype(float)
    vel = dfIn['Velocity [m/s]'].values.astype(float)
    frc = np.nan_to_num(dfIn['Force [N]'].values.astype(float)) / totalMass
    grd = dfIn['Gradient [permil]'].values.astype(float) / 1e3
    crv = dfIn['Curvature [1/m]'].values.astype(float)
    if accumulatedErrors:
        s_sim, v_sim = _resimulate(t, pos[0], vel[0], frc, grd, crv, train, totalMass)
    else:
        s_sim, v_sim = pos.copy(), vel.copy()
        for i in range(len(t) - 1):
            s2, v2 = _resimulate(t[i:i + 2], pos[i], vel[i], frc[i:i + 1], grd[i:i + 1], crv[i:i + 1], train, totalMass)
            s_sim[i + 1], v_sim[i + 1] = s2[1], v2[1]
    dfOut = dfIn.copy()
    dfOut['Position - cvodes [m]'] = s_sim
    dfOut['Velocity - cvodes [m/s]'] = v_sim
    dfOut['Error position [m]'] = np.abs(s_sim - pos)
    dfOut['Error velocity [m/s]'] = np.abs(v_sim - vel)
    return dfOut


def _rk4_richardson(rhs, y0, h, substeps=64, tol=1e-7):
    """Integrate a batch of independent ODEs over [0, h] (arrays, one entry per member; ``rhs(y, sel)`` evaluates the members
    ``sel``).  Classic RK4 with ``substeps`` and with twice as many steps, combined by Richardson extrapolation, for the whole
    batch at once; members whose two runs disagree by more than ``tol`` (a loss map with kinks or jumps along the path, a
    start from very low speed) are redone one by one with an adaptive integrator at rtol 1e-10 -- the reference asks CVODES
    for 1e-8/1e-6 (train.py:396, :437)."""
    everyone = slice(None)

    def run(m):
        y = [np.array(c, dtype=float) for c in y0]
        dt = h / m
        for _ in range(m):
            k1 = rhs(y, everyone)
            k2 = rhs([a + 0.5 * dt * k for a, k in zip(y, k1)], everyone)
            k3 = rhs([a + 0.5 * dt * k for a, k in zip(y, k2)], everyone)
            k4 = rhs([a + dt * k for a, k in zip(y, k3)], everyone)
            y = [a + dt / 6 * (p + 2 * q + 2 * r + w) for a, p, q, r, w in zip(y, k1, k2, k3, k4)]
        return y
    coarse, fine = run(substeps), run(2 * substeps)
    out = [f + (f - c) / 15 for c, f in zip(coarse, fine)]
    rough = np.zeros(np.shape(out[0]), dtype=bool)
    for c, f in zip(coarse, fine):
        rough |= ~(np.abs(f - c) <= tol * np.maximum(np.abs(f), 1e-12))
    if np.any(rough):
        from scipy.integrate import solve_ivp
        for j in np.flatnonzero(rough):
            one = np.array([j])
            single = lambda _, y: [float(np.asarray(d).reshape(-1)[0]) for d in rhs([np.array([v]) for v in y], one)]
            sol = solve_ivp(single, (0.0, float(h[j])), [float(np.asarray(c)[j]) for c in y0], method='LSODA', rtol=1e-10, atol=1e-13)
            for comp, val in zip(out, sol.y[:, -1]):
                comp[j] = val
    return out


def _integrated_losses(time, vel, fel, fpb, grad, curv, train, totalMass):
    """Energy lost in the drive over every interval, integrated along the time-domain trajectory under the interval's constant
    forces (reference utils.py:261-289 with train.py:367-413): d v/dt = a(v), d e/dt = PL(F, v).  The reference integrates
    the traction and the braking branch of splitLosses separately and picks by the sign of the force; each branch is the true
    map on its own side, so this is the unsplit map along the trajectory."""
    loss_fun = train.powerLossesFuns(split=False)                        # specific: W/kg from (N/kg, m/s)
    drive = fel + fpb - train.g * grad / train.rho - _curve_resistance(train, curv)

    def rhs(y, sel):
        v, _ = y
        return [drive[sel] - _rolling(train, v, totalMass), np.asarray(loss_fun(fel[sel], v), dtype=float)]

    _, e = _rk4_richardson(rhs, [vel, np.zeros_like(vel)], time)
    return totalMass * e


def _integrated_rolling_resistance(pos, vel, facc, fpb, grad, curv, train, totalMass):
    """Work of the rolling resistance over every interval in the position domain (reference utils.py:296-320 with
    train.py:415-456): d b/ds = 2 a(b), d e/ds = sr0 + sr1 sqrt(b) + sr2 b, driven -- as in the reference -- by the traction
    part of the electric force and the pneumatic brake only."""
    drive = facc + fpb - train.g * grad / train.rho - _curve_resistance(train, curv)

    def rhs(y, sel):
        b, _ = y
        roll = _rolling(train, np.sqrt(np.maximum(b, 0.0)), totalMass)
        return [2.0 * (drive[sel] - roll), roll]

    _, e = _rk4_richardson(rhs, [vel ** 2, np.zeros_like(vel)], pos)
    return totalMass * e


def postProcessDataFrame(dfIn, points, train, CVODES=True, integrateLosses=False, integrateRollingResistance=False):
    """Derived columns of the trajectory table (reference utils.py:223-336), vectorised over the rows.

    ``integrateLosses=True`` replaces the mid-point loss estimate by the losses integrated along the time-domain trajectory
    (utils.py:261-289), ``integrateRollingResistance=True`` adds the 'Rolling resistance [kWh]' column (utils.py:296-320)."""
    kWh = 1e-6 / 3.6
    totalMass = train.mass * train.rho
    df = dfIn.copy()
    for col in ('Speed limit [m/s]', 'Gradient [permil]', 'Curvature [1/m]'):
        df[col] = points[col].values
    fel = df['Force (el) [N]'].values.astype(float)
    vel = df['Velocity [m/s]'].values.astype(float)
    pos = df['Position [m]'].values.astype(float)
    f_acc = fel * (fel >= 0)           # NaN in the last row propagates, as in the reference
    f_rgb = fel * (fel < 0)
    df['Force (acc) [N]'] = f_acc
    df['Force (rgb) [N]'] = f_rgb
    df['Force [N]'] = f_acc + f_rgb + df['Force (pnb) [N]'].values
    v_next = np.append(vel[1:], np.nan)
    df['Max. Power [kW]'] = np.maximum(f_acc * vel / 1e3, f_acc * v_next / 1e3)
    df['Min. Power [kW]'] = np.minimum(f_rgb * vel / 1e3, f_rgb * v_next / 1e3)
    ds = np.append(np.diff(pos), np.nan)
    loss_fun = train.powerLossesFuns(split=False)          # specific, unsplit (utils.py:247-248)
    v_mid = 0.5 * (vel + v_next)
    last = np.isnan(fel) | np.isnan(v_mid)                   # the terminal row has no control / no mid-point speed
    f_eval, v_eval = np.where(last, 0.0, fel), np.where(last, 1.0, v_mid)
    if not integrateLosses:
        with np.errstate(invalid='ignore', divide='ignore'):
            losses = kWh * ds * totalMass * np.asarray(loss_fun(f_eval / totalMass, v_eval), dtype=float) / v_eval
        losses = np.where(last, np.nan, losses)
    else:
        t = df.index.values.astype(float)
        n = len(t) - 1
        e = _integrated_losses(np.diff(t), vel[:n], fel[:n] / totalMass, df['Force (pnb) [N]'].values[:n].astype(float) / totalMass,
                               df['Gradient [permil]'].values[:n] / 1e3, df['Curvature [1/m]'].values[:n].astype(float), train, totalMass)
        losses = np.append(kWh * e, np.nan)
    df['Losses [kWh]'] = losses
    df['Energy [kWh]'] = kWh * ds * f_acc + kWh * ds * f_rgb + losses
    df['Energy (pnb) [kWh]'] = -kWh * ds * df['Force (pnb) [N]'].values
    df['Energy (kin) [kWh]'] = kWh * 0.5 * train.mass * vel ** 2
    if integrateRollingResistance:
        n = len(pos) - 1
        e = _integrated_rolling_resistance(np.diff(pos), vel[:n], f_acc[:n] / totalMass, df['Force (pnb) [N]'].values[:n].astype(float) / totalMass,
                                           df['Gradient [permil]'].values[:n] / 1e3, df['Curvature [1/m]'].values[:n].astype(float), train, totalMass)
        df['Rolling resistance [kWh]'] = np.append(kWh * e, np.nan)
    grad_res = train.g * (df['Gradient [permil]'].values / 1000) / train.rho
    df['Acceleration [m/s^2]'] = df['Force [N]'].values / totalMass - _rolling(train, vel, totalMass) - grad_res \
        - _curve_resistance(train, df['Curvature [1/m]'].values)
    if CVODES:
        df = simulateCVODES(df, train, totalMass)
    return df


# plotting helpers of the reference (utils.py:441-479) are out of scope: matplotlib is optional
def saveFig(fig, axs, filename):
    if filename is not None:
        import matplotlib.pyplot as plt
        plt.savefig(filename, bbox_inches='tight')


def show():
    import matplotlib.pyplot as plt
    plt.show()


def latexify():
    return False
