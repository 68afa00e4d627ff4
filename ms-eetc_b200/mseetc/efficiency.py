"""Speed- and load-dependent loss map of the traction chain (gear, motor + converter, auxiliaries, transformer).

Same public functions as the reference's ``mseetc/efficiency.py`` (``forceToLoad`` :7-12, ``loadToForce`` :15-20,
``createSpline`` :23-51, ``motorLossesFunction`` :54-98, ``totalLossesFunction`` :101-141) but numeric: the
interpolant is the cubic not-a-knot tensor-product B-spline (what CasADi's ``interpolant('bspline')`` builds,
value 0 outside the grid), held as knots + coefficients so that the very same numbers are uploaded to the device.
The returned callables work on floats and numpy arrays and carry ``.mseetc_kind`` / ``.device_params`` so that
``casadiSolver`` can run them in the kernels instead of in Python.
"""
import numpy as np
from scipy.interpolate import make_interp_spline, BSpline

from mseetc.data import dataLosses

_MIN_SPEED, _MAX_SPEED = 20.0, 160.0     # [km/h] of the measured frequency range
_MIN_FREQ, _MAX_FREQ = 20.0, 170.0       # [Hz]
_POW_FREQ = 55.0                         # frequency where maximum power meets maximum force [Hz]
_NUM_MOTORS = 4
TRAFO_R, TRAFO_V = 10.0, 15000.0         # transformer resistance [Ohm], catenary voltage [V]


def hzToKmPerHour(f):
    return ((f - _MIN_FREQ) / (_MAX_FREQ - _MIN_FREQ)) * (_MAX_SPEED - _MIN_SPEED) + _MIN_SPEED


def forceToLoad(force, velocity, forceMax, powerMax):
    "Load [%] of a non-negative force [N]: force-limited below the turning speed, power-limited above."
    turning = powerMax / forceMax
    below = velocity <= turning
    return 100 * (force / forceMax) * below + 100 * (force * velocity / powerMax) * (~below if isinstance(below, np.ndarray) else (not below))


def loadToForce(load, velocity, forceMax, powerMax):
    turning = powerMax / forceMax
    below = velocity <= turning
    return (load / 100) * (forceMax * below + (powerMax / velocity) * (~below if isinstance(below, np.ndarray) else (not below)))


class TensorSpline:
    "Cubic not-a-knot interpolating tensor-product B-spline on a rectangular grid; 0 outside the grid."

    def __init__(self, x, y, values):
        x, y = np.asarray(x, float), np.asarray(y, float)
        values = np.asarray(values, float)                       # [len(x), len(y)]
        sx = make_interp_spline(x, values, k=3, axis=0)          # coefficients along x for every y sample
        sy = make_interp_spline(y, sx.c.T, k=3, axis=0)          # then along y; result [ny, nx]
        self.tx, self.ty, self.coef = sx.t, sy.t, np.ascontiguousarray(sy.c.T)   # knots (n+4), coefficients [nx, ny]
        self.box = (x[0], x[-1], y[0], y[-1])

    def __call__(self, x, y, dx=0, dy=0):
        x, y = np.broadcast_arrays(np.asarray(x, float), np.asarray(y, float))
        inside = (x >= self.box[0]) & (x <= self.box[1]) & (y >= self.box[2]) & (y <= self.box[3])
        xs, ys = np.clip(x, self.box[0], self.box[1]), np.clip(y, self.box[2], self.box[3])
        bx = BSpline.design_matrix(xs.ravel(), self.tx, 3).toarray() if dx == 0 else \
            np.stack([BSpline.basis_element(self.tx[i:i + 5], extrapolate=False).derivative(dx)(xs.ravel()) for i in range(len(self.tx) - 4)], 1)
        by = BSpline.design_matrix(ys.ravel(), self.ty, 3).toarray() if dy == 0 else \
            np.stack([BSpline.basis_element(self.ty[i:i + 5], extrapolate=False).derivative(dy)(ys.ravel()) for i in range(len(self.ty) - 4)], 1)
        bx, by = np.nan_to_num(bx), np.nan_to_num(by)
        val = np.einsum('pi,ij,pj->p', bx, self.coef, by).reshape(x.shape)
        return np.where(inside, val, 0.0)


def createSpline(loads, velocities, losses, forceMax, powerMax):
    "Motor-loss function of (force [N], velocity [m/s]); ``losses`` is load-major flattened like the reference (order='F')."
    loadsLoc = np.array(loads, dtype=float)
    loadsLoc[-1] += 1e-4                                  # keeps load 100.000000001 inside the grid
    velocities = np.asarray(velocities, dtype=float)
    table = np.asarray(losses, dtype=float).reshape(len(velocities), len(loadsLoc)).T
    lut = TensorSpline(loadsLoc, velocities, table)
    vMin, vMax = velocities.min(), velocities.max()

    def spline(f, v):
        f, v = np.asarray(f, dtype=float), np.asarray(v, dtype=float)
        vc = np.clip(v, vMin, vMax)                       # low / high speeds use the first / last measured column
        load = forceToLoad(np.abs(f), vc, forceMax, powerMax)
        out = lut(load, vc)
        return out if out.ndim else float(out)

    spline.lut = lut
    return spline


def motorLossesFunction(train, detailedOutput=False):
    """Motor + converter losses of the whole train [W].  NOTE: like the reference (efficiency.py:64-71) this adapts the
    train to the measured drive: powerMax, powerMin, forceMin (if regenerative braking is on) and velocityMax."""
    forceMax = train.forceMax
    powerMax = forceMax * hzToKmPerHour(_POW_FREQ) / 3.6
    train.powerMax = powerMax
    train.powerMin = -powerMax
    train.forceMin = -forceMax * (train.forceMin != 0)
    train.velocityMax = _MAX_SPEED / 3.6

    cfgA, cfgB = dataLosses()
    best = np.minimum(np.array(cfgA['losses']), np.array(cfgB['losses'])) * _NUM_MOTORS      # [load, frequency]
    speeds = [hzToKmPerHour(f) / 3.6 for f in cfgB['frequencies']]
    fun = createSpline(cfgB['loads'], speeds, best.ravel(order='F'), forceMax, powerMax)
    fun.forceMax, fun.powerMax = forceMax, powerMax
    if not detailedOutput:
        return fun
    import pandas as pd

    def frame(cfg):
        df = pd.DataFrame(index=[hzToKmPerHour(f) / 3.6 for f in cfg['frequencies']])
        for i, load in enumerate(cfg['loads']):
            df[load] = [x * _NUM_MOTORS for x in cfg['losses'][i]]
        return df

    return {'fun': fun, 'dfA': frame(cfgA), 'dfB': frame(cfgB)}


def totalLossesFunction(train, auxiliaries=27000, etaGear=1):
    "Total electrical losses [W] of (force at the wheel [N], velocity [m/s])."
    motor = motorLossesFunction(train)
    R, V = TRAFO_R, TRAFO_V

    def totalLossesFun(f, v):
        f, v = np.asarray(f, dtype=float), np.asarray(v, dtype=float)
        traction = f >= 0
        pWheelTr, pWheelBr = f * v, -f * v
        gear = np.where(traction, ((1 - etaGear) / etaGear) * pWheelTr, (1 - etaGear) * pWheelBr)
        mot = np.asarray(motor(f, v), dtype=float)
        pmTr = pWheelTr + gear + mot + auxiliaries
        pmBr = pWheelBr - gear - mot - auxiliaries          # may be negative (insufficient braking): same formula
        with np.errstate(invalid='ignore'):
            trafo = np.where(traction, (V - np.sqrt(V ** 2 - 4 * R * pmTr)) ** 2 / (4 * R), (V - np.sqrt(V ** 2 + 4 * R * pmBr)) ** 2 / (4 * R))
        total = np.where(mot > 0, gear + mot + auxiliaries + trafo, 0.0)   # outside the measured box the map is 0 by convention
        return total if total.ndim else float(total)

    totalLossesFun.mseetc_kind = 'dynamic'
    totalLossesFun.device_params = dict(auxiliaries=float(auxiliaries), etaGear=float(etaGear), forceMax=float(motor.forceMax),
                                        powerMax=float(motor.powerMax), knots_load=motor.lut.tx.copy(), knots_speed=motor.lut.ty.copy(),
                                        coef=motor.lut.coef.copy(), tableScale=1.0)
    return totalLossesFun
