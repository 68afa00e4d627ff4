"""Rolling-stock description and the shooting-interval integrator front end.

Same public surface as the reference's ``mseetc/train.py``: ``Train`` (JSON import with unit conversion and
per-field overrides, train.py:11-113; validation :116-172; ``exportModel`` :175-187; ``powerLossesFuns``
:190-219), ``TrainModel`` (:222-277), ``TrainIntegrator`` (:280-364) and the integrator option classes
(:457-534).  There is no CasADi here: ``TrainModel`` carries the numbers the device functions need and
``TrainIntegrator.solve`` evaluates one RK4 / ERK4+ interval through the C ABI
(``mseetc_eval_interval``), i.e. with the very device function the solver uses.
"""
import json
from pathlib import Path

import numpy as np

from mseetc.utils import Options, checkTTOBenchVersion, convertUnit, splitLosses

# JSON key -> (attribute, sign convention).  'neg' fields are stored as non-positive numbers.
_OPTIONAL_FIELDS = [
    ('max traction force', 'forceMax', False),
    ('max reg braking force', 'forceMin', True),
    ('max pn braking force', 'forceMinPn', True),
    ('max traction power', 'powerMax', False),
    ('max reg braking power', 'powerMin', True),
    ('max acceleration', 'accMax', False),
    ('max deceleration', 'accMin', True),
]


class Train():

    def __init__(self, config, pathJSON=Path(__file__).parent.parent / 'trains') -> None:
        self.g = 9.81   # [m/s^2]
        if not isinstance(config, dict):
            raise ValueError("Train configuration should be provided as a dictionary!")
        if 'id' not in config:
            raise ValueError("Train ID must be specified in configuration!")
        with open(Path(pathJSON) / (config['id'] + '.json')) as fh:
            data = json.load(fh)
        checkTTOBenchVersion(data, ['1.1', '1.2', '1.3'])

        # per-field overrides: None removes a limit, {'unit','value'} replaces it.  NOTE: like the reference
        # (train.py:42) the 'id' key is consumed from the caller's dictionary.
        config.pop('id')
        mayBeAdded = ("max acceleration", "max deceleration")
        accepted = set()
        for key, val in config.items():
            if val is None and key in data:
                del data[key]
                accepted.add(key)
                continue
            if not isinstance(val, dict) or val.keys() != {'unit', 'value'}:
                raise ValueError("Configuration field '{}' should be specified as a dictionary with 'unit' and 'value' keys!".format(key))
            if key in data or key in mayBeAdded:
                data[key] = val
                accepted.add(key)
        if set(config) != accepted:
            raise ValueError("Redundant fields in train configuration: {}!".format(', '.join(set(config) - accepted)))

        quantity = lambda key, sign=1.0: convertUnit(sign * abs(data[key]['value']) if sign < 0 else data[key]['value'], data[key]['unit'])
        self.mass = quantity('mass')                       # [kg]
        self.rho = quantity('rho')                         # rotating mass factor
        if self.rho < 1:
            self.rho += 1                                  # "6 %" means 1.06
        self.velocityMax = quantity('max speed')           # [m/s]
        for key, attr, negative in _OPTIONAL_FIELDS:
            setattr(self, attr, quantity(key, -1.0 if negative else 1.0) if key in data else None)
        self.r0 = quantity('rolling resistance r0')        # [N]
        self.r1 = quantity('rolling resistance r1')        # [N/(m/s)]
        self.r2 = quantity('rolling resistance r2')        # [N/(m/s)^2]

        hasT, hasR = 'efficiency traction' in data, 'efficiency reg brake' in data
        if hasT or hasR:
            if not (hasT and hasR):
                raise ValueError("Both efficiencies need to be specified in json file!")
            self.etaTraction = quantity('efficiency traction')
            self.etaRgBrake = quantity('efficiency reg brake')
        self.checkFields()

    def checkFields(self):
        finitePos = lambda x: x is not None and x > 0 and not np.isinf(x)
        if self.mass is None or self.mass < 0 or np.isinf(self.mass):
            raise ValueError("Train mass must be a positive number, not {}!".format(self.mass))
        if self.g is None or not 9 <= self.g <= 10:
            raise ValueError("Acceleration of gravity must be between 9 and 10 m/s^2, not {}!".format(self.g))
        if self.rho is None or not 1 <= self.rho <= 1.5:
            raise ValueError("Rotation mass factor must be between 1 and 1.5, not {}!".format(self.rho))
        if not finitePos(self.velocityMax):
            raise ValueError("Maximum velocity must be a strictly positive number, not {}!".format(self.velocityMax))
        if self.forceMax is not None and not finitePos(self.forceMax):
            raise ValueError("Maximum traction force must be strictly positive or free (None), not {}!".format(self.forceMax))
        if self.forceMinPn is not None and (self.forceMinPn > 0 or np.isinf(self.forceMinPn)):
            raise ValueError("Maximum pneumatic braking force must be negative, zero or free (None), not {}!".format(self.forceMinPn))
        if self.forceMin is not None and (self.forceMin > 0 or np.isinf(self.forceMin)):
            raise ValueError("Maximum regenerative braking force must be negative, zero or free (None), not {}!".format(self.forceMin))
        if self.forceMin == 0 and self.forceMinPn == 0:
            raise ValueError("Both brakes cannot be deactivated simultaneously!")
        if self.powerMax is not None and not finitePos(self.powerMax):
            raise ValueError("Maximum traction power must be strictly positive or free (None), not {}!".format(self.powerMax))
        if self.powerMin is not None and (self.powerMin >= 0 or np.isinf(self.powerMin)):
            raise ValueError("Maximum regenerative brake power must be strictly negative or free (None), not {}!".format(self.powerMin))
        if self.accMax is not None and not finitePos(self.accMax):
            raise ValueError("Maximum acceleration must be strictly positive or free (None), not {}!".format(self.accMax))
        if self.accMin is not None and (self.accMin >= 0 or np.isinf(self.accMin)):
            raise ValueError("Maximum deceleration must be strictly negative or free (None), not {}!".format(self.accMin))
        for name in ('r0', 'r1', 'r2'):
            coef = getattr(self, name)
            if coef is None or coef < 0:
                raise ValueError("Rolling resistance coefficient {} must be positive, not {}!".format(name, coef))

    def exportModel(self):
        "Specific (per kg of rotating mass) model data for the integrator."
        M = self.mass * self.rho
        model = TrainModel(self.r0 / M, self.r1 / M, self.r2 / M, self.rho, self.g, self.forceMinPn != 0)
        model.train = self
        return model

    def powerLossesFuns(self, split=True):
        """Specific power-loss function(s) [W/kg] of (specific force, velocity): the explicit ``powerLosses``
        attribute if present, otherwise the two constant efficiencies."""
        if hasattr(self, 'powerLosses'):
            absolute = self.powerLosses
        elif hasattr(self, 'etaTraction') and hasattr(self, 'etaRgBrake'):
            etaT, etaR = self.etaTraction, self.etaRgBrake
            absolute = lambda f, v: f * v * (f > 0) * (1 - etaT) / etaT - (1 - etaR) * f * v * (f < 0)
        else:
            raise ValueError("Power losses function of train must by either explicitly or implicitly defined!")
        M = self.mass * self.rho

        def specific(f, v):
            return (1 / M) * absolute(f * M, v)

        if hasattr(absolute, 'dforce'):
            specific.dforce = lambda f, v: absolute.dforce(f * M, v)
        if not split:
            return specific
        return splitLosses(specific)


class TrainModel():
    "Numbers that define the ODE in the position domain: db/ds = 2a, dt/ds = 1/sqrt(b)."

    def __init__(self, sr0, sr1, sr2, rho=1, g=9.81, withPnBrake=True) -> None:
        self.sr0, self.sr1, self.sr2 = sr0, sr1, sr2
        self.rho = rho
        self.g = g
        self.withPnBrake = withPnBrake

    def curvatureResistance(self, curvature):
        k = abs(curvature)
        return self.g * 0.5 * k / (1 - 30 * k) if k <= 1 / 300 else self.g * 0.65 * k / (1 - 55 * k)

    def rollingResistanceFun(self, velocitySquared):
        return self.sr0 + self.sr1 * np.sqrt(velocitySquared) + self.sr2 * velocitySquared

    def accelerationFun(self, x, u, gradient, curvature):
        "a(x,u) [m/s^2] with x = (time, velocity^2), u = (traction[, pnBrake])."
        u = np.atleast_1d(np.asarray(u, dtype=float))
        return float(u.sum()) - self.rollingResistanceFun(x[1]) - self.g * gradient / self.rho - self.curvatureResistance(curvature) / self.rho

    def offset(self, gradient, curvature):
        "c0 of the device functions: gravity + curve resistance, specific."
        return self.g * gradient / self.rho + self.curvatureResistance(curvature) / self.rho


def collocationTableau(order, method):
    """Runge-Kutta coefficients (A, w, c) of the collocation method on `order` points: c = the Radau ('radau', right end point
    included) or Gauss-Legendre ('legendre') points on (0, 1] that ca.simpleIRK uses (reference train.py:310), A_ij = int_0^{c_i} l_j,
    w_j = int_0^1 l_j with the Lagrange basis l_j on c.  The device integrates with (A, w): same scheme as CasADi's collocation
    equations, written as an implicit Runge-Kutta method."""
    from numpy.polynomial import legendre as L
    d = int(order)
    if method == 'legendre':
        x = L.legroots([0] * d + [1])
    elif method == 'radau':
        co = np.zeros(d + 1); co[d - 1] = 1.0; co[d] = -1.0
        x = L.legroots(co)
    else:
        raise ValueError("Unknown collocation method: {}!".format(method))
    c = np.sort((np.real(x) + 1.0) / 2.0)
    # integrals of the Lagrange basis by a 16-point Gauss rule (exact for these polynomials), the basis evaluated in product form
    # (expanding it into monomial coefficients loses four digits at nine points)
    xq, wq = L.leggauss(16)
    xq, wq = (xq + 1.0) / 2.0, wq / 2.0

    def basis(j, tau):
        v = np.ones_like(tau)
        for r in range(d):
            if r != j:
                v = v * (tau - c[r]) / (c[j] - c[r])
        return v
    w = np.array([wq @ basis(j, xq) for j in range(d)])
    A = np.array([[c[i] * (wq @ basis(j, c[i] * xq)) for j in range(d)] for i in range(d)])
    return A, w, c


# integrationMethod 'CVODES' (reference train.py:312-322: adaptive BDF on (t, b) with relTol 1e-6 / absTol 1e-8 by default): the
# device takes four Gauss-Legendre collocation steps with four points (order 8) per shooting interval -- on the intervals of the
# TTOBench tracks its error is below 1e-10 relative, i.e. four orders of magnitude inside the default tolerance of CVODES, so the
# two agree to within CVODES's own error; the sensitivities are the exact derivatives of that scheme (CasADi: CVODES forward /
# adjoint sensitivities at the same tolerances).
CVODES_EQUIVALENT = dict(order=4, collMethod='legendre', numSteps=4, numApproxSteps=0, maxIter=20)


def integratorSetup(method, opts):
    "(numSteps, numApproxSteps, tableau or None) that the device integrator needs for the reference's three integration methods."
    if method == 'RK':
        return int(opts.numSteps), int(opts.numApproxSteps), None
    if method == 'IRK':
        A, w, _ = collocationTableau(opts.order, opts.collMethod)
        return int(opts.numSteps), int(opts.numApproxSteps), dict(A=A, w=w, maxIter=int(opts.maxIter))
    if method == 'CVODES':
        q = CVODES_EQUIVALENT
        A, w, _ = collocationTableau(q['order'], q['collMethod'])
        return q['numSteps'], q['numApproxSteps'], dict(A=A, w=w, maxIter=q['maxIter'])
    raise ValueError("Unknown integration method!")


class TrainIntegrator():
    "One shooting interval on the device: explicit RK4 (OptionsRK), collocation (OptionsIRK) or the CVODES-equivalent scheme."

    def __init__(self, model, solver, optsDict={}) -> None:
        if solver not in {'RK', 'IRK', 'CVODES'}:
            raise ValueError("Unknown integration method!")
        self.model = model
        self.opts = {'RK': OptionsRK, 'IRK': OptionsIRK, 'CVODES': OptionsCVODES}[solver](optsDict)
        self.numSteps, self.numApproxSteps, self.tableau = integratorSetup(solver, self.opts)

    def solve(self, time, velocitySquared, ds, traction=0, pnBrake=0, gradient=0, curvature=0):
        if not self.model.withPnBrake and pnBrake != 0:
            raise ValueError("Cannot define value for pneumatic braking when this brake is deactivated!")
        from mseetc import _cabi
        m = self.model
        F = traction + (pnBrake if m.withPnBrake else 0)
        out = _cabi.eval_interval(np.array([[velocitySquared], [F], [ds], [m.offset(gradient, curvature)], [m.sr0], [m.sr1], [m.sr2]], dtype=float),
                                  self.numSteps, self.numApproxSteps, self.tableau)
        return {'time': time + float(out[0, 0]), 'velSquared': float(out[6, 0])}


class OptionsRK(Options):

    def __init__(self, paramsDict):
        self.order = 4
        self.numSteps = 1          # RK4 steps per shooting interval
        self.numApproxSteps = 0    # > 0: time from the average-speed rule on that many sub-points
        super().__init__(paramsDict)

    def checkValues(self):
        if self.order != 4:
            raise ValueError("Only explicit Runge-Kutta of order 4 is currently implemented in casadi!")
        self.checkPositiveInteger(self.numSteps, 'Number of integration steps', allowZero=False)
        self.checkPositiveInteger(self.numApproxSteps, 'Number of time approximation steps', allowZero=True)


class OptionsIRK(Options):

    def __init__(self, paramsDict):
        self.order = 2
        self.numSteps = 1
        self.numApproxSteps = 0
        self.collMethod = 'radau'
        self.maxIter = 10
        self.jit = False
        super().__init__(paramsDict)

    def checkValues(self):
        if int(self.order) != self.order or not 1 <= self.order <= 9:
            raise ValueError("Order of implicit Runge-Kutta should be a positive integer between 1 and 9!")
        self.checkPositiveInteger(self.numSteps, 'Number of integration steps', allowZero=False)
        self.checkPositiveInteger(self.numApproxSteps, 'Number of time approximation steps', allowZero=True)
        if self.collMethod not in {'radau', 'legendre'}:
            raise ValueError("Unknown collocation method: {}!".format(self.collMethod))
        self.checkPositiveInteger(self.maxIter, 'Maximum number of iterations', allowZero=False)
        if not isinstance(self.jit, bool):
            raise ValueError("JIT option must be a boolean!")


class OptionsCVODES(Options):

    def __init__(self, paramsDict):
        self.absTol = 1e-8
        self.relTol = 1e-6
        super().__init__(paramsDict)

    def checkValues(self):
        self.checkBounds(self.absTol, 'Absolute tolerance', 1e-20, 1e-1)
        self.checkBounds(self.relTol, 'Relative tolerance', 1e-20, 1e-1)
