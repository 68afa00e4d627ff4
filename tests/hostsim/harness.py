"""Test harness: build/run the CPU emulation of the CUDA solver's tick sequence (see hostsim.cpp).

TEST INFRASTRUCTURE ONLY.  Packs the flat C-ABI arrays from an ``oracle.nlp.ReferenceNLP`` so that the
emulated device code and the oracle are fed from one description of the problem.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, '_hostsim.so')
SRC = os.path.join(HERE, 'hostsim.cpp')
CSRC = os.path.join(HERE, '..', '..', 'ms-eetc_b200', 'csrc')

PARAMS = ['SR0', 'SR1', 'SR2', 'FEL_LO', 'FEL_UP', 'FPB_LO', 'POW_LO', 'POW_UP', 'ACC_LO', 'ACC_UP', 'LOSS_TR', 'LOSS_RG',
          'BMIN', 'OBJ_SCALE', 'T_END', 'T_START', 'B_START', 'B_END', 'MASS', 'DYN_AUX', 'DYN_ETAG', 'DYN_FMAX', 'DYN_PMAX', 'DYN_SCALE']


class Problem(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ('n_intervals_max', 'with_pn_brake', 'with_power_rows', 'energy_optimal',
                                               'loss_kind', 'num_steps', 'num_approx_steps', 'max_iterations')] + \
               [('tol', ctypes.c_double), ('mu_init', ctypes.c_double)]


def build():
    deps = [SRC] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cuh')]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.check_call(['g++', '-O2', '-std=c++17', '-shared', '-fPIC', '-o', SO, SRC])
    return ctypes.CDLL(SO)


def pack_instance(nlp, T, t0=0.0, v0=1.0, vN=1.0):
    """(params[PARAM_COUNT], ds, c0, bmax) of one instance from the oracle's NLP object."""
    N = nlp.N
    v0 = min(max(v0, nlp.vmin), nlp.limit[0])
    vN = min(max(vN, nlp.vmin), nlp.limit[-1])
    p = dict(SR0=nlp.sr[0], SR1=nlp.sr[1], SR2=nlp.sr[2], FEL_LO=nlp.forceMin if nlp.withRg else 0.0, FEL_UP=nlp.forceMax,
             FPB_LO=nlp.forceMinPn if nlp.withPn else -1.0, POW_LO=nlp.pLo if nlp.withPower else -1.0,
             POW_UP=nlp.pUp if nlp.withPower else 1.0, ACC_LO=nlp.accMin, ACC_UP=nlp.accMax, LOSS_TR=nlp.cT, LOSS_RG=nlp.cR,
             BMIN=nlp.vmin ** 2, OBJ_SCALE=nlp.scale, T_END=T, T_START=t0, B_START=v0 ** 2, B_END=vN ** 2, MASS=nlp.M,
             DYN_AUX=0.0, DYN_ETAG=1.0, DYN_FMAX=1.0, DYN_PMAX=1.0, DYN_SCALE=1.0)
    if nlp.lossKind == 'dynamic':
        _, aux, etag, scale = nlp.train.losses
        fmax = nlp.train.forceMax
        p.update(DYN_AUX=aux, DYN_ETAG=etag, DYN_FMAX=fmax, DYN_PMAX=fmax * ((((55 - 20) / 150) * 140 + 20) / 3.6), DYN_SCALE=scale)
    bmax = np.zeros(N + 1)
    lim = np.minimum(np.minimum(nlp.limit[1:N], nlp.train.velocityMax), nlp.limit[0:N - 1])
    bmax[1:N] = lim ** 2
    bmax[0] = bmax[N] = 1.0
    return np.array([p[k] for k in PARAMS]), nlp.ds.copy(), nlp.c0.copy(), bmax


def loss_map_arrays():
    "Knots and coefficients of the product's motor-loss spline (what casadiSolver uploads to the device)."
    import sys
    pkg = os.path.join(HERE, '..', '..', 'ms-eetc_b200')
    if pkg not in sys.path:
        sys.path.insert(0, pkg)
    from mseetc.efficiency import motorLossesFunction

    class _T:
        forceMax, forceMin = 213900.0, -1.0
    lut = motorLossesFunction(_T()).lut
    return np.ascontiguousarray(lut.tx), np.ascontiguousarray(lut.ty), np.ascontiguousarray(lut.coef)


def product_tableau(order, scheme):
    import sys
    pkg = os.path.join(HERE, '..', '..', 'ms-eetc_b200')
    if pkg not in sys.path:
        sys.path.insert(0, pkg)
    from mseetc.train import collocationTableau
    A, w, _ = collocationTableau(order, scheme)
    return np.ascontiguousarray(A), np.ascontiguousarray(w)


def solve(nlps, Ts, t0=0.0, v0=1.0, vN=1.0, max_iter=500, tol=1e-8, verbose_inst=-1, lib=None, tmin=None, pit_lanes=0, init_mode=0, mu_init=0.1):
    """Solve instances (one NLP object per instance, equal structure flags) with the emulated device code."""
    lib = lib or build()
    n = len(nlps)
    ref = nlps[0]
    Nmax = max(x.N for x in nlps)
    packs = [pack_instance(x, T, t0, v0, vN) for x, T in zip(nlps, Ts)]
    params = np.ascontiguousarray(np.stack([p[0] for p in packs], axis=1))  # [field][inst]
    nint = np.array([x.N for x in nlps], np.int32)
    trk_of = np.arange(n, dtype=np.int32)
    trk_off = np.concatenate([[0], np.cumsum(nint)]).astype(np.int32)
    ds = np.concatenate([p[1] for p in packs])
    c0 = np.concatenate([p[2] for p in packs])
    bmax = np.concatenate([p[3] for p in packs])
    pr = Problem(Nmax, int(ref.withPn), int(ref.withPower), int(ref.energy), {'none': 0, 'static': 1}.get(ref.lossKind, 2),
                 int(ref.opts['numSteps']), int(ref.opts['numApproxSteps']), max_iter, tol, mu_init)
    stp = 3 + ref.nu
    z = np.zeros((n, Nmax * stp + 2))
    lam = np.zeros((n, Nmax * ref.rows_per))
    obj = np.zeros(n); kkt = np.zeros(n)
    iters = np.zeros(n, np.int32); status = np.zeros(n, np.int32)
    ticks = ctypes.c_int32(0)
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    if ref.lossKind == 'dynamic':
        tl, tv, cf = loss_map_arrays()
        lm = (cf.shape[0], cf.shape[1], P(tl), P(tv), P(cf))
    else:
        lm = (0, 0, ctypes.c_void_p(0), ctypes.c_void_p(0), ctypes.c_void_p(0))
    if ref.opts.get('irk'):                 # collocation integrator: the tableau the product would hand to mseetc_set_integrator
        A, w = product_tableau(*ref.opts['irk'])
        lib.hostsim_set_integrator(len(w), P(A), P(w), 10)
    else:
        lib.hostsim_set_integrator(0, ctypes.c_void_p(0), ctypes.c_void_p(0), 1)
    lib.hostsim_set_integrate_losses(1 if ref.opts.get('integrateLosses') else 0)
    lib.hostsim_solve_batch(ctypes.byref(pr), n, P(params), P(nint), P(trk_of), P(trk_off), P(ds), P(c0), P(bmax),
                            P(np.ascontiguousarray(tmin, dtype=float)) if tmin is not None else ctypes.c_void_p(0), P(z), P(lam),
                            P(obj), P(kkt), P(iters), P(status), verbose_inst, ctypes.byref(ticks), int(pit_lanes), *lm, int(init_mode))
    return dict(z=z, lam=lam, obj=obj, kkt=kkt, iters=iters, status=status, ticks=ticks.value)
