"""ctypes binding of libmseetc_b200.so (include/mseetc_b200.h).

torch is used only to own device buffers and the CUDA stream; every call crosses the C ABI with raw
pointers.  There is no CPU fallback: without the shared library or without a CUDA device every entry point
raises RuntimeError.
"""
import ctypes
import functools
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('MSEETC_B200_LIB', os.path.join(_HERE, 'libmseetc_b200.so'))   # override: tuning experiments only

PARAMS = ['SR0', 'SR1', 'SR2', 'FEL_LO', 'FEL_UP', 'FPB_LO', 'POW_LO', 'POW_UP', 'ACC_LO', 'ACC_UP', 'LOSS_TR', 'LOSS_RG',
          'BMIN', 'OBJ_SCALE', 'T_END', 'T_START', 'B_START', 'B_END', 'MASS', 'DYN_AUX', 'DYN_ETAG', 'DYN_FMAX', 'DYN_PMAX',
          'DYN_SCALE']
PARAM_INDEX = {name: i for i, name in enumerate(PARAMS)}

# relative margin below the minimum trip time inside which an instance is still iterated (MS_TMIN_MARGIN, core.cuh)
TMIN_MARGIN = 1e-6

STATUS_STRINGS = {   # IPOPT's return_status vocabulary (what stats['Solver status'] holds in the reference, ocp.py:362)
    0: 'Solve_Succeeded',
    1: 'Maximum_Iterations_Exceeded',
    2: 'Restoration_Failed',
    3: 'Error_In_Step_Computation',
    4: 'Infeasible_Problem_Detected',
    5: 'Invalid_Number_Detected',
    6: 'Solved_To_Acceptable_Level',
}
SUCCESS_CODES = (0, 6)      # what CasADi's stats()['success'] is true for (reference ocp.py:364)


class Problem(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ('n_intervals_max', 'with_pn_brake', 'with_power_rows', 'energy_optimal',
                                               'loss_kind', 'num_steps', 'num_approx_steps', 'max_iterations')] + \
               [('tol', ctypes.c_double), ('mu_init', ctypes.c_double), ('initial_guess', ctypes.c_int32), ('stall_iterations', ctypes.c_int32)]


_lib = None


def lib():
    "Load the shared library (built by __graft_entry__.build()); loud failure if it is missing."
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libmseetc_b200.so not found at {} -- build it with `python __graft_entry__.py` "
                           "(nvcc, sm_100a); there is no CPU fallback".format(LIB_PATH))
    L = ctypes.CDLL(LIB_PATH)
    vp, i32, sz = ctypes.c_void_p, ctypes.c_int32, ctypes.c_size_t
    L.mseetc_version.restype = ctypes.c_int
    L.mseetc_last_error.restype = ctypes.c_char_p
    L.mseetc_create.argtypes = [ctypes.POINTER(Problem), ctypes.POINTER(vp)]
    L.mseetc_destroy.argtypes = [vp]
    L.mseetc_workspace_bytes.argtypes = [vp, i32]
    L.mseetc_workspace_bytes.restype = sz
    L.mseetc_solve_batch.argtypes = [vp, i32] + [vp] * 14 + [vp, sz, vp]
    L.mseetc_table_columns.restype = ctypes.c_int
    L.mseetc_postprocess_batch.argtypes = [vp, i32] + [vp] * 8 + [sz] + [vp] * 4
    L.mseetc_last_ticks.argtypes = [vp]
    L.mseetc_last_launches.argtypes = [vp]
    L.mseetc_set_profiling.argtypes = [vp, ctypes.c_int]
    L.mseetc_last_profile.argtypes = [vp, vp, vp, vp]
    L.mseetc_bytes_per_cell.argtypes = [vp, ctypes.c_int]
    L.mseetc_last_timeline.argtypes = [vp, vp, vp, i32]
    if hasattr(L, 'mseetc_measure_fp64_peak'):      # absent from older builds loaded through MSEETC_B200_LIB (tuning experiments)
        L.mseetc_measure_fp64_peak.argtypes = [ctypes.POINTER(ctypes.c_double), vp]
    L.mseetc_bytes_per_cell.restype = ctypes.c_double
    L.mseetc_eval_interval.argtypes = [i32, i32, i32, vp, vp, vp]
    L.mseetc_eval_interval_irk.argtypes = [i32, i32, i32, i32, vp, vp, i32, vp, vp, vp]
    L.mseetc_set_integrator.argtypes = [vp, i32, vp, vp, i32]
    L.mseetc_set_integrate_losses.argtypes = [vp, ctypes.c_int]
    L.mseetc_set_loss_map.argtypes = [vp, i32, i32, vp, vp, vp]
    L.mseetc_set_sweep_lanes.argtypes = [vp, ctypes.c_int]
    L.mseetc_last_sweep_fallbacks.argtypes = [vp]
    L.mseetc_last_sweep_fallbacks.restype = ctypes.c_longlong
    L.mseetc_last_sweep_lanes.argtypes = [vp]
    L.mseetc_last_compactions.argtypes = [vp]
    L.mseetc_set_compaction.argtypes = [vp, ctypes.c_int]
    L.mseetc_last_compaction_moves.argtypes = [vp]
    L.mseetc_last_sweep_fallback_reasons.argtypes = [vp, ctypes.POINTER(ctypes.c_int32)]
    L.mseetc_eval_loss_rows.argtypes = [vp, i32, vp, vp, vp, vp]
    _lib = L
    return L


def _torch_cuda():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("mseetc_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch


def _check(rc, what):
    if rc != 0:
        raise RuntimeError("{} failed ({}): {}".format(what, rc, lib().mseetc_last_error().decode()))


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


class Handle:
    "Owns one mseetc_handle plus the device workspace (a torch uint8 tensor, re-used across solves)."

    def __init__(self, n_intervals_max, with_pn, with_power, energy, loss_kind, num_steps, num_approx, max_iter,
                 tol=1e-8, mu_init=0.1, initial_guess=0, stall_iterations=0):
        self.problem = Problem(int(n_intervals_max), int(with_pn), int(with_power), int(energy), int(loss_kind),
                               int(num_steps), int(num_approx), int(max_iter), float(tol), float(mu_init), int(initial_guess), int(stall_iterations))
        self._h = ctypes.c_void_p(0)
        _check(lib().mseetc_create(ctypes.byref(self.problem), ctypes.byref(self._h)), 'mseetc_create')
        self._ws = None
        self.nu = 1 + int(with_pn)
        self.rows = (2 if with_power else 0) + 3 + (2 if energy else 0)

    def __del__(self):
        try:
            if self._h:
                lib().mseetc_destroy(self._h)
                self._h = ctypes.c_void_p(0)
        except Exception:
            pass

    def set_integrator(self, A, w, max_newton=10):
        "Collocation steps with the Runge-Kutta coefficients (A, w) instead of explicit RK4 (A = None: back to RK4)."
        if A is None:
            _check(lib().mseetc_set_integrator(self._h, 0, None, None, 1), 'mseetc_set_integrator')
            return
        A = np.ascontiguousarray(A, dtype=np.float64); w = np.ascontiguousarray(w, dtype=np.float64)
        _check(lib().mseetc_set_integrator(self._h, int(len(w)), A.ctypes.data_as(ctypes.c_void_p), w.ctypes.data_as(ctypes.c_void_p),
                                           int(max_newton)), 'mseetc_set_integrator')

    def set_integrate_losses(self, on):
        "integrateLosses = True of the reference (ocp.py:231-241): epigraph rows on the loss energies integrated over each interval"
        _check(lib().mseetc_set_integrate_losses(self._h, 1 if on else 0), 'mseetc_set_integrate_losses')

    def set_sweep_lanes(self, lanes):
        "0 = chosen per call; 1 = sequential Riccati sweeps; 8 / 16 / 32 = parallel-in-time sweeps with that many chunk lanes per instance."
        _check(lib().mseetc_set_sweep_lanes(self._h, int(lanes)), 'mseetc_set_sweep_lanes')
        self._lanes = int(lanes)

    def sweep_lanes(self):
        "lanes the last solve on this handle ran with (the setting, before the first solve)"
        n = int(lib().mseetc_last_sweep_lanes(self._h))
        return n if getattr(self, '_solved', False) else getattr(self, '_lanes', 1)

    def set_compaction(self, on):
        "Compaction of the running batch (csrc/compact.cuh); on by default, results do not depend on it."
        _check(lib().mseetc_set_compaction(self._h, int(bool(on))), 'mseetc_set_compaction')

    def last_compactions(self):
        "(passes launched, instances moved) in the last solve"
        return int(lib().mseetc_last_compactions(self._h)), int(lib().mseetc_last_compaction_moves(self._h))

    def last_sweep_fallbacks(self):
        return int(lib().mseetc_last_sweep_fallbacks(self._h))

    def last_sweep_fallback_reasons(self):
        "(reference recursion failed, chain step singular, chain and recursion disagreed) of the last solve"
        out = (ctypes.c_int32 * 3)()
        _check(lib().mseetc_last_sweep_fallback_reasons(self._h, out), 'mseetc_last_sweep_fallback_reasons')
        return tuple(int(x) for x in out)

    def set_loss_map(self, knots_load, knots_speed, coef):
        "Upload the motor-loss spline (efficiency.createSpline) used by loss_kind 2."
        tl = np.ascontiguousarray(knots_load, dtype=np.float64)
        tv = np.ascontiguousarray(knots_speed, dtype=np.float64)
        cf = np.ascontiguousarray(coef, dtype=np.float64)
        assert cf.shape == (len(tl) - 4, len(tv) - 4)
        _torch_cuda()
        _check(lib().mseetc_set_loss_map(self._h, cf.shape[0], cf.shape[1], tl.ctypes.data, tv.ctypes.data, cf.ctypes.data),
               'mseetc_set_loss_map')

    def eval_loss_rows(self, fel, b0, b1, params):
        "Kernel-level parity hook: numpy in ([n] each, params [PARAM_COUNT, n]) -> numpy [20, n]."
        torch = _torch_cuda()
        inp = torch.from_numpy(np.ascontiguousarray(np.stack([fel, b0, b1]), dtype=np.float64)).cuda()
        par = torch.from_numpy(np.ascontiguousarray(params, dtype=np.float64)).cuda()
        out = torch.empty((20, inp.shape[1]), dtype=torch.float64, device=inp.device)
        _check(lib().mseetc_eval_loss_rows(self._h, inp.shape[1], _ptr(inp), _ptr(par), _ptr(out),
                                           ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), 'mseetc_eval_loss_rows')
        return out.cpu().numpy()

    def workspace(self, n, device):
        torch = _torch_cuda()
        need = lib().mseetc_workspace_bytes(self._h, int(n))
        if self._ws is None or self._ws.numel() < need or self._ws.device != device:
            self._ws = torch.empty(need, dtype=torch.uint8, device=device)
        return self._ws, need

    def solve_device(self, params, nint, trk_of, trk_off, ds, c0, bmax, want_z=True, want_lam=False, tmin=None, out=None):
        """All arguments are torch CUDA tensors (float64 / int32).  Returns dict of device tensors."""
        torch = _torch_cuda()
        n = int(nint.numel())
        dev = params.device
        Nmax = self.problem.n_intervals_max
        stp = 3 + self.nu
        out = out if out is not None else dict(
            z=torch.zeros((n, Nmax * stp + 2), dtype=torch.float64, device=dev) if want_z else None,
            lam=torch.zeros((n, Nmax * self.rows), dtype=torch.float64, device=dev) if want_lam else None,
            obj=torch.empty(n, dtype=torch.float64, device=dev),
            kkt=torch.empty(n, dtype=torch.float64, device=dev),
            iters=torch.empty(n, dtype=torch.int32, device=dev),
            status=torch.empty(n, dtype=torch.int32, device=dev),
        )
        ws, need = self.workspace(n, dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        rc = lib().mseetc_solve_batch(self._h, n, _ptr(params), _ptr(nint), _ptr(trk_of), _ptr(trk_off), _ptr(ds), _ptr(c0),
                                      _ptr(bmax), _ptr(tmin), _ptr(out['z']), _ptr(out['lam']), _ptr(out['obj']), _ptr(out['kkt']),
                                      _ptr(out['iters']), _ptr(out['status']), _ptr(ws), need, ctypes.c_void_p(stream))
        _check(rc, 'mseetc_solve_batch')
        self._solved = True
        out['ticks'] = lib().mseetc_last_ticks(self._h)
        out['launches'] = lib().mseetc_last_launches(self._h)
        return out


TABLE_COLUMNS = ('Time [s]', 'Position [m]', 'Velocity [m/s]', 'Force (el) [N]', 'Force (pnb) [N]', 'Slacks', 'Speed limit [m/s]',
                 'Gradient [permil]', 'Curvature [1/m]', 'Force (acc) [N]', 'Force (rgb) [N]', 'Force [N]', 'Max. Power [kW]', 'Min. Power [kW]',
                 'Losses [kWh]', 'Energy [kWh]', 'Energy (pnb) [kWh]', 'Energy (kin) [kWh]', 'Acceleration [m/s^2]', 'Position - cvodes [m]',
                 'Velocity - cvodes [m/s]', 'Error position [m]', 'Error velocity [m/s]')      # reference ocp.py:401-405, utils.py:230-334,188-192


def postprocess_device(handle, z, params, nint, trk_of, trk_off, ds, c0, nodes, mass, status, out=None):
    "Batched trajectory tables on the device (mseetc_postprocess_batch); all arguments torch CUDA tensors; returns [n, Nmax+1, columns]."
    torch = _torch_cuda()
    n = int(nint.numel())
    ncol = lib().mseetc_table_columns()
    assert ncol == len(TABLE_COLUMNS)
    Nmax = handle.problem.n_intervals_max
    if out is None:
        out = torch.empty((n, Nmax + 1, ncol), dtype=torch.float64, device=z.device)
    stream = torch.cuda.current_stream(z.device).cuda_stream
    _check(lib().mseetc_postprocess_batch(handle._h, n, _ptr(z), _ptr(params), _ptr(nint), _ptr(trk_of), _ptr(trk_off), _ptr(ds), _ptr(c0),
                                          _ptr(nodes), int(nodes.shape[1]), _ptr(mass), _ptr(status), _ptr(out), ctypes.c_void_p(stream)),
           'mseetc_postprocess_batch')
    return out


class StreamPool:
    """Solve one batch as k sub-batches on k CUDA streams, each driven by its own host thread (ctypes releases the GIL).

    A tick of the solver is a chain of latency-bound phases -- the Riccati sweep kernel takes the same time for 1 or for 4096
    instances -- so concurrent sub-batches overlap one sub-batch's sweep with the interval kernels of the others.  Instances
    are independent: the results are bitwise identical to a single-stream solve."""

    def __init__(self, make_handle, k, device):
        torch = _torch_cuda()
        self.handles = [make_handle() for _ in range(k)]
        # graded priorities: sub-batches of equal size would otherwise march in phase (all in their bandwidth-bound interval
        # kernels together, then all in their latency-bound sweeps together); with priorities the first stream runs as if alone
        # and the others fill the gaps its sweeps leave
        self.streams = [torch.cuda.Stream(device=device, priority=max(-4, -(k - 1 - i))) for i in range(k)]
        self.device = device

    @staticmethod
    def bounds(n, k):
        per = -(-n // k)
        per = (per + 31) // 32 * 32                     # sub-batches start on a 32-instance tile
        cuts = [min(i * per, n) for i in range(k + 1)]
        cuts[-1] = n
        return [(a, b) for a, b in zip(cuts[:-1], cuts[1:]) if b > a]

    @staticmethod
    def interleave(n, k, tile=32):
        "Cached front end of _interleave (pure function of its arguments; the arrays are shared, do not modify them)."
        return StreamPool._interleave(int(n), int(k), int(tile))[:2]

    @staticmethod
    def interleave_inverse(n, k, tile=32):
        "argsort of the permutation: position of every instance of the caller's order in the dealt order."
        return StreamPool._interleave(int(n), int(k), int(tile))[2]

    @staticmethod
    @functools.lru_cache(maxsize=16)
    def _interleave(n, k, tile):
        perm, parts = StreamPool._interleave_build(n, k, tile)
        return perm, parts, np.argsort(perm)

    @staticmethod
    def _interleave_build(n, k, tile=32):
        """Order of the instances that balances the sub-batches: 32-instance tiles are dealt round-robin to the k streams, so a
        contiguous run of cheap instances (a sorted sweep whose short trip times are screened as infeasible) is shared by all
        streams while screened tiles stay whole (their warps exit at once).  Returns (perm, parts): sub-batch i is
        perm[parts[i][0]:parts[i][1]] of the caller's order."""
        tiles = [np.arange(t, min(t + tile, n)) for t in range(0, n, tile)]
        groups = [[] for _ in range(k)]
        for j, t in enumerate(tiles[:-1] if len(tiles[-1]) < tile else tiles):
            groups[j % k].append(t)
        if len(tiles[-1]) < tile:
            groups[-1].append(tiles[-1])              # the ragged tile goes last so that every sub-batch starts on a tile
        perm, parts, pos = [], [], 0
        for g in groups:
            if not g:
                continue
            idx = np.concatenate(g)
            perm.append(idx)
            parts.append((pos, pos + len(idx)))
            pos += len(idx)
        return np.concatenate(perm), parts

    def solve(self, params, nint, trk_of, trk_off, ds, c0, bmax, tmin=None, out=None, want_lam=False, parts=None, on_started=None):
        import threading
        torch = _torch_cuda()
        n = int(nint.numel())
        dev = params.device
        h0 = self.handles[0]
        Nmax, stp = h0.problem.n_intervals_max, 3 + h0.nu
        if out is None:
            out = dict(z=torch.zeros((n, Nmax * stp + 2), dtype=torch.float64, device=dev),
                       lam=torch.zeros((n, Nmax * h0.rows), dtype=torch.float64, device=dev) if want_lam else None,
                       obj=torch.empty(n, dtype=torch.float64, device=dev), kkt=torch.empty(n, dtype=torch.float64, device=dev),
                       iters=torch.empty(n, dtype=torch.int32, device=dev), status=torch.empty(n, dtype=torch.int32, device=dev))
        main = torch.cuda.current_stream(dev)
        parts = parts if parts is not None else self.bounds(n, len(self.handles))
        info, errors = [None] * len(parts), []

        import time as _t
        t_ref = _t.perf_counter()
        spans = [None] * len(parts)

        def work(i, a, b):
            try:
                t_in = _t.perf_counter() - t_ref
                torch.cuda.set_device(dev)
                st = self.streams[i]
                st.wait_stream(main)
                with torch.cuda.stream(st):
                    sub = {k: (v[a:b] if v is not None else None) for k, v in out.items()}
                    r = self.handles[i].solve_device(params[:, a:b].contiguous(), nint[a:b], trk_of[a:b], trk_off, ds, c0, bmax,
                                                     tmin=tmin[a:b] if tmin is not None else None, out=sub)
                    info[i] = (r['ticks'], r['launches'])
                spans[i] = (round(1e3 * t_in, 2), round(1e3 * (_t.perf_counter() - t_ref), 2))
            except Exception as exc:
                errors.append(exc)

        threads = [threading.Thread(target=work, args=(i, a, b)) for i, (a, b) in enumerate(parts)]
        for t in threads:
            t.start()
        if on_started is not None:
            on_started()                 # e.g. start the presolve thread once the stream threads are on their way into the library
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
        for st in self.streams[:len(parts)]:
            main.wait_stream(st)
        out = dict(out)
        out['ticks'] = max(i[0] for i in info if i)
        out['launches'] = sum(i[1] for i in info if i)
        out['spans_ms'] = spans                  # (entered, left) the library per stream thread, relative to the call
        return out


KERNEL_CLASSES = ('cell_trial', 'inst_decide', 'cell_eval', 'inst_step', 'misc', 'cell_step', 'inst_alpha', 'inst_kkt')


def set_profiling(handle, on):
    _check(lib().mseetc_set_profiling(handle._h, int(bool(on))), 'mseetc_set_profiling')


def last_profile(handle):
    "Per-kernel-class (ms, launches, cells, bytes_per_cell) of the last solve on this handle."
    ms = (ctypes.c_double * 8)()
    la = (ctypes.c_int32 * 8)()
    ce = (ctypes.c_int64 * 8)()
    _check(lib().mseetc_last_profile(handle._h, ms, la, ce), 'mseetc_last_profile')
    return {name: dict(ms=ms[i], launches=la[i], cells=ce[i], bytes_per_cell=lib().mseetc_bytes_per_cell(handle._h, i))
            for i, name in enumerate(KERNEL_CLASSES)}


def measure_fp64_peak():
    "Sustained DFMA throughput of the current device in GFLOP/s (roofline denominator of the FP64 pipe)."
    torch = _torch_cuda()
    out = ctypes.c_double(0.0)
    _check(lib().mseetc_measure_fp64_peak(ctypes.byref(out), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), 'mseetc_measure_fp64_peak')
    return out.value


def last_timeline(handle, origin=None, max_entries=4096):
    "Launches of the last solve on `handle` as rows (class index, start ms, end ms), relative to the first launch on `origin`."
    buf = np.zeros((max_entries, 3))
    n = lib().mseetc_last_timeline(handle._h, (origin or handle)._h, buf.ctypes.data_as(ctypes.c_void_p), int(max_entries))
    if n < 0:
        _check(n, 'mseetc_last_timeline')
    return buf[:n]


def eval_interval(inp, num_steps, num_approx, tableau=None):
    """inp: numpy [7, n] planes (b0, F, ds, c0, sr0, sr1, sr2) -> numpy [12, n] (tau jets, b1 jets).
    tableau: None = explicit RK4 steps; dict(A, w, maxIter) = collocation steps (mseetc.train.collocationTableau)."""
    torch = _torch_cuda()
    inp = np.ascontiguousarray(inp, dtype=np.float64)
    n = inp.shape[1]
    d_in = torch.from_numpy(inp).cuda()
    d_out = torch.empty((12, n), dtype=torch.float64, device=d_in.device)
    stream = torch.cuda.current_stream().cuda_stream
    if tableau is None:
        _check(lib().mseetc_eval_interval(n, int(num_steps), int(num_approx), _ptr(d_in), _ptr(d_out), ctypes.c_void_p(stream)),
               'mseetc_eval_interval')
    else:
        A = np.ascontiguousarray(tableau['A'], dtype=np.float64); w = np.ascontiguousarray(tableau['w'], dtype=np.float64)
        _check(lib().mseetc_eval_interval_irk(n, int(num_steps), int(num_approx), int(len(w)), A.ctypes.data_as(ctypes.c_void_p),
                                              w.ctypes.data_as(ctypes.c_void_p), int(tableau.get('maxIter', 10)), _ptr(d_in), _ptr(d_out),
                                              ctypes.c_void_p(stream)), 'mseetc_eval_interval_irk')
    return d_out.cpu().numpy()
