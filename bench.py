#!/usr/bin/env python
"""Benchmark of the hot path: batched OCP solves per second (BASELINE.json metric).

Workload (BASELINE.json configs[1]): trip-time sweep, 4096 VIRM6 instances on tracks/CH_StGallen_Wil.json,
numIntervals 300 (simulations/config.json), terminal times T_k = Tmin*(0.8 + 0.4 k/4095).  Instances below the
minimum trip time are infeasible and must be flagged as such; the minimum time itself is
computed inside every timed step by a time-optimal solve.

One "step" = one pass of the hot path over the whole batch.  `value` is measured with the inputs already in
HBM (CUDA events on the launch stream); `e2e` goes through the public API (`casadiSolver.solve_batch`) with
host arrays, host->device and device->host copies inside the timed region.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
For N > 1 the driver launches this file under torchrun; every rank solves its own 4096-instance sweep
(weak scaling, no collective on the solve path) and rank 0 prints the line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, 'ms-eetc_b200')
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

N_INST = 4096
N_INT = 300
OPTS = {'numIntervals': N_INT, 'maxIterations': 500, 'integrationMethod': 'RK',
        'integrationOptions': {'order': 4, 'numSteps': 1, 'numApproxSteps': 1}}
TMIN_REF = 1035.5536      # restatement-derived; re-derived on the device in every step
METRIC = 'OCP solves/sec (batched, 1/2/4/8 B200) vs CasADi+IPOPT on host cores'


def sweep_times(n, tmin=TMIN_REF):
    return tmin * (0.8 + 0.4 * np.arange(n) / (n - 1))


# ----------------------------------------------------------------------------------------------------------
# CPU arm: the oracle (numpy/scipy restatement of the reference NLP + IPOPT-style solver) in a process pool
# ----------------------------------------------------------------------------------------------------------
_ORACLE = {}


def _oracle_one(T):
    if 'nlp' not in _ORACLE:
        from oracle.problem import load_train, load_track, discretization_points
        from oracle.nlp import ReferenceNLP
        train = load_train(os.path.join(PKG, 'trains', 'NL_Intercity_VIRM6.json'))
        track = load_track(os.path.join(PKG, 'tracks', 'CH_StGallen_Wil.json'))
        pos, g, v, c = discretization_points(track, N_INT)
        _ORACLE['nlp'] = ReferenceNLP(train, pos, g, v, c, track.length,
                                      dict(numSteps=1, numApproxSteps=1, energyOptimal=True, minimumVelocity=1))
    from oracle import ipm
    nlp = _ORACLE['nlp']
    lbz, ubz, lbg, ubg = nlp.bounds(float(T))
    r = ipm.solve(nlp, nlp.x0(float(T)), lbz, ubz, lbg, ubg, max_iter=500)
    return r.success, r.iters


def cpu_pool_rate(n_samples, cores, repeats=1):
    "Solves/s of the oracle on `cores` worker processes over n_samples feasible instances of the sweep."
    import multiprocessing as mp
    Ts = TMIN_REF * np.linspace(1.001, 1.2, n_samples)
    ctx = mp.get_context('spawn')
    with ctx.Pool(cores) as pool:
        pool.map(_oracle_one, Ts[:cores])               # warm-up: imports + sympy code generation in every worker
        best = None
        for _ in range(repeats):
            t0 = time.perf_counter()
            res = pool.map(_oracle_one, Ts, chunksize=1)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
    ok = sum(1 for s, _ in res if s)
    return n_samples / best, ok, best


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    workers = max(1, min(cores, 32))
    n_samples = 2 * workers
    import multiprocessing as mp
    Ts = TMIN_REF * np.linspace(1.001, 1.2, n_samples)
    ctx = mp.get_context('spawn')
    with ctx.Pool(workers) as pool:
        for _ in range(max(1, min(args.warmup, 1))):
            pool.map(_oracle_one, Ts[:workers])
        t0 = time.perf_counter()
        for _ in range(args.steps):
            res = pool.map(_oracle_one, Ts, chunksize=1)
        dt = time.perf_counter() - t0
    value = n_samples * args.steps / dt
    sample = '%d feasible instances of the sweep (T in Tmin*[1.001,1.2]) per step on %d worker processes' % (n_samples, workers)
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'solves/s', 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': 1e3 * dt / args.steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': 'trip-time sweep, VIRM6 on CH_StGallen_Wil, numIntervals 300 (bounded sample)',
                       'instances_per_step': n_samples},
            'cpu_baseline': {'value': value, 'unit': 'solves/s', 'cores': workers, 'kind': 'port', 'sample': sample,
                             'note': 'oracle port (numpy/scipy restatement of the reference NLP + IPOPT-style filter IP), '
                                     'NOT CasADi+IPOPT: casadi is not installable in this image'},
            'e2e': {'value': value, 'unit': 'solves/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'converged': int(sum(1 for s, _ in res if s))}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu), '--query-gpu=' + self.QUERY, '--format=csv,noheader,nounits',
                                          '-lms', '100'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[5:9]):
                    if val.lower().startswith('active'):
                        reasons.add(name)
            except Exception:
                pass
        busy = [x for x in sm if x > 0.5 * max(sm)] if sm else []
        return {'sm_mhz': float(np.median(busy)) if busy else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


def run_gpu(args):
    import torch
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py needs a CUDA device (there is no CPU fallback for the product arm)')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)
    import __graft_entry__ as ge
    if rank == 0:
        ge.build()
    if dist is not None:
        dist.barrier()
    ge.build()
    from mseetc.ocp import casadiSolver
    from mseetc.train import Train
    from mseetc.track import Track
    from mseetc import _cabi

    n = args.instances
    train = Train(config={'id': 'NL_Intercity_VIRM6'})
    track = Track(config={'id': 'CH_StGallen_Wil'})
    solver = casadiSolver(train, track, OPTS)
    tsolver = solver._time_sibling()
    T = sweep_times(n)

    # ---------------- device-resident inputs (uploaded once, outside the timed region)
    zero = np.zeros(n)
    P, M = solver._planes(n, T, zero, zero + 1.0, zero + 1.0, {}, (1 - train.etaTraction) / train.etaTraction, 1 - train.etaRgBrake)
    lim = np.minimum(solver.points['Speed limit [m/s]'].values[:-1], solver._base['velocityMax'])
    horizon = 1.5 * float(np.sum(solver.steps / lim))      # same bound as casadiSolver.minimum_time
    Pt, _ = tsolver._planes(1, np.array([horizon]), np.zeros(1), np.ones(1), np.ones(1), {}, 0.0, 0.0)
    ds, c0, bmax = solver._tables(solver._base['rho'], solver._base['g'], solver._base['velocityMax'])
    # several streams: instances are dealt to the sub-batches tile by tile, exactly as casadiSolver.solve_batch does; the resident
    # inputs are stored in that order and the per-instance results are put back in sweep order for the checks below
    perm, parts = _cabi.StreamPool.interleave(n, max(1, args.streams)) if args.streams > 1 else (np.arange(n), None)
    back = np.argsort(perm)
    P = np.ascontiguousarray(P[:, perm])
    up = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(device=dev, dtype=dt)
    d = dict(P=up(P, torch.float64), Pt=up(Pt, torch.float64), nint=up(np.full(n, N_INT, np.int32), torch.int32),
             nint1=up(np.full(1, N_INT, np.int32), torch.int32), trk_of=up(np.zeros(n, np.int32), torch.int32),
             trk_of1=up(np.zeros(1, np.int32), torch.int32), trk_off=up(np.array([0, N_INT], np.int32), torch.int32),
             ds=up(ds, torch.float64), c0=up(c0, torch.float64), bmax=up(bmax, torch.float64))
    solver.streams = args.streams
    pool = _cabi.StreamPool(solver._make_handle, max(1, args.streams), dev)
    h = pool.handles[0]
    ht = tsolver._ensure_handle()
    for hh in pool.handles + [ht]:
        _cabi.set_profiling(hh, not args.no_profile)
    stp = 3 + h.nu
    outbuf = dict(z=torch.zeros((n, N_INT * stp + 2), dtype=torch.float64, device=dev), lam=None,
                  obj=torch.empty(n, dtype=torch.float64, device=dev), kkt=torch.empty(n, dtype=torch.float64, device=dev),
                  iters=torch.empty(n, dtype=torch.int32, device=dev), status=torch.empty(n, dtype=torch.int32, device=dev))
    prof = {k: dict(ms=0.0, launches=0, cells=0) for k in _cabi.KERNEL_CLASSES}          # the sweep (all pool streams)
    prof_pre = {k: dict(ms=0.0, launches=0, cells=0) for k in _cabi.KERNEL_CLASSES}      # the concurrent time-optimal presolve
    launches = [0]

    tmin_dev = torch.zeros(n, dtype=torch.float64, device=dev)
    side = torch.cuda.Stream(device=dev, priority=-5)

    def step(accumulate):
        # (1) minimum trip time of the (single) distinct problem of this sweep: time-optimal solve on a side stream, driven by
        #     a second host thread, concurrently with (2); its result lands in tmin_dev while the batch is iterating
        tmin_dev.zero_()
        box = {}

        def presolve():
            torch.cuda.set_device(dev)
            side.wait_stream(torch.cuda.default_stream(dev))
            with torch.cuda.stream(side):
                tr = ht.solve_device(d['Pt'], d['nint1'], d['trk_of1'], d['trk_off'], d['ds'], d['c0'], d['bmax'])
                tmin_dev.copy_(tr['z'][:, -2].expand(n))          # t_N - t_0 with t_0 = 0
                side.synchronize()
            box['tr'] = tr

        th = threading.Thread(target=presolve)
        th.start()
        # (2) the sweep; instances below the minimum time are flagged infeasible by the library as soon as it is known
        out = pool.solve(d['P'], d['nint'], d['trk_of'], d['trk_off'], d['ds'], d['c0'], d['bmax'], tmin=tmin_dev, out=dict(outbuf), parts=parts)
        th.join()
        tr = box['tr']
        if accumulate:
            launches[0] += tr['launches'] + out['launches']
            for hh in pool.handles + [ht]:
                acc = prof_pre if hh is ht else prof
                for k, v in _cabi.last_profile(hh).items():
                    acc[k]['ms'] += v['ms']; acc[k]['launches'] += v['launches']; acc[k]['cells'] += v['cells']
                    acc[k]['bytes_per_cell'] = v['bytes_per_cell']
        return out, tr

    for _ in range(args.warmup):
        step(False)
    torch.cuda.synchronize(dev)
    if dist is not None:
        dist.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(dev)
    e0.record()
    for _ in range(args.steps):
        out, tr = step(True)
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    status = out['status'].cpu().numpy()[back]
    iters = out['iters'].cpu().numpy()[back]
    kkt = out['kkt'].cpu().numpy()[back]
    tmin_dev = float(tr['z'][0, -2].item())
    feas = T >= tmin_dev
    n_ok = int(np.sum((status == 0) & feas & (kkt <= 1e-8)))
    n_flag = int(np.sum((status != 0) & ~feas))      # flagged by the library or failed on their own before the certificate arrived

    # ---------------- e2e: public API, host arrays in, host arrays out
    for _ in range(min(args.warmup, 3)):
        res = solver.solve_batch(T)
    torch.cuda.synchronize(dev)
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    e2e_steps = []
    for _ in range(args.steps):
        ts = time.perf_counter()
        res = solver.solve_batch(T)
        e2e_steps.append(round(1e3 * (time.perf_counter() - ts), 2))
        if e2e_steps[-1] >= max(e2e_steps):
            slowest = {k: round(1e3 * v, 2) for k, v in res['timing'].items()}
            slowest['stream_threads_ms'] = res.get('spans_ms')
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    clocks = sampler.stop() if rank == 0 else None
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---------------- roofline of the dominant kernel (largest share of the device time)
    peaks_file = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(peaks_file):
        peak, peak_src = float(json.load(open(peaks_file))['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    else:
        peak, peak_src = 6650.0, 'fallback (B200_PROFILING.md)'
    kern = {}
    for k, v in prof.items():
        if v['launches'] == 0 or k == 'misc':
            continue
        gbs = v['cells'] * v['bytes_per_cell'] / (v['ms'] * 1e-3) / 1e9 if v['ms'] > 0 else 0.0
        kern[k] = {'ms_total': v['ms'], 'launches': v['launches'], 'avg_us': 1e3 * v['ms'] / v['launches'],
                   'bytes_per_cell': v['bytes_per_cell'], 'cells': v['cells'], 'achieved_gbs': gbs, 'frac_hbm': gbs / peak}
    if not kern:
        kern = {'inst_step': dict(ms_total=1.0, launches=1, avg_us=0.0, bytes_per_cell=0.0, cells=0, achieved_gbs=0.0, frac_hbm=0.0)}
    top = max(kern, key=lambda k: kern[k]['ms_total'])
    ncu_traffic, capture = None, None
    tfile = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(tfile):
        tj = json.load(open(tfile))
        ncu_traffic = tj.get(top)
        capture = (tj.get('captures_r01_q') or {}).get(top)
    # FP64 view (report both candidates): static FP64 operation counts per cell (profiles/fp64_ops.json, from the SASS
    # of this build) against the DFMA throughput measured on this device
    fp64 = None
    ofile = os.path.join(ROOT, 'profiles', 'fp64_ops.json')
    if os.path.exists(ofile):
        ops = json.load(open(ofile))
        peak64 = _cabi.measure_fp64_peak()
        fp64 = {'peak_gflops': peak64, 'peak_source': 'measured here (k_fp64_peak: independent DFMA chains, CUDA events)', 'kernels': {}}
        for k, v in kern.items():
            if k in ops and v['ms_total'] > 0:
                gf = v['cells'] * ops[k]['flop_per_cell'] / (v['ms_total'] * 1e-3) / 1e9
                fp64['kernels'][k] = {'flop_per_cell': ops[k]['flop_per_cell'], 'achieved_gflops': gf, 'frac_fp64': gf / peak64}
    roofline = {'bound': 'hbm', 'kernel': top, 'achieved': kern[top]['achieved_gbs'], 'peak': peak, 'unit': 'GB/s',
                'frac': kern[top]['achieved_gbs'] / peak, 'traffic': ncu_traffic, 'traffic_capture': capture, 'peak_source': peak_src,
                'share_of_device_time': kern[top]['ms_total'] / max(1e-12, sum(v['ms_total'] for v in kern.values())),
                'bytes_per_launch': kern[top]['cells'] * kern[top]['bytes_per_cell'] / kern[top]['launches'], 'kernels': kern, 'fp64': fp64,
                'presolve': {'ms_total': sum(v['ms'] for v in prof_pre.values()), 'launches': sum(v['launches'] for v in prof_pre.values()),
                             'note': 'single-instance time-optimal solve on its own stream, concurrent with the sweep (not in `kernels`)'}}

    # ---------------- CPU baseline beside it (bounded sample, rank 0, N = 1 only)
    cpu = None
    if world == 1 and not args.no_cpu:
        cores = max(1, min(os.cpu_count() or 1, 16))
        rate, ok, dt = cpu_pool_rate(cores, cores)
        cpu = {'value': rate, 'unit': 'solves/s', 'cores': cores, 'kind': 'port',
               'sample': '%d feasible instances of the same sweep, one per worker process, %.1f s wall, %d converged' % (cores, dt, ok),
               'note': 'oracle port (numpy/scipy restatement + IPOPT-style filter IP), not CasADi+IPOPT'}

    total = n * world
    line = {'metric': METRIC, 'value': total * args.steps / (ms * 1e-3), 'unit': 'solves/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': 'trip-time sweep: %d VIRM6 instances per GPU on CH_StGallen_Wil, numIntervals %d, '
                                   'T = Tmin*[0.8,1.2] (BASELINE configs[1])' % (n, N_INT),
                       'instances_per_gpu': n, 'feasible_per_gpu': int(feas.sum()), 'converged_feasible': n_ok,
                       'flagged_infeasible': n_flag, 'tmin_s': tmin_dev, 'ip_iterations_mean': float(iters[feas].mean()),
                       'ip_iterations_max': int(iters[feas].max()), 'ticks': int(out['ticks']),
                       'streams': args.streams, 'l2': 'per-tick working set %.2f GB >> 126 MB L2' % (_cabi.lib().mseetc_workspace_bytes(h._h, n) / 1e9)},
            'feasible_solves_per_s': int(feas.sum()) * world * args.steps / (ms * 1e-3),
            'e2e': {'value': total * args.steps / e2e_s, 'unit': 'solves/s', 'h2d_bytes_per_step': res['h2d_bytes'],
                    'last_call_breakdown_s': res.get('timing'),
                    'd2h_bytes_per_step': res['d2h_bytes'], 'ms_per_step': 1e3 * e2e_s / args.steps, 'steps_ms': e2e_steps,
                    'slowest_step_breakdown_ms': slowest},
            'gpu_launches': launches[0], 'clocks': clocks, 'roofline': roofline}
    if cpu is not None:
        line['cpu_baseline'] = cpu
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--instances', type=int, default=N_INST)
    ap.add_argument('--no-cpu', action='store_true', help='skip the CPU baseline leg')
    ap.add_argument('--no-profile', action='store_true', help='no per-kernel CUDA events (roofline block is then empty)')
    ap.add_argument('--streams', type=int, default=2, help='concurrent sub-batches (CUDA streams / host threads) per GPU')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == '__main__':
    main()
