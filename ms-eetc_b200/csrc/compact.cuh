// Compaction of a running batch.
//
// The kernels work on warps of 32 instances at one interval; a warp costs the same whether 3 or 32 of its instances are still
// iterating, and the instances of a batch finish at different iterations (28 .. 42 on the trip-time sweep).  When the finished
// instances are scattered over the slots, the last third of a solve runs on warps that carry mostly finished instances.  Every few
// ticks the plan kernel looks at the occupancy; if the running instances are spread over at least twice the tiles they need,
// those that sit beyond the first A slots (A = number of running instances) are moved into the slots of finished instances among
// the first A, after the results of those have been written out: the running instances fill whole warps again and the tiles
// behind them exit at once.  (Measured, profiles/probe_tail.py: a SORTED sweep finishes tile by tile -- its kernels shrink with
// the running set, 222 -> 113 -> 31 us -- and compacting it costs more than it saves, 20.5 against 20.0 ms: hence the threshold.)  Only the live planes move (track 3, current iterate 34, step 17, stage
// QP 52, value functions 9 = references of the parallel-in-time sweeps) plus the per-instance state; an instance computes the same
// numbers in any slot, so the results do not depend on when (or whether) the batch is compacted.
//   k_compact_plan    one block: A, the holes among the first A slots, the running instances beyond; nothing to do unless the running
//                     instances occupy noticeably more tiles than they need
//   k_cell_extract    (io.cuh, finalPass = false) results of the finished instances
//   k_compact_move    thread = (move, interval): copies the live planes
//   k_compact_finish  thread = move: per-instance state, frees the source slot
#pragma once
#include "io.cuh"

namespace mseetc {

enum { PLAN_MOVES = 0, PLAN_ACTIVE = 1, PLAN_HDR = 8 };

#if defined(__CUDACC__)
// exclusive block-wide prefix sum of one int per thread (1024 threads), total returned to everybody
__device__ inline int block_scan_1024(int v, int* total, int* sm) {
    const int t = threadIdx.x;
    sm[t] = v;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {
        const int x = (t >= d) ? sm[t - d] : 0;
        __syncthreads();
        sm[t] += x;
        __syncthreads();
    }
    const int incl = sm[t];
    *total = sm[1023];
    __syncthreads();
    return incl - v;
}

__global__ void __launch_bounds__(1024) k_compact_plan(Ctx c) {
    __shared__ int sm[1024];
    const int n = c.cfg.nInst, t = threadIdx.x;
    const int seg = (n + 1023) / 1024;
    const int lo = min(n, t * seg), hi = min(n, lo + seg);
    int* plan = c.plan;
    // running instances, and the tiles they touch
    int act = 0;
    for (int s = lo; s < hi; ++s) act += (c.I(SI_PHASE, s) != PH_DONE);
    int A;
    const int before = block_scan_1024(act, &A, sm);
    (void)before;
    int tiles = 0;
    for (int tile = t; tile * 32 < n; tile += 1024) {
        bool any = false;
        for (int s = tile * 32; s < min(n, tile * 32 + 32); ++s) any |= (c.I(SI_PHASE, s) != PH_DONE);
        tiles += any;
    }
    int U;
    block_scan_1024(tiles, &U, sm);
    const int ideal = (A + 31) / 32;
    // worth it when the running instances are spread over at least twice the tiles they need (a sorted trip-time sweep finishes
    // tile by tile and never gets there: its kernels shrink with the running set anyway; a batch in random order does)
    const bool go = A > 0 && U >= 2 * ideal && (U - ideal) >= 4;
    // holes among the first A slots / running instances beyond them: the i-th of the one takes the i-th of the other
    int holes = 0, movers = 0;
    for (int s = lo; s < hi; ++s) {
        const bool done = c.I(SI_PHASE, s) == PH_DONE;
        holes += (s < A && done);
        movers += (s >= A && !done);
    }
    int H, M;
    int hb = block_scan_1024(holes, &H, sm);
    int mb = block_scan_1024(movers, &M, sm);
    if (go) {
        for (int s = lo; s < hi; ++s) {
            const bool done = c.I(SI_PHASE, s) == PH_DONE;
            if (s < A && done) plan[PLAN_HDR + c.cfg.S + hb++] = s;          // destinations
            if (s >= A && !done) plan[PLAN_HDR + mb++] = s;                  // sources
        }
    }
    if (t == 0) { plan[PLAN_MOVES] = go ? M : 0; plan[PLAN_ACTIVE] = A; }      // H == M by construction
}

__global__ void __launch_bounds__(128) k_compact_extract(Ctx c, BatchIO io) {
    if (c.plan[PLAN_MOVES] == 0) return;
    const size_t total = (size_t)c.cfg.NK * c.cfg.S;
    const size_t stride = (size_t)gridDim.x * 128;
    for (size_t idx = (size_t)blockIdx.x * 128 + threadIdx.x; idx < total; idx += stride)
        cell_extract(c, io, (int)(idx / c.cfg.S), (int)(idx % c.cfg.S), false);
}

__global__ void __launch_bounds__(128) k_compact_move(Ctx c) {
    const int m = c.plan[PLAN_MOVES];
    const long total = (long)m * c.cfg.NK;
    for (long idx = (long)blockIdx.x * 128 + threadIdx.x; idx < total; idx += (long)gridDim.x * 128) {
        const int i = (int)(idx / c.cfg.NK), k = (int)(idx % c.cfg.NK);
        const int src = c.plan[PLAN_HDR + i], dst = c.plan[PLAN_HDR + c.cfg.S + i];
        if (k > c.I(SI_N_INT, src)) continue;
        const double* a = &c.W(0, k, src);
        double* b = &c.W(0, k, dst);
        const int it = c.I(SI_PARITY, src) ? WS_IT1 : WS_IT0;
#pragma unroll
        for (int f = 0; f < TRK_N; ++f) b[(WS_TRK + f) * 32] = a[(WS_TRK + f) * 32];
#pragma unroll 8
        for (int f = 0; f < IT_N; ++f) b[(it + f) * 32] = a[(it + f) * 32];
#pragma unroll 8
        for (int f = 0; f < ST_N; ++f) b[(WS_ST + f) * 32] = a[(WS_ST + f) * 32];        // an instance in its line search still needs its step
#pragma unroll 8
        for (int f = 0; f < QP_N; ++f) b[(WS_QP + f) * 32] = a[(WS_QP + f) * 32];
#pragma unroll
        for (int f = RIC_P; f < RIC_N; ++f) b[(WS_RIC + f) * 32] = a[(WS_RIC + f) * 32];
    }
}

__global__ void __launch_bounds__(128) k_compact_finish(Ctx c) {
    const int m = c.plan[PLAN_MOVES];
    // finished instances whose results are out (all of them when a compaction runs)
    for (int s = blockIdx.x * 128 + threadIdx.x; s < c.cfg.nInst && m > 0; s += gridDim.x * 128)
        if (c.I(SI_PHASE, s) == PH_DONE) c.I(SI_EXTRACTED, s) = 1;
    // (the loop above only touches finished slots, the one below only running sources and finished destinations whose flag it
    // overwrites with the source's 0: a destination may be marked by another thread first, so the moves wait for the marks)
    __threadfence();
}

__global__ void __launch_bounds__(128) k_compact_state(Ctx c) {
    const int m = c.plan[PLAN_MOVES];
    if (blockIdx.x == 0 && threadIdx.x == 0) c.done[54] += m;      // instances moved in this solve (statistics)
    for (int i = blockIdx.x * 128 + threadIdx.x; i < m; i += gridDim.x * 128) {
        const int src = c.plan[PLAN_HDR + i], dst = c.plan[PLAN_HDR + c.cfg.S + i];
        for (int f = 0; f < PAR_N; ++f) c.P(f, dst) = c.P(f, src);
        for (int f = 0; f < SD_N; ++f) c.D(f, dst) = c.D(f, src);
        for (int f = 0; f < SI_N; ++f) c.I(f, dst) = c.I(f, src);
        c.I(SI_PHASE, src) = PH_DONE;          // the source slot is free (not a finished instance: the completion counter stays)
        c.I(SI_EXTRACTED, src) = 1;
    }
}
#endif

}  // namespace mseetc
