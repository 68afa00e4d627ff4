// HBM data layout of the batched solver.
//
// All per-(instance, interval) data live in one FP64 workspace, tiled by 32 instances:
//     ws[tile = slot/32][k][field][lane = slot%32]      k = interval / node index (0..NK-1)
// A warp that works on 32 consecutive instances at the same k touches one contiguous 256-byte segment per
// field -- both in the interval-parallel kernels (thread = (k, slot)) and in the instance-parallel Riccati
// sweeps (thread = slot, loop over k) -- and all fields of one (tile, k) are contiguous, so a field access is
// `cell base + compile-time offset` (no per-access address arithmetic) and the per-interval data of a sweep is
// one contiguous block.  Per-instance scalars and parameters are planes of length S.
#pragma once
#include <stdint.h>
#include "jet.cuh"
#include "lossmap.cuh"

namespace mseetc {

// ---- inequality rows of one interval (ocp.py:183-229), in the reference's order
enum Row { R_P0 = 0, R_P1, R_ACC, R_LTR, R_LRG, NROW };
// ---- bound multipliers of one interval / node
enum Zi {
    Z_FEL_L = 0, Z_FEL_U, Z_FPB_L, Z_FPB_U, Z_SL_L, Z_T_L, Z_T_U, Z_B_L, Z_B_U,
    Z_P0_L, Z_P0_U, Z_P1_L, Z_P1_U, Z_ACC_L, Z_ACC_U, Z_LTR_L, Z_LRG_L, NZ
};

// ---- iterate (two buffers: current / trial, selected per instance by a parity bit)
enum ItField {
    IT_FEL = 0, IT_FPB, IT_SL, IT_T, IT_B,      // primal (ocp.py:166-181,248-249)
    IT_W,                                        // IT_W + Row : slacks of the inequality rows
    IT_YT = IT_W + NROW, IT_YB,                  // multipliers of the coupling rows (ocp.py:204-213)
    IT_YD,                                       // IT_YD + Row : multipliers of the inequality rows
    IT_Z = IT_YD + NROW,                         // IT_Z + Zi
    IT_N = IT_Z + NZ
};
// ---- Newton step
enum StField { ST_FEL = 0, ST_FPB, ST_SL, ST_T, ST_B, ST_W, ST_YT = ST_W + NROW, ST_YB, ST_YD, ST_N = ST_YD + NROW };
// ---- stage QP produced by the interval kernel, consumed by the Riccati sweeps
enum QpField {
    // folded stage Hessian over (t,b,f | Fel,Fpb), structurally non-zero entries only, with the loss-epigraph variable s already
    // eliminated (static condensation in cell_eval: s couples to nothing across intervals and its pivot does not depend on
    // the value function, so the sweep carries two controls instead of three)
    QP_H_TT = 0, QP_H_BB, QP_H_BFEL, QP_H_BFPB, QP_H_FF, QP_H_FFEL, QP_H_FELFEL, QP_H_FELFPB, QP_H_FPBFPB,
    QP_TAU_B, QP_TAU_F, QP_PHI_B, QP_PHI_F, QP_RT, QP_RB,          // linearised coupling rows
    QP_G0_B, QP_G0_F, QP_G0_FEL, QP_G0_FPB,                        // condensed gradient, mu-independent part
    QP_G1_T, QP_G1_B, QP_G1_FEL, QP_G1_FPB,                        // condensed gradient, coefficient of mu
    QP_SWEEP_N,                                                     // the backward sweep reads the fields above
    // column of s: Hessian entries (b,s), (Fel,s), (Fpb,s), (s,s) and its gradient parts; needed to recover d s (cell_step),
    // when the regularisation delta_w is non-zero, and by the dense algebra of the last interval
    QP_H_BSL = QP_SWEEP_N, QP_H_FELSL, QP_H_FPBSL, QP_H_SLSL, QP_G0_SL, QP_G1_SL,
    QP_HC_B, QP_HC_FEL, QP_HC_FPB, QP_HC_SL, QP_HPP, QP_GP0, QP_GP1,   // terms in b_{k+1} (for multiplier recovery)
    QP_J_P0_B, QP_J_P0_FEL, QP_J_P1_FEL, QP_J_P1_BN, QP_J_ACC_B,   // inequality row gradients
    QP_J_LTR_FEL, QP_J_LTR_B, QP_J_LTR_BN, QP_J_LRG_FEL, QP_J_LRG_B, QP_J_LRG_BN,
    QP_RES,                                                         // QP_RES + Row : d_j(x) - w_j
    QP_N = QP_RES + NROW
};
// ---- Riccati factors: feedback rows of Fel and Fpb (2 x 3), their constant parts (2), value function P (6, upper triangle), p (3)
enum RicField { RIC_K = 0, RIC_KF = 6, RIC_P = 8, RIC_PV = 14, RIC_N = 17 };
// ---- per-interval partial sums (reduced sequentially per instance: deterministic)
enum PartField {
    PC_TH = 0, PC_F, PC_SLOG, PC_SDAMP,          // evaluated point (the trial point, or the starting point): violation, objective, barrier sums
    PC_DINF, PC_PINF, PC_CMIN, PC_CMAX, PC_ZSUM, PC_YSUM, PC_OWN_B, PC_CN_B, PC_OWN_T,
    PS_AP, PS_AZ, PS_GPHID,                      // step: primal / dual fraction-to-boundary limits, barrier slope
    PART_N
};
enum TrkField { TRK_DS = 0, TRK_C0, TRK_BMAX, TRK_N };

enum WsBase {
    WS_TRK = 0,
    WS_IT0 = WS_TRK + TRK_N,
    WS_IT1 = WS_IT0 + IT_N,
    WS_ST = WS_IT1 + IT_N,
    WS_QP = WS_ST + ST_N,
    WS_RIC = WS_QP + QP_N,
    WS_PART = WS_RIC + RIC_N,
    WS_FIELDS = WS_PART + PART_N
};

// ---- per-instance parameters (specific units, ocp.py:96-116)
enum ParField {
    P_SR0 = 0, P_SR1, P_SR2, P_FEL_L, P_FEL_U, P_FPB_L, P_P_LO, P_P_UP, P_A_LO, P_A_UP, P_CT, P_CR,
    P_BMIN, P_SCALE, P_T, P_T0, P_B0, P_BN, P_MASS,
    P_DYN_AUX, P_DYN_ETAG, P_DYN_FMAX, P_DYN_PMAX, P_DYN_SCALE,   // dynamic loss map (efficiency.py:64-65,101)
    PAR_N
};
// ---- per-instance solver state (doubles)
enum SdField {
    SD_MU = 0, SD_TAU, SD_ALPHA, SD_ALPHA_Z, SD_ALPHA_MIN, SD_THETA, SD_FOBJ, SD_SLOG, SD_SDAMP, SD_GPHID,
    SD_THETA_MIN, SD_THETA_MAX, SD_DELTA_LAST, SD_DELTA, SD_KKT, SD_DINF, SD_PINF, SD_CINF, SD_KKT_BEST,
    SD_FILTER,                                   // SD_FILTER + 2*i : (theta_i, phi_i)
    SD_N = SD_FILTER + 2 * 12
};
enum SiField { SI_PHASE = 0, SI_PARITY, SI_ITERS, SI_STATUS, SI_NLS, SI_NFILT, SI_N_INT, SI_NREG, SI_TICKS, SI_LAST_GAIN,
               SI_NACC,     // consecutive iterations at IPOPT's "acceptable" level
               SI_ORIG,     // index of the instance in the caller's arrays (slots are re-used when the batch is compacted)
               SI_EXTRACTED,   // its results have been written to the caller's arrays
               SI_FACT,     // a factorisation has been stored (its chunk-end value functions are the references of pit.cuh)
               SI_N };

enum Phase { PH_EVAL = 0, PH_TRIAL = 1, PH_DONE = 2, PH_STEPPED = 3, PH_FACTOR = 4 };
enum { RED_W = 16 };   // interleave factor of the per-instance reductions (fixed -> bitwise reproducible sums)
// status codes (mapped to IPOPT's vocabulary by the host shim, ocp.py:362)
enum Status {
    ST_RUNNING = -1, ST_SOLVE_SUCCEEDED = 0, ST_MAXITER = 1, ST_RESTORATION_FAILED = 2, ST_STEP_FAILED = 3,
    ST_INFEASIBLE = 4, ST_INVALID_NUMBER = 5, ST_ACCEPTABLE = 6
};

struct Config {
    int S;            // padded slot count
    int NK;           // max nodes = Nmax + 1
    int nInst;
    int withPn, withPower, energy, lossKind;   // lossKind: 0 none, 1 static, 2 dynamic spline map
    int numSteps, numApprox;
    int maxIter;
    double tol, muInit;
    int stallIters;   // > 0: give up (status MAXITER) after that many iterations without a 10 % gain of the best KKT error
    int initMode;     // 0: initial guess of the reference (ocp.py:325-339); 1: dynamically consistent speed-envelope guess
    int intLosses;    // 1: integrateLosses = True (ocp.py:231-241): epigraph rows on the loss energy integrated over the interval
};

struct IrkTab;     // model.cuh: Butcher tableau of the collocation integrator

struct Ctx {
    Config cfg;
    double* ws;    // WS_FIELDS planes of NK*S
    double* par;   // PAR_N planes of S
    double* sd;    // SD_N planes of S
    int* si;       // SI_N planes of S
    int* done;     // number of finished instances (device counter polled by the host loop)
    unsigned long long* cnt;   // [4] processed cells: trial, eval, riccati backward, riccati forward
    LossMapDev lm;             // spline of the dynamic loss map (lossKind 2), device pointers
    const double* tmin;        // [nInst] or null: minimum trip duration once known (0 = not known yet), see inst_kkt; indexed by SI_ORIG
    const IrkTab* irk;         // null: explicit RK4 (integrationMethod 'RK'); else collocation steps ('IRK', and 'CVODES' by a high-order tableau)
    int* plan;                 // compaction plan (device only; null in the host emulation): [0] moves, [1] active, then sources, destinations

    MS_HD double& W(int field, int k, int slot) const {
        return ws[((size_t)(slot >> 5) * cfg.NK + k) * (WS_FIELDS * 32) + (slot & 31) + field * 32];
    }
    MS_HD double& P(int field, int slot) const { return par[(size_t)field * cfg.S + slot]; }
    MS_HD double& D(int field, int slot) const { return sd[(size_t)field * cfg.S + slot]; }
    MS_HD int& I(int field, int slot) const { return si[(size_t)field * cfg.S + slot]; }
};

}  // namespace mseetc
