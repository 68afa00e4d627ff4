// Translation unit(s) of libmseetc_b200.so: integrateLosses = True (ocp.py:231-241) -- loss energies integrated in the time domain
// inside the interval evaluation (see variants.h, loss_energy_rows in core.cuh).  Compiled twice (MS_PART = 0 / 1): the two
// evaluation kernels take as long to compile as everything else together.
#include "variants.h"

#ifndef MS_PART
#define MS_PART 0
#endif

namespace mseetc {
void launch_variant_intl_eval(int which, unsigned grid, cudaStream_t st, const Ctx& c, const BatchIO& io);

#if MS_PART == 0
namespace {
MS_CELL_KERNEL(k_cell_init_intl, 1, (cell_init<true, true>(c, k, s)))
MS_CELL_KERNEL(k_cell_step_intl, 2, (cell_step<true, true>(c, k, s)))
MS_CELL_KERNEL(k_cell_lam_intl, 1, cell_fix_time_multiplier_intl<false>(c, io, k, s))
}  // namespace

void launch_variant_intl(int which, unsigned grid, cudaStream_t st, const Ctx& c, const BatchIO& io) {
    switch (which) {
        case VK_INIT_INTL: k_cell_init_intl<<<grid, 128, 0, st>>>(c, io); break;
        case VK_STEP_INTL: k_cell_step_intl<<<grid, 128, 0, st>>>(c, io); break;
        case VK_LAM_INTL: k_cell_lam_intl<<<grid, 128, 0, st>>>(c, io); break;
        default: launch_variant_intl_eval(which, grid, st, c, io); break;
    }
}
#else
namespace {
MS_CELL_KERNEL(k_cell_trial_eval_intl, 1, (cell_eval<true, true, false, true>(c, k, s)))
MS_CELL_KERNEL(k_cell_eval_intl, 1, (cell_eval<true, false, false, true>(c, k, s)))
}  // namespace

void launch_variant_intl_eval(int which, unsigned grid, cudaStream_t st, const Ctx& c, const BatchIO& io) {
    switch (which) {
        case VK_EVAL_INTL: k_cell_eval_intl<<<grid, 128, 0, st>>>(c, io); break;
        case VK_TRIAL_INTL: k_cell_trial_eval_intl<<<grid, 128, 0, st>>>(c, io); break;
        default: break;
    }
}
#endif

}  // namespace mseetc
