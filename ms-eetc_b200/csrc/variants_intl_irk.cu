// Translation unit(s) of libmseetc_b200.so: integrateLosses = True together with the collocation integrator ('IRK' / 'CVODES'): the
// duration in the loss rows comes from the collocation steps (see variants.h; the step kernel is the one of variants_intl.cu).
// Compiled three times (MS_PART = 0 / 1 / 2), one evaluation kernel per unit.
#include "variants.h"

#ifndef MS_PART
#define MS_PART 0
#endif

namespace mseetc {
void launch_variant_intl_irk_eval(int which, unsigned grid, cudaStream_t st, const Ctx& c, const BatchIO& io);
void launch_variant_intl_irk_trial(int which, unsigned grid, cudaStream_t st, const Ctx& c, const BatchIO& io);

#if MS_PART == 0
namespace {
MS_CELL_KERNEL(k_cell_init_intl_irk, 1, (cell_init<true, true, true>(c, k, s)))
MS_CELL_KERNEL(k_cell_lam_intl_irk, 1, cell_fix_time_multiplier_intl<true>(c, io, k, s))
}  // namespace

void launch_variant_intl_irk(int which, unsigned grid, cudaStream_t st, const Ctx& c, const BatchIO& io) {
    switch (which) {
        case VK_INIT_INTL_IRK: k_cell_init_intl_irk<<<grid, 128, 0, st>>>(c, io); break;
        case VK_LAM_INTL_IRK: k_cell_lam_intl_irk<<<grid, 128, 0, st>>>(c, io); break;
        case VK_EVAL_INTL_IRK: launch_variant_intl_irk_eval(which, grid, st, c, io); break;
        default: launch_variant_intl_irk_trial(which, grid, st, c, io); break;
    }
}
#elif MS_PART == 1
namespace {
MS_CELL_KERNEL(k_cell_eval_intl_irk, 1, (cell_eval<true, false, true, true>(c, k, s)))
}  // namespace
void launch_variant_intl_irk_eval(int, unsigned grid, cudaStream_t st, const Ctx& c, const BatchIO& io) {
    k_cell_eval_intl_irk<<<grid, 128, 0, st>>>(c, io);
}
#else
namespace {
MS_CELL_KERNEL(k_cell_trial_eval_intl_irk, 1, (cell_eval<true, true, true, true>(c, k, s)))
}  // namespace
void launch_variant_intl_irk_trial(int, unsigned grid, cudaStream_t st, const Ctx& c, const BatchIO& io) {
    k_cell_trial_eval_intl_irk<<<grid, 128, 0, st>>>(c, io);
}
#endif

}  // namespace mseetc
