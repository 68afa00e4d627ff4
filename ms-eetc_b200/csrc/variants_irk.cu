// Translation unit of libmseetc_b200.so: interval evaluation with the collocation integrator (integrationMethod 'IRK' / 'CVODES';
// see variants.h) -- separate instantiations, so that the Newton iterations and their local arrays stay out of the explicit-RK kernels.
#include "variants.h"

namespace mseetc {
namespace {
MS_CELL_KERNEL(k_cell_trial_eval_irk, 1, (cell_eval<false, true, true>(c, k, s)))
MS_CELL_KERNEL(k_cell_trial_eval_dyn_irk, 1, (cell_eval<true, true, true>(c, k, s)))
MS_CELL_KERNEL(k_cell_eval_irk, 1, (cell_eval<false, false, true>(c, k, s)))
MS_CELL_KERNEL(k_cell_eval_dyn_irk, 1, (cell_eval<true, false, true>(c, k, s)))

__global__ void k_eval_interval_irk(int n, int numSteps, int numApprox, const double* in, double* out, const IrkTab* irk) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    eval_interval_point<true>(i, n, numSteps, numApprox, in, out, irk);
}
}  // namespace

void launch_variant_irk(int which, unsigned grid, cudaStream_t st, const Ctx& c, const BatchIO& io) {
    switch (which) {
        case VK_EVAL_IRK: k_cell_eval_irk<<<grid, 128, 0, st>>>(c, io); break;
        case VK_EVAL_DYN_IRK: k_cell_eval_dyn_irk<<<grid, 128, 0, st>>>(c, io); break;
        case VK_TRIAL_IRK: k_cell_trial_eval_irk<<<grid, 128, 0, st>>>(c, io); break;
        case VK_TRIAL_DYN_IRK: k_cell_trial_eval_dyn_irk<<<grid, 128, 0, st>>>(c, io); break;
        default: break;
    }
}

void launch_eval_interval_irk(int n, int numSteps, int numApprox, const double* in, double* out, const IrkTab* irk, cudaStream_t st) {
    k_eval_interval_irk<<<(n + 127) / 128, 128, 0, st>>>(n, numSteps, numApprox, in, out, irk);
}

}  // namespace mseetc
