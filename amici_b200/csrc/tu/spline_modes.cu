// instantiates the solve kernel for model 'spline_modes': two forcing splines outside the default HermiteSpline configuration
// (general path of spline.cuh; tests/test_cpu.py, tests/test_gpu.py).
#define AB_FULL 1   // state non-negativity, wall-clock guard (model_common.cuh)
#include "../launch.cuh"
#include "../models/spline_modes.cuh"
namespace amici_b200 {
template <>
int launch_model_preeq<models::spline_modes>(const ab_solver_options& o, const ab_batch& b, int* queue_head, const Progress& pg, cudaStream_t st) {
    return launch_model_impl<models::spline_modes, true>(o, b, queue_head, pg, st);
}
ModelEntry entry_spline_modes() { return make_entry<models::spline_modes>(); } }
