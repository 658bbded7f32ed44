// GENERATED: instantiates the solve kernel with the pre-equilibration phase for model 'robertson_ode'.
#define AB_FULL 1   // state non-negativity, wall-clock guard (model_common.cuh)
#include "../launch.cuh"
#include "../models/robertson_ode.cuh"
namespace amici_b200 {
template <>
int launch_model_preeq<models::robertson_ode>(const ab_solver_options& o, const ab_batch& b, int* queue_head, const Progress& pg, cudaStream_t st) {
    return launch_model_impl<models::robertson_ode, true>(o, b, queue_head, pg, st);
}
}  // namespace amici_b200
