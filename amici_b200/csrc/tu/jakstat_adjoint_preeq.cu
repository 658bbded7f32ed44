// GENERATED: instantiates the solve kernel with the pre-equilibration phase for model 'jakstat_adjoint'.
#define AB_FULL 1   // state non-negativity, wall-clock guard (model_common.cuh)
#include "../launch.cuh"
#include "../models/jakstat_adjoint.cuh"
namespace amici_b200 {
template <>
int launch_model_preeq<models::jakstat_adjoint>(const ab_solver_options& o, const ab_batch& b, int* queue_head, const Progress& pg, cudaStream_t st) {
    return launch_model_impl<models::jakstat_adjoint, true>(o, b, queue_head, pg, st);
}
}  // namespace amici_b200
