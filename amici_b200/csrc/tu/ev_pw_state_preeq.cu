// instantiates the solve kernel with the pre-equilibration / presimulation phases for the closed-form event model 'ev_pw_state' (reference python/tests/test_events.py; known-answer tests).
// Event models: IEEE divisions and double pow in the step control (see tu/events.cu).
#define AB_EXACT_DIV 1
#define AB_EXACT_ROOT 1
#define AB_FULL 1   // state non-negativity, wall-clock guard (model_common.cuh)
#include "../launch.cuh"
#include "../models/ev_pw_state.cuh"
namespace amici_b200 {
template <>
int launch_model_preeq<models::ev_pw_state>(const ab_solver_options& o, const ab_batch& b, int* queue_head, const Progress& pg, cudaStream_t st) {
    return launch_model_impl<models::ev_pw_state, true>(o, b, queue_head, pg, st);
}
}  // namespace amici_b200
