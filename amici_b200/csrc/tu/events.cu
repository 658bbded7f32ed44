// GENERATED: instantiates the solve kernel for model 'events' (one translation unit per model keeps builds parallel).
#include "../launch.cuh"
#include "../models/events.cuh"
namespace amici_b200 {
// pre-equilibration of models with events is not part of this build
template <>
int launch_model_preeq<models::events>(const ab_solver_options&, const ab_batch&, int*, const Progress&, cudaStream_t) {
    last_error() = "pre-/post-equilibration and presimulation of models with events are not implemented";
    return AB_E_UNSUPPORTED;
}
ModelEntry entry_events() { return make_entry<models::events>(); } }
