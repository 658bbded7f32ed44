// GENERATED: instantiates the solve kernel for model 'nested_events' (one translation unit per model keeps builds parallel).
// Event models: IEEE divisions and double pow in the step control.  The approximations used by the event-free kernels
// (model_common.cuh fast_div, bdf_core.cuh nth_root) do not change the error against a tight-tolerance run, but they
// double the deviation from the reference's own KLU run on these round-off-sensitive models (profiles/r02_event_spread.md).
#define AB_EXACT_DIV 1
#define AB_EXACT_ROOT 1
#define AB_FULL 1   // state non-negativity, wall-clock guard (model_common.cuh)
#include "../launch.cuh"
#include "../models/nested_events.cuh"
namespace amici_b200 {
ModelEntry entry_nested_events() { return make_entry<models::nested_events>(); } }
