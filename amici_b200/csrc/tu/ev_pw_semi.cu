// instantiates the solve kernel for the closed-form event model 'ev_pw_semi' (reference python/tests/test_events.py; known-answer tests).
// Event models: IEEE divisions and double pow in the step control (see tu/events.cu).
#define AB_EXACT_DIV 1
#define AB_EXACT_ROOT 1
#define AB_FULL 1   // state non-negativity, wall-clock guard (model_common.cuh)
#include "../launch.cuh"
#include "../models/ev_pw_semi.cuh"
namespace amici_b200 {
ModelEntry entry_ev_pw_semi() { return make_entry<models::ev_pw_semi>(); } }
