// GENERATED: instantiates the solve kernel for model 'steadystate' (one translation unit per model keeps builds parallel).
#include "../launch.cuh"
#include "../models/steadystate.cuh"
namespace amici_b200 { ModelEntry entry_steadystate() { return make_entry<models::steadystate>(); } }
