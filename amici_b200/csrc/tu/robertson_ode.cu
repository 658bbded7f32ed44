// GENERATED: instantiates the solve kernel for model 'robertson_ode' (one translation unit per model keeps builds parallel).
#include "../launch.cuh"
#include "../models/robertson_ode.cuh"
namespace amici_b200 { ModelEntry entry_robertson_ode() { return make_entry<models::robertson_ode>(); } }
