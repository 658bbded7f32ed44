// GENERATED: instantiates the solve kernel for model 'jakstat_adjoint' (one translation unit per model keeps builds parallel).
#include "../launch.cuh"
#include "../models/jakstat_adjoint.cuh"
namespace amici_b200 { ModelEntry entry_jakstat_adjoint() { return make_entry<models::jakstat_adjoint>(); } }
