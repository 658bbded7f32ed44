// GENERATED: instantiates the solve kernel for model 'steadystate' (one translation unit per model keeps builds parallel).
#include "../launch.cuh"
#include "../models/steadystate.cuh"
namespace amici_b200 {
template <>
int launch_model_preeq<models::steadystate>(const ab_solver_options&, const ab_batch&, int*, const Progress&, cudaStream_t);
ModelEntry entry_steadystate() { return make_entry<models::steadystate>(); } }
