// GENERATED: instantiates the solve kernel for model 'robertson_ode' (one translation unit per model keeps builds parallel).
#include "../launch.cuh"
#include "../models/robertson_ode.cuh"
namespace amici_b200 {
template <>
int launch_model_preeq<models::robertson_ode>(const ab_solver_options&, const ab_batch&, int*, const Progress&, cudaStream_t);
ModelEntry entry_robertson_ode() { return make_entry<models::robertson_ode>(); } }
