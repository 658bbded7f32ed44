// GENERATED: instantiates the solve kernel for model 'jakstat_adjoint' (one translation unit per model keeps builds parallel).
#include "../launch.cuh"
#include "../models/jakstat_adjoint.cuh"
namespace amici_b200 {
template <>
int launch_model_preeq<models::jakstat_adjoint>(const ab_solver_options&, const ab_batch&, int*, const Progress&, cudaStream_t);
ModelEntry entry_jakstat_adjoint() { return make_entry<models::jakstat_adjoint>(); } }
