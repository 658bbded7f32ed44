// instantiates the DAE kernel (dae.cuh) for model 'robertson' (the shipped index-1 DAE form, reference models/model_robertson_py).
#include "../dae_launch.cuh"
#include "../models/robertson.cuh"
namespace amici_b200 {
ModelEntry entry_robertson() { return make_dae_entry<models::robertson>(); } }
