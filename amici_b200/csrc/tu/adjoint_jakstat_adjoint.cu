// instantiates the adjoint kernel (adjoint.cuh) for model 'jakstat_adjoint'.
// the pieces of the scalar solver inlined (measured on JAK-STAT: 31.1k -> 34.8k solves/s, results bit-identical)
#define AB_NI __host__ __device__ __forceinline__
#include "../adjoint_launch.cuh"
#include "../models/jakstat_adjoint.cuh"
namespace amici_b200 {
int launch_adjoint_jakstat_adjoint(const ab_solver_options& o, const ab_batch& b, int* queue_head, const Progress& pg, cudaStream_t st) {
    return launch_adjoint_model<models::jakstat_adjoint>(o, b, queue_head, pg, st);
}
}  // namespace amici_b200
