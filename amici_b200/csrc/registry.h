// registry.h — what the per-model translation units (tu/<model>.cu) hand to the C ABI (engine.cu).
#ifndef AMICI_B200_REGISTRY_H
#define AMICI_B200_REGISTRY_H
#include <cuda_runtime.h>

#include <atomic>
#include <string>

#include "../../include/amici_b200.h"

namespace amici_b200 {

extern std::atomic<long> g_launches;  // kernels launched by this library
std::string& last_error();            // thread-local message behind ab_last_error()
std::string& last_kernel();           // description of the solve kernel this thread launched last (ab_last_kernel_name)

inline bool cuda_ok(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return true;
    last_error() = std::string(what) + ": " + cudaGetErrorString(e);
    return false;
}
#define AB_CUDA(x)                                   \
    do {                                             \
        if (!cuda_ok((x), #x)) return AB_E_CUDA;     \
    } while (0)

// Optional completion tracking of a launch (see bdf_solve_kernel): done[c] counts the finished conditions of sub-batch c
// (device memory, zeroed by the caller), host_flags[c] is raised in mapped pinned host memory when sub-batch c is complete.
struct Progress {
    int* done = nullptr;
    int* host_flags = nullptr;   // device-side address of the mapped host array
    int chunk = 0;
    // one-thread-per-condition kernels (G == 1, dae_kernel): lanes of a warp that take conditions.  A small batch is spread over
    // more warps than N / 32 (set by the launch): less divergence inside a warp, all SMs busy
    int lanes = 32;
};

// lanes per warp for a one-thread-per-condition launch of N conditions on `slots` resident warps: the smallest power of two
// that still fits the batch into one wave
inline int lanes_for_batch(long N, long slots) {
    int lanes = 1;
    while (lanes < 32 && (N + lanes - 1) / lanes > slots) lanes *= 2;
    return lanes;
}

struct ModelEntry {
    const char* name;
    ab_dims dims;
    void (*nominal)(double*, double*);
    int (*launch)(const ab_solver_options&, const ab_batch&, int*, const Progress&, cudaStream_t);
};

// SensitivityMethod::adjoint (adjoint.cuh), dispatched by model name; defined in tu/adjoint.cu
int launch_adjoint(const char* model, const ab_solver_options&, const ab_batch&, int* queue_head, const Progress&, cudaStream_t);

// defined in tu/<model>.cu
ModelEntry entry_jakstat_adjoint();
ModelEntry entry_steadystate();
ModelEntry entry_robertson_ode();
ModelEntry entry_events();
ModelEntry entry_neuron();
ModelEntry entry_massaction();
ModelEntry entry_massaction_small();
// closed-form event models of the reference's python/tests/test_events.py (known-answer tests)
ModelEntry entry_spline_modes();   // general HermiteSpline configurations
ModelEntry entry_ev_heavisides();
ModelEntry entry_ev_nested();
ModelEntry entry_ev_pw_simple();
ModelEntry entry_ev_ddeltax();
ModelEntry entry_ev_pw_semi();
ModelEntry entry_ev_pw_state();
// emitted from the reference's symbolic model object through codegen/de_export.py (no ModelSpec in codegen/models.py)
ModelEntry entry_dirac();
ModelEntry entry_nested_events();
// index-1 DAE model (dae.cuh): the shipped model_robertson_py, emitted from the reference's symbolic model object
ModelEntry entry_robertson();
ModelEntry entry_calvetti();

}  // namespace amici_b200
#endif
