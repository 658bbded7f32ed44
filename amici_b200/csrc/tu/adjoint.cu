// SensitivityMethod::adjoint (adjoint.cuh): dispatch by model name to the per-model translation units tu/adjoint_<model>.cu.
#include <cstring>

#include "../registry.h"

namespace amici_b200 {
int launch_adjoint_jakstat_adjoint(const ab_solver_options&, const ab_batch&, int*, const Progress&, cudaStream_t);
int launch_adjoint_steadystate(const ab_solver_options&, const ab_batch&, int*, const Progress&, cudaStream_t);
int launch_adjoint_robertson_ode(const ab_solver_options&, const ab_batch&, int*, const Progress&, cudaStream_t);
int launch_adjoint_spline_modes(const ab_solver_options&, const ab_batch&, int*, const Progress&, cudaStream_t);

int launch_adjoint(const char* model, const ab_solver_options& o, const ab_batch& b, int* queue_head, const Progress& pg, cudaStream_t st) {
    if (!std::strcmp(model, "jakstat_adjoint")) return launch_adjoint_jakstat_adjoint(o, b, queue_head, pg, st);
    if (!std::strcmp(model, "steadystate")) return launch_adjoint_steadystate(o, b, queue_head, pg, st);
    if (!std::strcmp(model, "robertson_ode")) return launch_adjoint_robertson_ode(o, b, queue_head, pg, st);
    if (!std::strcmp(model, "spline_modes")) return launch_adjoint_spline_modes(o, b, queue_head, pg, st);
    last_error() = "adjoint sensitivity analysis is implemented for the event-free lane-group models only";
    return AB_E_UNSUPPORTED;
}

}  // namespace amici_b200
