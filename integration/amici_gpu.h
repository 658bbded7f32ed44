// amici_gpu.h — the reference-side binding of libamici_b200.so: what a maintainer adds to libamici next to
// amici::run_simulations (reference include/amici/amici.h:21-39, src/amici.cpp:220-264).
#ifndef AMICI_GPU_H
#define AMICI_GPU_H
#include <amici/amici.h>

#include <memory>
#include <vector>

namespace amici {

/** Same signature, ownership and error convention as amici::run_simulations: inputs borrowed and not mutated, one owning
 * ReturnData per ExpData in list order, never throws for per-condition integration failures (ReturnData::status carries the
 * code, outputs are NaN from the failing timepoint on, one message per failed condition).  `num_threads` is ignored (the batch
 * runs on the calling thread's current CUDA device); `failfast` has the reference's single-thread semantics: the conditions
 * after the first failed one are returned as freshly constructed ReturnData with status AMICI_NOT_RUN.
 * Throws AmiException for what the reference also throws for: inconsistent ExpData dimensions, settings the engine does not
 * implement (adjoint sensitivities, second order), an unknown model. */
std::vector<std::unique_ptr<ReturnData>> run_simulations_gpu(
    Solver const& solver, std::vector<ExpData*> const& edatas, Model const& model, bool failfast, int num_threads);

/** Engine registry name of a reference model: "model_jakstat_adjoint_py" -> "jakstat_adjoint". */
std::string gpu_model_name(Model const& model);

}  // namespace amici
#endif
