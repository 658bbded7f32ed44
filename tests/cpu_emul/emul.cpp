// tests/cpu_emul/emul.cpp — TEST INFRASTRUCTURE, never loaded by the product.
//
// Compiles the solver source (amici_b200/csrc/*.cuh) for the host with a lane-group of one
// lane that owns every column, so the step-control logic can be checked against the compiled
// reference on machines without a GPU.  The product path is the CUDA build of the same source.
#define AB_FULL 1
#include "../../amici_b200/csrc/flat.cuh"
#include "../../amici_b200/csrc/adjoint.cuh"
#include "../../amici_b200/csrc/models/jakstat_adjoint.cuh"
#include "../../amici_b200/csrc/models/robertson_ode.cuh"
#include "../../amici_b200/csrc/models/steadystate.cuh"
#include "../../amici_b200/csrc/models/events.cuh"
#include "../../amici_b200/csrc/models/neuron.cuh"
#include "../../amici_b200/csrc/models/spline_modes.cuh"
#include "../../amici_b200/csrc/models/ev_heavisides.cuh"
#include "../../amici_b200/csrc/models/ev_nested.cuh"
#include "../../amici_b200/csrc/models/ev_pw_simple.cuh"
#include "../../amici_b200/csrc/models/ev_ddeltax.cuh"
#include "../../amici_b200/csrc/models/ev_pw_semi.cuh"
#include "../../amici_b200/csrc/models/ev_pw_state.cuh"
#include "../../amici_b200/csrc/models/dirac.cuh"
#include "../../amici_b200/csrc/models/nested_events.cuh"
#include "../../amici_b200/csrc/models/robertson.cuh"
#include "../../amici_b200/csrc/models/calvetti.cuh"
#include "../../amici_b200/csrc/dae.cuh"

#include <memory>
#include <string>
#include <vector>

using namespace amici_b200;

#if defined(AB_SLDET_RECORD)
// harvest of the ring buffers cvSLdet is called with (tests/golden/make_sldet_cases.py): 43 doubles per call
static double* g_rec = nullptr;
static long g_rec_cap = 0, g_rec_n = 0;
namespace amici_b200 {
void sldet_record(const double* ssdat, int head, int q, int rc) {
    if (!g_rec || g_rec_n >= g_rec_cap) return;
    double* d = g_rec + 43 * g_rec_n++;
    for (int i = 0; i < 40; ++i) d[i] = ssdat[i];
    d[40] = head; d[41] = q; d[42] = rc;
}
}
extern "C" void emul_sldet_record(double* buf, long cap) { g_rec = buf; g_rec_cap = cap; g_rec_n = 0; }
extern "C" long emul_sldet_recorded() { return g_rec_n; }
#endif

template <class M>
static int run_model(const SolverOptions* o, const BatchView* b) {
    constexpr int NC = M::NP + 1;
    using Sim = Simulation<M, 1, NC, true>;
    auto S = std::make_unique<typename Sim::Shared>();
    if (o->sensi >= 1 && b->nplist + 1 > NC) return -1;
    for (int n = 0; n < b->N; ++n) {
        Sim sim(*S, *o, *b, n);
        sim.run();
    }
    return 0;
}

// The flat driver (flat.cuh) with the compact column layout where the model has identically-zero sensitivity columns:
// what the CUDA build launches for event-free models.  Must agree with run_model bit for bit.
template <class M>
static int run_model_flat(const SolverOptions* o, const BatchView* b) {
    constexpr int NC = M::NP_DYN + 1;
    using D = FlatDriver<M, 1, NC>;
    static_assert(M::NP_DYN == M::NP || Layout<M, 1, NC>::COMPACT, "compact layout expected");
    auto S = std::make_unique<typename D::Shared>();
    if (o->sensi >= 1 && b->nplist > M::NP) return -1;
    if (b->sx0_in && M::NP_DYN != M::NP) return -3;
    int head = 0;
    WarpAlign align;
    flat_run<M, 1, NC>(*S, *o, *b, &head, align, [](int) {});
    return 0;
}

// adjoint sensitivity analysis (adjoint.cuh): one condition after the other, history on the heap
template <class M>
static int run_model_adjoint(const SolverOptions* o, const BatchView* b) {
    const int cap = o->maxsteps + 2;
    std::vector<double> ht(cap), hy((size_t)cap * M::NX), dj((size_t)(b->nt > 0 ? b->nt : 1) * M::NX);
    std::vector<int> hq(cap);
    auto hist = std::make_unique<FwdHistory<M::NX>>();
    hist->t = ht.data(); hist->y = hy.data(); hist->ord = hq.data(); hist->stride = 1; hist->cap = cap;
    for (int n = 0; n < b->N; ++n) {
        auto sim = std::make_unique<AdjointSimulation<M>>(*o, *b, *hist, dj.data(), n);
        sim->run();
    }
    return 0;
}

extern "C" int emul_run_adjoint(const char* model, const SolverOptions* o, const BatchView* b) {
    std::string m(model);
    if (m == "jakstat_adjoint") return run_model_adjoint<models::jakstat_adjoint>(o, b);
    if (m == "steadystate") return run_model_adjoint<models::steadystate>(o, b);
    if (m == "robertson_ode") return run_model_adjoint<models::robertson_ode>(o, b);
    return -2;
}

extern "C" int emul_run_flat(const char* model, const SolverOptions* o, const BatchView* b) {
    std::string m(model);
    if (m == "jakstat_adjoint") return run_model_flat<models::jakstat_adjoint>(o, b);
    if (m == "steadystate") return run_model_flat<models::steadystate>(o, b);
    if (m == "robertson_ode") return run_model_flat<models::robertson_ode>(o, b);
    return -2;
}

extern "C" int emul_run(const char* model, const SolverOptions* o, const BatchView* b) {
    if (o->sensi >= 1 && o->sensi_meth == 2 && b->nplist > 0) return emul_run_adjoint(model, o, b);   // as engine.cu launch_solve
    std::string m(model);
    if (m == "jakstat_adjoint") return run_model<models::jakstat_adjoint>(o, b);
    if (m == "steadystate") return run_model<models::steadystate>(o, b);
    if (m == "robertson_ode") return run_model<models::robertson_ode>(o, b);
    if (m == "events") return run_model<models::events>(o, b);
    if (m == "neuron") return run_model<models::neuron>(o, b);
    if (m == "spline_modes") return run_model<models::spline_modes>(o, b);
    if (m == "ev_heavisides") return run_model<models::ev_heavisides>(o, b);
    if (m == "ev_nested") return run_model<models::ev_nested>(o, b);
    if (m == "ev_pw_simple") return run_model<models::ev_pw_simple>(o, b);
    if (m == "ev_ddeltax") return run_model<models::ev_ddeltax>(o, b);
    if (m == "ev_pw_semi") return run_model<models::ev_pw_semi>(o, b);
    if (m == "ev_pw_state") return run_model<models::ev_pw_state>(o, b);
    if (m == "dirac") return run_model<models::dirac>(o, b);
    if (m == "nested_events") return run_model<models::nested_events>(o, b);
    if (m == "robertson" || m == "calvetti") {   // DAE path (dae.cuh): forward simulation only, as launch_dae_model
        if (o->sensi >= 1 && b->nplist > 0 && m != "robertson") return -3;   // as launch_dae_model: FSA of event-free DAE models only
        for (int n = 0; n < b->N; ++n) {
            if (m == "robertson") std::make_unique<DaeSimulation<models::robertson>>(*o, *b, n)->run();
            else std::make_unique<DaeSimulation<models::calvetti>>(*o, *b, n)->run();
        }
        return 0;
    }
    return -2;
}

// host twin of ab_selftest_spline (include/amici_b200.h): the same hermite_spline.cuh functions, compiled for the host
#include "../../amici_b200/csrc/hermite_spline.cuh"
extern "C" int emul_selftest_spline(int ncases, const int* cfg, const double* nodes, const double* values, const double* slopes, int npar,
                                    const double* dvaluesdp, const double* dslopesdp, int nt, const double* ts, double* value_out,
                                    double* sens_out, int* err_out) {
    using namespace amici_b200;
    constexpr int NN = 8, NP = 4;
    if (npar > NP) return -1;
    for (int c = 0; c < ncases; ++c) {
        const int* k = cfg + 8 * c;
        const SplineCfg sc{k[1], k[2], k[3], k[4], k[5] != 0, k[6] != 0, k[7] != 0};
        HermiteSplineData<NN, NP> H;
        hs_build(H, sc, k[0], nodes + (size_t)c * NN, values + (size_t)c * NN, slopes + (size_t)c * NN, npar,
                 dvaluesdp + (size_t)c * NN * npar, dslopesdp + (size_t)c * NN * npar);
        for (int i = 0; i < nt; ++i) {
            bool err = false;
            const double t = ts[(size_t)c * nt + i];
            value_out[(size_t)c * nt + i] = hs_value(H, t, &err);
            for (int ip = 0; ip < npar; ++ip) sens_out[((size_t)c * nt + i) * npar + ip] = hs_sensitivity(H, t, ip, &err);
            err_out[(size_t)c * nt + i] = err ? 1 : 0;
        }
    }
    return 0;
}
