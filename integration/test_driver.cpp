// test_driver.cpp — TEST of the reference-side binding: the SAME std::vector<ExpData*> through amici::run_simulations (the
// unmodified reference on the host cores) and through amici::run_simulations_gpu (integration/amici_gpu.cpp -> C ABI -> B200),
// every ReturnData field of both copied out side by side for the Python test (tests/test_gpu.py::test_reference_side_binding).
// Linked against oracle/_ref/libamici.a + one pre-generated reference model + libamici_b200.so by integration/build.sh.
#include <amici/amici.h>

#include <cmath>
#include <cstring>
#include <random>

#include "amici_gpu.h"
#include "wrapfunctions.h"

using namespace amici;

namespace {
struct Case {
    std::unique_ptr<Model> model;
    std::unique_ptr<Solver> solver;
    std::vector<std::unique_ptr<ExpData>> owned;
    std::vector<ExpData*> edatas;
};

// A deliberately heterogeneous list (ConditionContext in full): own timepoints, plist, pscale with and without parameters,
// fixed parameters, pre-equilibration / presimulation / re-initialisation where the model has fixed parameters, x0 / sx0, a NULL
// entry, and one condition that must fail (maxsteps) when `with_failure` is set.
Case make_case(int n, unsigned seed, int scenario, int linsol, double rtol, double atol, int reporting = 0) {
    Case c;
    c.model = amici::generic_model::get_model();
    c.solver = c.model->create_solver();
    Model& m = *c.model;
    Solver& s = *c.solver;
    s.set_relative_tolerance(rtol);
    s.set_absolute_tolerance(atol);
    if (linsol > 0) s.set_linear_solver(static_cast<LinearSolver>(linsol));
    if (rtol < 1e-10) s.set_max_steps(500000);
    s.set_return_data_reporting_mode(static_cast<RDataReporting>(reporting));
    s.set_sensitivity_order(SensitivityOrder::first);
    s.set_sensitivity_method(SensitivityMethod::forward);
    std::vector<realtype> ts;
    for (int i = 0; i <= 10; ++i) ts.push_back(i * (m.get_name().find("jakstat") != std::string::npos ? 6.0 : 1.0));
    m.set_timepoints(ts);
    std::mt19937 rng(seed);
    std::uniform_real_distribution<double> U(0.0, 1.0);
    auto const p0 = m.get_free_parameters();
    auto const k0 = m.get_fixed_parameters();
    for (int i = 0; i < n; ++i) {
        if (scenario == 1 && i == 3) {   // a NULL ExpData: the model as it is
            c.owned.emplace_back(nullptr);
            c.edatas.push_back(nullptr);
            continue;
        }
        auto e = std::make_unique<ExpData>(m);
        std::vector<realtype> p(p0), k(k0);
        for (auto& v : p) v *= std::pow(10.0, 0.2 * (U(rng) - 0.5));
        for (auto& v : k) v *= 0.8 + 0.45 * U(rng);
        e->free_parameters = p;
        e->fixed_parameters = k;
        std::vector<realtype> my(e->nt() * m.nytrue), sg(e->nt() * m.nytrue, 0.1);
        for (auto& v : my) v = 0.2 + U(rng);
        e->set_measurements(my);
        e->set_noise_scales(sg);
        e->id = "cond" + std::to_string(i);
        if (scenario == 1) {
            if (i % 4 == 1) {   // own, shorter time grid
                std::vector<realtype> t2 = {0.0, 0.5 * ts[3], ts[5], 1.3 * ts[7]};
                e->set_timepoints(t2);
                e->set_measurements(std::vector<realtype>(t2.size() * m.nytrue, 0.4));
                e->set_noise_scales(std::vector<realtype>(t2.size() * m.nytrue, 0.2));
            }
            if (i % 4 == 2) e->plist = {0, 2};   // sensitivities w.r.t. two parameters only
            if (i % 5 == 0) {                    // log10-scaled parameters
                e->pscale = std::vector<ParameterScaling>(m.np(), ParameterScaling::log10);
                for (auto& v : e->free_parameters) v = std::log10(v);
            }
            if (i == 7) {                        // a scale WITHOUT parameters: the model's values are re-scaled (Model::set_parameter_scale)
                e->pscale = std::vector<ParameterScaling>(m.np(), ParameterScaling::ln);
                e->free_parameters.clear();
            }
            if (i == 9) e->fixed_parameters.clear();   // the model's fixed parameters
            if (i == 11) {
                std::vector<realtype> x0(m.nx_rdata);
                for (auto& v : x0) v = 0.1 + U(rng);
                e->x0 = x0;
            }
        }
        if (scenario == 2 && m.nk() > 0 && m.get_name().find("steadystate") != std::string::npos) {
            // pre-equilibration (Newton), presimulation, re-initialisation of states 0 and 2
            c.solver->set_newton_max_steps(50);
            std::vector<realtype> kp(k0);
            for (auto& v : kp) v *= 0.6 + U(rng);
            if (i % 2 == 0) e->fixed_parameters_pre_equilibration = kp;
            if (i % 3 == 0) {
                e->t_presim = 0.5 + U(rng);
                std::vector<realtype> ks(k0);
                for (auto& v : ks) v *= 0.6 + U(rng);
                e->fixed_parameters_presimulation = ks;
            }
            if (i % 4 == 0) e->reinitialization_state_idxs_sim = {0, 2};
        }
        if (scenario == 3 && i == 4) {   // this one runs out of steps: status, NaN tail, message, and the failfast tail after it
            e->free_parameters[0] *= 1.0;
        }
        c.edatas.push_back(e.get());
        c.owned.push_back(std::move(e));
    }
    if (scenario == 3) c.solver->set_max_steps(200);   // jakstat needs ~1400: every condition fails
    return c;
}

struct Flat {
    std::vector<double> d;
    void add(std::vector<realtype> const& v) { d.push_back((double)v.size()); d.insert(d.end(), v.begin(), v.end()); }
    void addi(std::vector<int> const& v) { d.push_back((double)v.size()); for (int x : v) d.push_back(x); }
    void add1(double v) { d.push_back(1.0); d.push_back(v); }
};

// every numeric ReturnData field on this path, in a fixed order (tests/test_gpu.py knows the names)
void flatten(ReturnData const& r, Flat& f) {
    f.add1(r.status); f.add1(r.llh); f.add1(r.chi2); f.add(r.sllh); f.add(r.ts); f.add(r.x); f.add(r.sx); f.add(r.y); f.add(r.sy);
    f.add(r.sigmay); f.add(r.ssigmay); f.add(r.res); f.add(r.sres); f.add(r.FIM); f.add(r.x0); f.add(r.sx0); f.add(r.xdot); f.add(r.J);
    f.addi(r.numsteps); f.addi(r.num_rhs_evals); f.addi(r.num_err_test_fails); f.addi(r.num_non_lin_solv_conv_fails); f.addi(r.order);
    f.add(r.x_ss); f.add(r.sx_ss);
    std::vector<int> ps, qs;
    for (auto s : r.preeq_status) ps.push_back(static_cast<int>(s));
    for (auto s : r.posteq_status) qs.push_back(static_cast<int>(s));
    f.addi(ps); f.addi(r.preeq_numsteps); f.add1(r.preeq_wrms); f.add1(r.preeq_t);
    f.addi(qs); f.addi(r.posteq_numsteps); f.add1(r.posteq_wrms); f.add1(r.posteq_t);
    f.add(r.z); f.add(r.sz); f.add(r.rz); f.add(r.srz); f.add(r.sigmaz);
    f.add1((double)r.messages.size()); f.add1((double)r.nplist); f.add1((double)r.nt);
    f.add1(r.id.empty() ? -1.0 : std::atof(r.id.c_str() + 4));
}
}  // namespace

// which = 0: amici::run_simulations (reference, host), 1: amici::run_simulations_gpu.  Writes the flattened results of all
// conditions to out (capacity cap doubles); returns the number of doubles written, -1 on overflow, -2 on an exception.
extern "C" long binding_run(int which, int n, unsigned seed, int scenario, int failfast, int num_threads, int linsol, double rtol, double atol,
                            double* out, long cap, char* err, int errcap) {
    try {
        // scenario 10 + m: the plain list of scenario 0 in RDataReporting mode m
        Case c = scenario >= 10 ? make_case(n, seed, 0, linsol, rtol, atol, scenario - 10) : make_case(n, seed, scenario, linsol, rtol, atol);
        auto res = which == 0 ? run_simulations(*c.solver, c.edatas, *c.model, failfast != 0, num_threads)
                              : run_simulations_gpu(*c.solver, c.edatas, *c.model, failfast != 0, num_threads);
        Flat f;
        for (auto const& r : res) flatten(*r, f);
        if ((long)f.d.size() > cap) return -1;
        std::memcpy(out, f.d.data(), f.d.size() * sizeof(double));
        return (long)f.d.size();
    } catch (std::exception const& ex) {
        std::snprintf(err, errcap, "%s", ex.what());
        return -2;
    }
}
