// amici_gpu.cpp — amici::run_simulations_gpu: AMICI's own C++ types on one side, the C ABI of the B200 engine
// (include/amici_b200.h) on the other.  Compiled against the UNMODIFIED reference headers and linked with libamici; it
// contains no numerics: ExpData -> ab_condition (borrowed pointers into the ExpData's vectors), Solver -> ab_solver_options,
// ReturnData (constructed by the reference's own constructor, so every vector has the reference's size and default) ->
// ab_batch output pointers into its std::vector members.  Grouping of compatible conditions, struct-of-arrays assembly and the
// `failfast` tail are done behind the ABI (ab_run_conditions, amici_b200/csrc/conditions.cpp).
#include "amici_gpu.h"

#include <amici/exception.h>
#include <amici_b200.h>

#include <cmath>
#include <cstring>

namespace amici {

std::string gpu_model_name(Model const& model) {
    std::string n = model.get_name();
    if (n.rfind("model_", 0) == 0) n = n.substr(6);
    if (n.size() > 3 && n.compare(n.size() - 3, 3, "_py") == 0) n.resize(n.size() - 3);
    return n;
}

namespace {

ab_solver_options solver_options(Solver const& solver, Model const& model) {
    ab_solver_options o;
    std::memset(&o, 0, sizeof o);
    if (solver.get_sensitivity_order() > SensitivityOrder::first)
        throw AmiException("amici_b200: second-order sensitivities are not implemented");
    if (solver.get_sensitivity_order() == SensitivityOrder::first && solver.get_sensitivity_method() == SensitivityMethod::none)
        throw AmiException("amici_b200: sensitivity order first needs a sensitivity method");
    o.rtol = solver.get_relative_tolerance();
    o.atol = solver.get_absolute_tolerance();
    o.rtol_fsa = solver.get_relative_tolerance_fsa();
    o.atol_fsa = solver.get_absolute_tolerance_fsa();
    o.t0 = model.t0();
    o.hmax_inv = solver.get_max_step_size() > 0 ? 1.0 / solver.get_max_step_size() : 0.0;
    o.sensi = solver.get_sensitivity_order() == SensitivityOrder::first ? 1 : 0;
    o.ism = static_cast<int>(solver.get_internal_sensitivity_method());
    o.sensi_meth = static_cast<int>(solver.get_sensitivity_method());   // 1 forward, 2 adjoint (plain simulations of event-free models)
    o.maxsteps = static_cast<int>(solver.get_max_steps());
    o.stldet = solver.get_stability_limit_flag() ? 1 : 0;
    o.max_nonlin_iters = solver.get_max_nonlin_iters();
    o.max_conv_fails = solver.get_max_conv_fails();
    o.max_err_test_fails = 7;   // CVODES default; AMICI does not change it
    o.reporting = static_cast<int>(solver.get_return_data_reporting_mode());
    o.rtol_ss = solver.get_relative_tolerance_steady_state();
    o.atol_ss = solver.get_absolute_tolerance_steady_state();
    o.rtol_ss_sensi = solver.get_relative_tolerance_steady_state_sensi();
    o.atol_ss_sensi = solver.get_absolute_tolerance_steady_state_sensi();
    o.newton_damping_lower_bound = solver.get_newton_damping_factor_lower_bound();
    o.newton_maxsteps = solver.get_newton_max_steps();
    o.newton_damping_mode = static_cast<int>(solver.get_newton_damping_factor_mode());
    o.sensi_ss_check = solver.get_sensi_steady_state_check() ? 1 : 0;
    o.ss_mode = static_cast<int>(model.get_steady_state_computation_mode());
    o.ss_sensi_mode = static_cast<int>(model.get_steady_state_sensitivity_mode());
    o.nmaxevent = model.n_max_event();
    o.maxtime = solver.get_max_time();   // Solver::set_max_time, seconds; 0 = unlimited
    o.x_nonneg = 0ull;                   // Model::set_state_is_non_negative as a bit set
    auto const& nonneg = model.get_state_is_non_negative();
    for (size_t i = 0; i < nonneg.size() && i < 64; ++i)
        if (nonneg[i]) o.x_nonneg |= 1ull << i;
    return o;
}

template <class T>
T* ptr(std::vector<T>& v) { return v.empty() ? nullptr : v.data(); }

// output pointers of one condition: into the vectors the reference's ReturnData constructor has sized
void bind_outputs(ReturnData& r, std::vector<int>& preeq_status, std::vector<int>& posteq_status, ab_batch& b) {
    std::memset(&b, 0, sizeof b);
    b.status = &r.status;
    // which objective scalars a reporting mode computes: ReturnData::initialize_objective_function (reference src/rdata.cpp:970-982)
    // llh is accumulated in every mode (in `residuals` it starts from the member's default 0 and is never reset); chi2 only where it
    // is initialised to 0, elsewhere it stays NaN
    auto const mode = r.rdata_reporting;
    b.llh = &r.llh;
    if (mode == RDataReporting::full || mode == RDataReporting::residuals) b.chi2 = &r.chi2;
    b.sllh = ptr(r.sllh);
    b.x = ptr(r.x);
    b.sx = ptr(r.sx);
    b.y = ptr(r.y);
    b.sy = ptr(r.sy);
    b.sigmay = ptr(r.sigmay);
    b.ssigmay = ptr(r.ssigmay);
    b.res = ptr(r.res);
    b.sres = ptr(r.sres);
    b.FIM = ptr(r.FIM);
    b.x0 = ptr(r.x0);
    b.sx0 = ptr(r.sx0);
    b.xdot = ptr(r.xdot);
    b.J = ptr(r.J);
    b.numsteps = ptr(r.numsteps);
    b.numrhsevals = ptr(r.num_rhs_evals);
    b.numerrtestfails = ptr(r.num_err_test_fails);
    b.numnonlinsolvconvfails = ptr(r.num_non_lin_solv_conv_fails);
    b.order = ptr(r.order);
    b.x_ss = ptr(r.x_ss);
    b.sx_ss = ptr(r.sx_ss);
    preeq_status.assign(3, 0);
    posteq_status.assign(3, 0);
    b.preeq_status = preeq_status.data();   // SteadyStateStatus is an enum class: converted after the call
    b.posteq_status = posteq_status.data();
    b.preeq_numsteps = ptr(r.preeq_numsteps);
    b.posteq_numsteps = ptr(r.posteq_numsteps);
    b.preeq_wrms = &r.preeq_wrms;
    b.preeq_t = &r.preeq_t;
    b.posteq_wrms = &r.posteq_wrms;
    b.posteq_t = &r.posteq_t;
    b.z = ptr(r.z);
    b.sz = ptr(r.sz);
    b.rz = ptr(r.rz);
    b.srz = ptr(r.srz);
    b.sigmaz = ptr(r.sigmaz);
}

char const* status_text(int status) {
    switch (status) {
    case AMICI_TOO_MUCH_WORK: return "mxstep steps taken before reaching tout";
    case AMICI_TOO_MUCH_ACC: return "could not satisfy the accuracy demanded by the user";
    case AMICI_ERR_FAILURE: return "error test failures occurred too many times during one internal time step";
    case AMICI_CONV_FAILURE: return "convergence test failures occurred too many times during one internal time step";
    case AMICI_RHSFUNC_FAIL: return "the right-hand side function failed in an unrecoverable manner";
    case AMICI_FIRST_RHSFUNC_ERR: return "the right-hand side function failed at the first call";
    case -10: return "repeated recoverable right-hand side function errors";                                            // CV_REPTD_RHSFUNC_ERR
    case -11: return "the right-hand side function had a recoverable error, but no recovery was possible";              // CV_UNREC_RHSFUNC_ERR
    case AMICI_ILL_INPUT: return "illegal input (e.g. a root found at and very near the initial time)";
    default: return "simulation failed";
    }
}

}  // namespace

std::vector<std::unique_ptr<ReturnData>> run_simulations_gpu(
    Solver const& solver, std::vector<ExpData*> const& edatas, Model const& model, bool const failfast, int /* num_threads */) {
    int const n = static_cast<int>(edatas.size());
    std::vector<std::unique_ptr<ReturnData>> results(n);
    if (n == 0) return results;
    std::string const name = gpu_model_name(model);
    ab_solver_options const opt = solver_options(solver, model);

    // ---- the model's own settings: what an ExpData field that is not set falls back to ----
    auto const m_ts = model.get_timepoints();
    auto const m_plist = model.get_parameter_list();
    std::vector<int> m_pscale;
    for (auto s : model.get_parameter_scale()) m_pscale.push_back(static_cast<int>(s));
    auto const m_p = model.get_free_parameters();
    auto const m_k = model.get_fixed_parameters();
    auto const m_reinit = model.get_reinitialization_state_idxs();
    ab_model_state ms;
    std::memset(&ms, 0, sizeof ms);
    ms.nt = static_cast<int>(m_ts.size());
    ms.ts = m_ts.data();
    ms.nplist = static_cast<int>(m_plist.size());
    ms.plist = m_plist.data();
    ms.pscale = m_pscale.data();
    ms.p = m_p.data();
    ms.k = m_k.data();
    ms.t0 = model.t0();
    ms.reinit_all = model.get_reinitialize_fixed_parameter_initial_states() ? 1 : 0;
    ms.n_reinit = static_cast<int>(m_reinit.size());
    ms.reinit_idxs = m_reinit.data();

    // ---- ExpData -> ab_condition (borrowed pointers), ReturnData -> output pointers ----
    std::vector<ab_condition> conds(n);
    std::vector<ab_batch> outs(n);
    std::vector<std::vector<int>> pscales(n), preeq_st(n), posteq_st(n);
    for (int i = 0; i < n; ++i) {
        ab_condition& c = conds[i];
        std::memset(&c, 0, sizeof c);
        ExpData const* e = edatas[i];
        // ReturnData exactly as the reference constructs it for this condition: ConditionContext applied to a clone of the model
        // (timepoints, plist and scales decide the sizes), restored when the context goes out of scope (src/amici.cpp:86-90)
        {
            auto my_model = std::unique_ptr<Model>(model.clone());
            ConditionContext cc(my_model.get(), e, FixedParameterContext::simulation);
            results[i] = std::make_unique<ReturnData>(solver, *my_model);
        }
        if (e) results[i]->id = e->id;
        bind_outputs(*results[i], preeq_st[i], posteq_st[i], outs[i]);
        if (!e) {
            // no data: the reference computes no objective (ReturnData::process_simulation_objects only does so `if (edata)`,
            // src/rdata.cpp:213-260): llh / chi2 / sllh / res / sres / FIM keep the constructor's values (NaN / 0)
            c.none = 1;
            ab_batch& o = outs[i];
            o.llh = o.chi2 = o.sllh = o.res = o.sres = o.FIM = nullptr;
            continue;
        }
        if (e->nytrue() != model.nytrue || e->nztrue() != model.nztrue)
            throw AmiException("ExpData dimensions do not match the model");
        c.nt = e->nt();
        c.ts = c.nt ? e->get_timepoints().data() : nullptr;
        c.nplist = static_cast<int>(e->plist.size());
        c.plist = e->plist.empty() ? nullptr : e->plist.data();
        if (!e->pscale.empty()) {
            if (e->pscale.size() != static_cast<size_t>(model.np()))
                throw AmiException("Number of parameters (%d) in model does not match ExpData (%zd).", model.np(), e->pscale.size());
            for (auto s : e->pscale) pscales[i].push_back(static_cast<int>(s));
            c.pscale = pscales[i].data();
        }
        auto sized = [&](std::vector<realtype> const& v, size_t want, char const* what) -> realtype const* {
            if (v.empty()) return nullptr;
            if (v.size() != want) throw AmiException("Size of ExpData::%s (%zd) does not match the model (%zd).", what, v.size(), want);
            return v.data();
        };
        size_t const nplist_eff = e->plist.empty() ? m_plist.size() : e->plist.size();
        c.p = sized(e->free_parameters, model.np(), "free_parameters");
        c.k = sized(e->fixed_parameters, model.nk(), "fixed_parameters");
        c.k_preeq = sized(e->fixed_parameters_pre_equilibration, model.nk(), "fixed_parameters_pre_equilibration");
        c.k_presim = sized(e->fixed_parameters_presimulation, model.nk(), "fixed_parameters_presimulation");
        c.x0 = sized(e->x0, model.nx_rdata, "x0");
        c.sx0 = sized(e->sx0, model.nx_rdata * nplist_eff, "sx0");
        c.t_presim = e->t_presim;
        c.t_start = e->t_start;
        c.my = e->get_measurements_ptr(0);
        c.sigmay = e->get_noise_scales_ptr(0);
        c.mz = e->get_event_measurements_ptr(0);
        c.sigmaz = e->get_event_noise_scales_ptr(0);
        c.reinit_all = e->reinitialize_fixed_parameter_initial_states ? 1 : 0;
        c.n_reinit_sim = static_cast<int>(e->reinitialization_state_idxs_sim.size());
        c.reinit_sim = e->reinitialization_state_idxs_sim.data();
        c.n_reinit_presim = static_cast<int>(e->reinitialization_state_idxs_presim.size());
        c.reinit_presim = e->reinitialization_state_idxs_presim.data();
    }

    int const rc = ab_run_conditions(name.c_str(), &opt, &ms, n, conds.data(), outs.data(), failfast ? 1 : 0, nullptr);
    if (rc != AB_OK) throw AmiException("amici_b200 (%d): %s", rc, ab_last_error());

    for (int i = 0; i < n; ++i) {
        ReturnData& r = *results[i];
        if (r.status == AMICI_NOT_RUN) continue;   // failfast tail: a freshly constructed ReturnData, as in the reference
        for (int j = 0; j < 3 && j < static_cast<int>(r.preeq_status.size()); ++j) r.preeq_status[j] = static_cast<SteadyStateStatus>(preeq_st[i][j]);
        for (int j = 0; j < 3 && j < static_cast<int>(r.posteq_status.size()); ++j) r.posteq_status[j] = static_cast<SteadyStateStatus>(posteq_st[i][j]);
        if (r.status < 0) {
            // one message per failed condition, in the reference's words (src/amici.cpp:112-130); t_last = the last time reported
            double t_fail = NAN;
            for (int it = 0; it < r.nt; ++it)
                if (!r.numsteps.empty() && (it == 0 || r.numsteps[it] > 0)) t_fail = r.ts[it];
            char buf[512];
            if (r.status != AMICI_ERROR) {
                // the reference's SUNDIALS error handler logs the integrator's own message first (debug severity,
                // src/solver.cpp:18-58: "<file stem>:<function>:<code name>")
                std::string code = simulation_status_to_str(r.status);
                if (code.rfind("AMICI_", 0) == 0) code = code.substr(6);
                std::snprintf(buf, sizeof buf, "At t = %g, %s.", t_fail, status_text(r.status));
                r.messages.emplace_back(LogSeverity::debug, "cvodes:CVode:" + code, buf);
            }
            std::snprintf(buf, sizeof buf, "AMICI forward simulation failed at t = %g: %s (status %d)", t_fail, status_text(r.status), r.status);
            r.messages.emplace_back(LogSeverity::error, "FORWARD_FAILURE", buf);
        }
    }
    return results;
}

}  // namespace amici
