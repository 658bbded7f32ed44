// oracle/ref_driver.cpp — TEST INFRASTRUCTURE (checker), not product code.
//
// Drives the UNMODIFIED reference through its own public C++ API
// (amici::run_simulations, reference include/amici/amici.h:36-39) on a flat SoA
// batch. One copy of this file is linked per pre-generated model.
#include "ref_driver.h"

#include "wrapfunctions.h"
#include <amici/amici.h>
#include <amici/model_ode.h>

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <memory>
#include <omp.h>
#include <vector>

using namespace amici;

namespace {
template <class T>
void put(T* dst, size_t stride, size_t i, std::vector<T> const& src) {
    if (!dst) return;
    size_t n = std::min(stride, src.size());
    std::copy_n(src.begin(), n, dst + i * stride);
    // anything the reference did not report (reporting mode) stays NaN / 0
}
void fill_nan(double* p, size_t n) {
    if (p) std::fill_n(p, n, std::nan(""));
}
} // namespace

extern "C" void ref_get_dims(ref_dims* d) {
    auto m = generic_model::get_model();
    d->nx = m->nx_rdata; d->nx_solver = m->nx_solver; d->np = m->np(); d->nk = m->nk();
    d->ny = m->ny; d->nytrue = m->nytrue; d->nz = m->nz; d->nztrue = m->nztrue;
    d->ne = m->ne; d->ne_solver = m->ne_solver; d->nw = m->nw; d->nspl = m->nspl;
    d->is_ode = dynamic_cast<Model_ODE*>(m.get()) != nullptr;
}

extern "C" void ref_get_nominal(double* p, double* k) {
    auto m = generic_model::get_model();
    auto const& up = m->get_unscaled_parameters();
    auto const& fk = m->get_fixed_parameters();
    std::copy(up.begin(), up.end(), p);
    std::copy(fk.begin(), fk.end(), k);
}

extern "C" int ref_max_threads(void) { return omp_get_max_threads(); }

extern "C" int ref_run(ref_options const* o, ref_batch* b) {
    try {
        auto model = generic_model::get_model();
        int const np = model->np(), nk = model->nk();
        int const nx = model->nx_rdata, nxs = model->nx_solver, ny = model->ny;
        int const nyt = model->nytrue, nz = model->nz, nzt = model->nztrue;
        int const nt = b->nt, npl = b->nplist, N = b->N;

        std::vector<realtype> ts(b->ts, b->ts + nt);
        model->set_timepoints(ts);
        model->set_t0(o->t0);
        if (o->nmaxevent >= 0) model->set_n_max_event(o->nmaxevent);
        if (o->ss_mode >= 0)
            model->set_steady_state_computation_mode(static_cast<SteadyStateComputationMode>(o->ss_mode));
        if (o->ss_sensi_mode >= 0)
            model->set_steady_state_sensitivity_mode(static_cast<SteadyStateSensitivityMode>(o->ss_sensi_mode));
        std::vector<ParameterScaling> pscale(np);
        for (int i = 0; i < np; ++i) pscale[i] = static_cast<ParameterScaling>(b->pscale ? b->pscale[i] : 0);
        model->set_parameter_scale(pscale);
        std::vector<int> plist(b->plist, b->plist + npl);
        model->set_parameter_list(plist);
        int const nmaxevent = model->n_max_event();

        auto solver = model->create_solver();
        solver->set_relative_tolerance(o->rtol);
        solver->set_absolute_tolerance(o->atol);
        if (!std::isnan(o->rtol_fsa)) solver->set_relative_tolerance_fsa(o->rtol_fsa);
        if (!std::isnan(o->atol_fsa)) solver->set_absolute_tolerance_fsa(o->atol_fsa);
        if (!std::isnan(o->ss_tol_factor)) solver->set_steady_state_tolerance_factor(o->ss_tol_factor);
        if (!std::isnan(o->rtol_ss)) solver->set_relative_tolerance_steady_state(o->rtol_ss);
        if (!std::isnan(o->atol_ss)) solver->set_absolute_tolerance_steady_state(o->atol_ss);
        solver->set_sensitivity_order(static_cast<SensitivityOrder>(o->sensi_order));
        solver->set_sensitivity_method(static_cast<SensitivityMethod>(o->sensi_meth));
        solver->set_internal_sensitivity_method(static_cast<InternalSensitivityMethod>(o->ism));
        solver->set_linear_solver(static_cast<LinearSolver>(o->linsol));
        solver->set_max_steps(o->maxsteps);
        solver->set_newton_max_steps(o->newton_maxsteps);
        if (o->stldet >= 0) solver->set_stability_limit_flag(o->stldet != 0);
        if (o->maxtime > 0) solver->set_max_time(o->maxtime);
        if (o->x_nonneg) {
            std::vector<bool> nn(model->nx_solver, false);
            for (int ix = 0; ix < model->nx_solver && ix < 31; ++ix) nn[ix] = (o->x_nonneg >> ix & 1) != 0;
            model->set_state_is_non_negative(nn);
        }
        solver->set_return_data_reporting_mode(static_cast<RDataReporting>(o->reporting));

        std::vector<std::unique_ptr<ExpData>> owned;
        std::vector<ExpData*> edatas;
        owned.reserve(N);
        for (int i = 0; i < N; ++i) {
            auto e = std::make_unique<ExpData>(*model);
            e->set_timepoints(ts);
            e->free_parameters.assign(b->p + (size_t)i * np, b->p + (size_t)(i + 1) * np);
            e->fixed_parameters.assign(b->k + (size_t)i * nk, b->k + (size_t)(i + 1) * nk);
            if (b->k_preeq)
                e->fixed_parameters_pre_equilibration.assign(b->k_preeq + (size_t)i * nk,
                                                             b->k_preeq + (size_t)(i + 1) * nk);
            if (b->k_presim)
                e->fixed_parameters_presimulation.assign(b->k_presim + (size_t)i * nk, b->k_presim + (size_t)(i + 1) * nk);
            if (b->t_presim) e->t_presim = b->t_presim[i];
            e->reinitialize_fixed_parameter_initial_states = o->reinit_all != 0;
            for (int ix = 0; ix < 31; ++ix) {
                if (o->reinit_states_sim >> ix & 1) e->reinitialization_state_idxs_sim.push_back(ix);
                if (o->reinit_states_presim >> ix & 1) e->reinitialization_state_idxs_presim.push_back(ix);
            }
            e->pscale = pscale;
            e->plist = plist;
            if (b->x0_in) e->x0.assign(b->x0_in + (size_t)i * nx, b->x0_in + (size_t)(i + 1) * nx);
            if (b->sx0_in && npl > 0)
                e->sx0.assign(b->sx0_in + (size_t)i * npl * nx, b->sx0_in + (size_t)(i + 1) * npl * nx);
            if (b->my) {
                size_t n = (size_t)nt * nyt;
                e->set_measurements(std::vector<realtype>(b->my + i * n, b->my + (i + 1) * n));
            }
            if (b->sigmay_in) {
                size_t n = (size_t)nt * nyt;
                e->set_noise_scales(std::vector<realtype>(b->sigmay_in + i * n, b->sigmay_in + (i + 1) * n));
            }
            if (b->mz && nzt > 0) {
                size_t n = (size_t)nmaxevent * nzt;
                e->set_event_measurements(std::vector<realtype>(b->mz + i * n, b->mz + (i + 1) * n));
            }
            if (b->sigmaz_in && nzt > 0) {
                size_t n = (size_t)nmaxevent * nzt;
                e->set_event_noise_scales(std::vector<realtype>(b->sigmaz_in + i * n, b->sigmaz_in + (i + 1) * n));
            }
            edatas.push_back(e.get());
            owned.push_back(std::move(e));
        }

        auto t_begin = std::chrono::steady_clock::now();
        auto rdatas = run_simulations(*solver, edatas, *model, o->failfast != 0, o->num_threads);
        auto t_end = std::chrono::steady_clock::now();
        if (b->wall_seconds) *b->wall_seconds = std::chrono::duration<double>(t_end - t_begin).count();

        fill_nan(b->x, (size_t)N * nt * nx); fill_nan(b->sx, (size_t)N * nt * npl * nx);
        fill_nan(b->y, (size_t)N * nt * ny); fill_nan(b->sy, (size_t)N * nt * npl * ny);
        fill_nan(b->sigmay, (size_t)N * nt * ny); fill_nan(b->ssigmay, (size_t)N * nt * npl * ny);
        fill_nan(b->res, (size_t)N * nt * nyt); fill_nan(b->sres, (size_t)N * nt * nyt * npl);
        fill_nan(b->FIM, (size_t)N * npl * npl); fill_nan(b->sllh, (size_t)N * npl);
        fill_nan(b->x0, (size_t)N * nx); fill_nan(b->sx0, (size_t)N * npl * nx);
        fill_nan(b->xdot, (size_t)N * nxs); fill_nan(b->J, (size_t)N * nxs * nxs);
        fill_nan(b->z, (size_t)N * nmaxevent * nz); fill_nan(b->sz, (size_t)N * nmaxevent * npl * nz);
        fill_nan(b->rz, (size_t)N * nmaxevent * nz); fill_nan(b->srz, (size_t)N * nmaxevent * npl * nz);
        fill_nan(b->sigmaz, (size_t)N * nmaxevent * nz);
        fill_nan(b->x_ss, (size_t)N * nx); fill_nan(b->sx_ss, (size_t)N * npl * nx);

        for (int i = 0; i < N; ++i) {
            auto const& r = *rdatas[i];
            if (std::getenv("REF_VERBOSE"))   // diagnostic: the reference's per-simulation log (ReturnData::messages)
                for (auto const& msg : r.messages) fprintf(stderr, "ref[%d] %s: %s\n", i, msg.identifier.c_str(), msg.message.c_str());
            if (b->status) b->status[i] = r.status;
            if (b->llh) b->llh[i] = r.llh;
            if (b->chi2) b->chi2[i] = r.chi2;
            if (b->preeq_wrms) b->preeq_wrms[i] = r.preeq_wrms;
            if (b->preeq_t) b->preeq_t[i] = r.preeq_t;
            if (b->posteq_wrms) b->posteq_wrms[i] = r.posteq_wrms;
            if (b->posteq_t) b->posteq_t[i] = r.posteq_t;
            put(b->sllh, npl, i, r.sllh);
            put(b->x, (size_t)nt * nx, i, r.x);
            put(b->sx, (size_t)nt * npl * nx, i, r.sx);
            put(b->y, (size_t)nt * ny, i, r.y);
            put(b->sy, (size_t)nt * npl * ny, i, r.sy);
            put(b->sigmay, (size_t)nt * ny, i, r.sigmay);
            put(b->ssigmay, (size_t)nt * npl * ny, i, r.ssigmay);
            put(b->res, (size_t)nt * nyt, i, r.res);
            put(b->sres, (size_t)nt * nyt * npl, i, r.sres);
            put(b->FIM, (size_t)npl * npl, i, r.FIM);
            put(b->x0, nx, i, r.x0);
            put(b->sx0, (size_t)npl * nx, i, r.sx0);
            put(b->xdot, nxs, i, r.xdot);
            put(b->J, (size_t)nxs * nxs, i, r.J);
            put(b->z, (size_t)nmaxevent * nz, i, r.z);
            put(b->sz, (size_t)nmaxevent * npl * nz, i, r.sz);
            put(b->rz, (size_t)nmaxevent * nz, i, r.rz);
            put(b->srz, (size_t)nmaxevent * npl * nz, i, r.srz);
            put(b->sigmaz, (size_t)nmaxevent * nz, i, r.sigmaz);
            put(b->x_ss, nx, i, r.x_ss);
            put(b->sx_ss, (size_t)npl * nx, i, r.sx_ss);
            put(b->numsteps, nt, i, r.numsteps);
            put(b->numrhsevals, nt, i, r.num_rhs_evals);
            put(b->numerrtestfails, nt, i, r.num_err_test_fails);
            put(b->numnonlinsolvconvfails, nt, i, r.num_non_lin_solv_conv_fails);
            put(b->order, nt, i, r.order);
            put(b->preeq_numsteps, 3, i, r.preeq_numsteps);
            if (b->preeq_status)
                for (size_t j = 0; j < 3 && j < r.preeq_status.size(); ++j)
                    b->preeq_status[i * 3 + j] = static_cast<int>(r.preeq_status[j]);
            put(b->posteq_numsteps, 3, i, r.posteq_numsteps);
            if (b->posteq_status)
                for (size_t j = 0; j < 3 && j < r.posteq_status.size(); ++j)
                    b->posteq_status[i * 3 + j] = static_cast<int>(r.posteq_status[j]);
        }
        return 0;
    } catch (std::exception const& e) {
        fprintf(stderr, "ref_run: %s\n", e.what());
        return -1;
    }
}
