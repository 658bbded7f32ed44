// Steady-state (pre-equilibration) pieces of one condition: Newton's method with damping, the
// convergence test of the steady-state simulation, and the steady-state sensitivities from one
// linear solve per parameter.
//
// Behavioural specification: reference src/steadystateproblem.cpp:29-205 (getWrmsNorm, NewtonsMethod),
// src/newton_solver.cpp:44-107 (get_step, compute_newton_sensis, prepare_linear_system),
// src/forwardproblem.cpp:755-1223 (SteadyStateProblem).  Everything here is rare, out-of-line code that
// works on the trajectory's shared-memory slice: the state is column 0 of zn[0], the sensitivities the
// other columns; the dense LU of J (not I - gamma*J) is built in S.lu with partial pivoting
// (SUNLinSol_Dense semantics; the reference's default KLU differs only in pivot order).
#ifndef AMICI_B200_STEADYSTATE_CUH
#define AMICI_B200_STEADYSTATE_CUH
#include "bdf_core.cuh"

namespace amici_b200 {

// SteadyStateStatus (reference include/amici/defines.h:218-226)
enum : int {
    SS_FAILED_TOO_LONG = -5,
    SS_FAILED_DAMPING = -4,
    SS_FAILED_FACTORIZATION = -3,
    SS_FAILED_CONVERGENCE = -2,
    SS_FAILED = -1,
    SS_NOT_RUN = 0,
    SS_SUCCESS = 1,
};
// SteadyStateComputationMode / SteadyStateSensitivityMode (reference include/amici/defines.h:204-215)
enum : int { SS_NEWTON_ONLY = 0, SS_INTEGRATION_ONLY = 1, SS_INTEGRATE_IF_NEWTON_FAILS = 2 };

// spline values at time t for the lane's own use (no group communication)
template <class M, int G, int C>
AB_HD void spline_values_at(const typename Layout<M, G, C>::Shared& S, double t, double* spl) {
    constexpr int NCOLS = Layout<M, G, C>::NCOLS;
    AB_UNROLL
    for (int is = 0; is < M::NSPL; ++is) {
        const double v = spline_eval_scaled<M, NCOLS>(S.spl, is, 0, t, true);
        spl[is] = spline_is_log<M>(is) ? exp(v) : v;
    }
}

// getWrmsNorm (reference src/steadystateproblem.cpp:29-52) with an empty mask
template <int NX>
AB_HD double ss_wrms(const double* v, const double* xref, double rtol, double atol) {
    double s = 0.0;
    for (int r = 0; r < NX; ++r) {
        const double ew = 1.0 / (rtol * fabs(xref[r]) + atol);
        const double pr = v[r] * ew;
        s += pr * pr;
    }
    return sqrt(s / NX);
}

// NewtonsMethod::run (reference src/steadystateproblem.cpp:74-142) on the state column, at time t
// (the reference passes t = infinity).  Executed by lane 0, results broadcast to the group.
// Returns AB_SUCCESS, AB_TOO_MUCH_WORK, AB_DAMPING_FACTOR_ERROR or AB_SINGULAR_JACOBIAN.
template <class M, int G, int C>
AB_COLD int cold_newton(typename Layout<M, G, C>::Shared& S, const SolverOptions& o, double t, int* nsteps,
                        double* wrms_out) {
    constexpr int NX = M::NX;
    Group<G> g;
    int ret = AB_SUCCESS, i_step = 0;
    double wrms = 0.0;
    g.sync();
    if (g.lane == 0) {
        double x[NX], x_old[NX], xdot[NX], delta[NX], xp[NX], spl[M::NSPL > 0 ? M::NSPL : 1];   // xp: what the model functions see (x_pos)
        double Jv[M::NNZ_J > 0 ? M::NNZ_J : 1];
        spline_values_at<M, G, C>(S, t, spl);
        for (int r = 0; r < NX; ++r) { x[r] = S.zn[0][r][0]; x_old[r] = x[r]; delta[r] = 0.0; }
        for (int r = 0; r < NX; ++r) xp[r] = x[r];
        x_pos<M>(S, xp);
        M::xdot(t, xp, S.p, S.k, S.h, spl, xdot);
        wrms = ss_wrms<NX>(xdot, x, o.rtol_ss, o.atol_ss);
        bool converged = wrms < 1.0;
        double gamma = 1.0;
        bool update_direction = true, step_successful = true;
        while (!converged && i_step < o.newton_maxsteps) {
            if (step_successful)
                for (int r = 0; r < NX; ++r) x_old[r] = x[r];
            if (update_direction) {
                // NewtonSolver::get_step: delta = -J(x)^{-1} xdot
                for (int r = 0; r < NX; ++r) xp[r] = x[r];
                x_pos<M>(S, xp);
                M::jac(t, xp, S.p, S.k, S.h, spl, Jv);
                M::jac_to_dense(Jv, &S.lu[0][0]);
                bool fin = true;
                for (int i = 0; i < M::NNZ_J; ++i) fin = fin && isfinite(Jv[i]);
                if (!fin || !dense_lu_inplace<M, G, C>(S)) { ret = AB_SINGULAR_JACOBIAN; break; }
                for (int r = 0; r < NX; ++r) delta[r] = -xdot[r];
                cold_dense_solve<M, G, C>(S, delta);
            }
            for (int r = 0; r < NX; ++r) x[r] = x_old[r] + gamma * delta[r];
            for (int r = 0; r < NX; ++r) xp[r] = x[r];
            x_pos<M>(S, xp);
            M::xdot(t, xp, S.p, S.k, S.h, spl, xdot);
            const double wrms_tmp = ss_wrms<NX>(xdot, x, o.rtol_ss, o.atol_ss);
            step_successful = wrms_tmp < wrms;
            if (step_successful) {
                wrms = wrms_tmp;
                converged = wrms < 1.0;
            }
            // update_damping_factor (:150-168)
            if (o.newton_damping_mode == 1) {
                gamma = step_successful ? ab_min(1.0, 2.0 * gamma) : gamma / 4.0;
                if (gamma < o.newton_damping_lower_bound) { ret = AB_DAMPING_FACTOR_ERROR; break; }
                update_direction = step_successful;
            } else {
                update_direction = true;
            }
            ++i_step;
        }
        if (ret == AB_SUCCESS && !converged) ret = AB_TOO_MUCH_WORK;
        for (int r = 0; r < NX; ++r) S.zn[0][r][0] = x[r];
    }
    g.sync();
    ret = g.bcast(ret, 0);
    *nsteps = g.bcast(i_step, 0);
    *wrms_out = g.bcast(wrms, 0);
    return ret;
}

// Convergence test of the steady-state simulation (check_convergence + get_wrms_fsa, reference
// src/forwardproblem.cpp:1044-1078, :984-1012): WRMS of xdot weighted by the state with the steady-state
// tolerances; with check_sens the same for every sensitivity right-hand side J*s_ip + df/dp_ip.
// The current solver state is zn[0].
template <class M, int G, int C>
AB_COLD bool cold_ss_converged(typename Layout<M, G, C>::Shared& S, const SolverOptions& o, int ncols, double t,
                               bool check_sens, double* wrms_out) {
    constexpr int NX = M::NX;
    Group<G> g;
    ModelEval<M, C> e;
    double yy[C][NX], f[C][NX];
    AB_UNROLL
    for (int c = 0; c < C; ++c) {
        const int cc = g.lane + G * c;
        AB_UNROLL
        for (int r = 0; r < NX; ++r) yy[c][r] = (cc < ncols) ? S.zn[0][r][cc] : 0.0;
    }
    const int ncols_eval = check_sens ? ncols : 1;
    model_rhs<M, G, C>(S, g, ncols_eval, t, yy, f, e);
    double w0 = 0.0;
    bool gt1 = false, last_bad = false;
    AB_UNROLL
    for (int c = 0; c < C; ++c) {
        const int cc = g.lane + G * c;
        if (cc >= ncols_eval) continue;
        const double w = (cc == 0) ? ss_wrms<NX>(f[c], yy[c], o.rtol_ss, o.atol_ss)
                                   : ss_wrms<NX>(f[c], yy[c], o.rtol_ss_sensi, o.atol_ss_sensi);
        if (cc == 0) w0 = w;
        else {
            gt1 = gt1 || (w > 1.0);
            if (cc == ncols_eval - 1) last_bad = !(w < 1.0);
        }
    }
    w0 = g.bcast(w0, 0);
    *wrms_out = w0;
    if (!(w0 < 1.0)) return false;
    if (!check_sens || ncols_eval <= 1) return true;
    return !g.any(gt1) && !g.any(last_bad);
}

// NewtonSolver::compute_newton_sensis (reference src/newton_solver.cpp:53-76): sx_ip = -J(x)^{-1} df/dp_ip at the
// state in column 0 of zn[0]; writes the sensitivity columns of zn[0]. False if J cannot be factorised.
template <class M, int G, int C>
AB_COLD bool cold_newton_sensis(typename Layout<M, G, C>::Shared& S, int ncols, double t) {
    constexpr int NX = M::NX;
    Group<G> g;
    ModelEval<M, C> e;
    double yy[C][NX];
    AB_UNROLL
    for (int c = 0; c < C; ++c) {
        const int cc = g.lane + G * c;
        AB_UNROLL
        for (int r = 0; r < NX; ++r) yy[c][r] = (cc < ncols) ? S.zn[0][r][cc] : 0.0;
    }
    bcast_state<M, G, C>(S, g, yy, e);
    eval_splines<M, G, C>(S, g, ncols, t, e);
    int ok = 1;
    g.sync();
    if (g.lane == 0) {
        double Jv[M::NNZ_J > 0 ? M::NNZ_J : 1];
        M::jac(t, e.y, S.p, S.k, S.h, e.spl, Jv);
        M::jac_to_dense(Jv, &S.lu[0][0]);
        bool fin = true;
        for (int i = 0; i < M::NNZ_J; ++i) fin = fin && isfinite(Jv[i]);
        ok = (fin && dense_lu_inplace<M, G, C>(S)) ? 1 : 0;
    }
    g.sync();
    ok = g.bcast(ok, 0);
    if (!ok) return false;
    AB_UNROLL
    for (int c = 0; c < C; ++c) {
        const int cc = g.lane + G * c;
        if (cc >= ncols || cc == 0) continue;
        double b[NX];
        M::dxdotdp(S.ip_of_col[cc], t, e.y, S.p, S.k, S.h, e.spl, e.sspl[c], b);
        for (int r = 0; r < NX; ++r) b[r] = -b[r];
        cold_dense_solve<M, G, C>(S, b);
        for (int r = 0; r < NX; ++r) S.zn[0][r][cc] = b[r];
    }
    g.sync();
    return true;
}

}  // namespace amici_b200
#endif
