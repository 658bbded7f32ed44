// One condition end-to-end on one lane-group: the batched counterpart of the reference's
// run_simulation -> ForwardProblem -> EventHandlingSimulator::run -> ReturnData assembly
// (reference src/amici.cpp:65-218, src/forwardproblem.cpp:53-151, src/rdata.cpp:213-432, :749-875).
#ifndef AMICI_B200_SIMULATE_CUH
#define AMICI_B200_SIMULATE_CUH
#include "bdf_core.cuh"
#include "steadystate.cuh"

namespace amici_b200 {

AB_HD double ab_nan() {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double(0x7ff8000000000000LL);
#else
    return NAN;
#endif
}

// is_next_t_too_close (reference src/forwardproblem.cpp:26-38)
AB_HD bool next_t_too_close(double cur_t, double t_next) {
    const double tdiff = t_next - cur_t;
    if (tdiff == 0.0) return true;
    const double tround = DBL_EPSILON * ab_max(fabs(cur_t), fabs(t_next));
    return fabs(tdiff) < 2.0 * tround;
}

// PREEQ: compile the pre-equilibration phase (SteadyStateProblem) into the condition driver. It is a template
// flag, not a run-time branch, so that the kernel used without pre-equilibration keeps its register budget.
template <class M, int G, int C, bool PREEQ = false>
struct Simulation {
    using Solver = Bdf<M, G, C>;
    static constexpr int NX = M::NX, NY = M::NY, NCOLS = Solver::NCOLS;
    using Shared = typename Solver::Shared;

    Shared& S;
    const SolverOptions& o;
    const BatchView& b;
    Solver bdf;
    int n;       // trajectory index
    int nplist;
    double llh, chi2;
    double sllh[C];

    AB_HD Simulation(Shared& s, const SolverOptions& opt, const BatchView& bv, int n_)
        : S(s), o(opt), b(bv), bdf(s, opt, (opt.sensi >= 1 ? 1 + bv.nplist : 1)), n(n_),
          nplist(opt.sensi >= 1 ? bv.nplist : 0) {}

    AB_HD bool full() const { return o.reporting == 0; }

    // ConditionContext + Model::initialize (reference src/edata.cpp:438-556, src/model.cpp:251-278):
    // unscale parameters (reference include/amici/misc.h:164-195), x0, sx0, spline tables.
    AB_HD void load_condition(const bool preeq) {
        auto& g = bdf.g;
        g.sync();
        if (g.lane == 0) {
            for (int i = 0; i < M::NP; ++i) {
                const double v = b.p[(size_t)n * M::NP + i];
                const int sc = b.pscale ? b.pscale[i] : AB_PSCALE_NONE;
                S.p[i] = sc == AB_PSCALE_LOG10 ? pow(10.0, v) : sc == AB_PSCALE_LN ? exp(v) : v;
            }
            // ConditionContext with FixedParameterContext::preequilibration (reference src/edata.cpp:438-556)
            for (int i = 0; i < M::NK; ++i) S.k[i] = (preeq ? b.k_preeq : b.k)[(size_t)n * M::NK + i];
            for (int i = 0; i < M::NH; ++i) S.h[i] = 0.0;
            for (int cc = 0; cc < NCOLS; ++cc) {
                int ip = -1;
                double pc = 1.0;
                if (cc >= 1 && cc <= nplist) {
                    ip = b.plist[cc - 1];
                    const int sc = b.pscale ? b.pscale[ip] : AB_PSCALE_NONE;
                    // pcoefficient of ReturnData::apply_chain_rule_factor_to_simulation_results (rdata.cpp:783-796)
                    pc = sc == AB_PSCALE_LOG10 ? S.p[ip] * log(10.0) : sc == AB_PSCALE_LN ? S.p[ip] : 1.0;
                }
                S.ip_of_col[cc] = ip;
                S.pcoef[cc] = pc;
            }
        }
        g.sync();
        AB_UNROLL
        for (int c = 0; c < C; ++c) {
            if (!bdf.active(c)) continue;
            const int cc = bdf.col(c);
            double v[NX];
            if (cc == 0) M::x0(S.p, S.k, v);
            else M::sx0(S.ip_of_col[cc], S.p, S.k, v);
            AB_UNROLL
            for (int r = 0; r < NX; ++r) S.zn[0][r][cc] = v[r];
            AB_UNROLL
            for (int is = 0; is < M::NSPL; ++is)
                spline_build_column<M, NCOLS>(S.spl, is, cc, cc == 0 ? -1 : S.ip_of_col[cc], S.p, S.k);
        }
        g.sync();
    }

    AB_HD double chain(double v, double pc) const { return v != 0.0 ? v * pc : v; }

    // ReturnData::get_data_output + get_data_sensis_fsa for timepoint `it` from the solver state
    // (reference src/rdata.cpp:368-432; observables/likelihood: src/model.cpp:1112-1259, :2140-2413;
    //  residuals/FIM: src/rdata.cpp:985-1205).  Gaussian noise model (generated Jy.cpp/dJydy.cpp/dJydsigma.cpp).
    AB_HD void datapoint(int it) {
        auto& g = bdf.g;
        const int nt = b.nt;
        const double t = b.ts[it];
        typename Solver::Eval e;
        double xc[C][NX];
        AB_UNROLL
        for (int c = 0; c < C; ++c) {
            AB_UNROLL
            for (int r = 0; r < NX; ++r) xc[c][r] = bdf.active(c) ? S.zn[0][r][bdf.col(c)] : 0.0;
        }
        bcast_state<M, G, C>(S, g, xc, e);
        eval_splines<M, G, C>(S, g, bdf.ncols, t, e);
        double y[NY], sig[NY], my[NY];
        bool my_set[NY], sig_set[NY];
        M::y(t, e.y, S.p, S.k, S.h, e.spl, y);
        M::sigmay(S.p, S.k, y, sig);
        AB_UNROLL
        for (int iy = 0; iy < NY; ++iy) {
            const size_t idx = ((size_t)n * nt + it) * NY + iy;
            my[iy] = b.my ? b.my[idx] : ab_nan();
            my_set[iy] = (my[iy] == my[iy]);
            const double se = b.sigmay_in ? b.sigmay_in[idx] : ab_nan();
            sig_set[iy] = (se == se);
            if (sig_set[iy]) sig[iy] = se;
        }
        double res[NY];
        AB_UNROLL
        for (int iy = 0; iy < NY; ++iy) res[iy] = my_set[iy] ? (my[iy] - y[iy]) / sig[iy] : 0.0;

        AB_UNROLL
        for (int c = 0; c < C; ++c) {
            if (!bdf.active(c)) continue;
            const int cc = bdf.col(c);
            if (cc == 0) {
                if (b.x && full())
                    AB_UNROLL
                    for (int r = 0; r < NX; ++r) b.x[((size_t)n * nt + it) * NX + r] = e.y[r];
                AB_UNROLL
                for (int iy = 0; iy < NY; ++iy) {
                    const size_t idx = ((size_t)n * nt + it) * NY + iy;
                    if (b.y && full()) b.y[idx] = y[iy];
                    if (b.sigmay && full()) b.sigmay[idx] = sig[iy];
                    if (my_set[iy]) {
                        const double r_ = y[iy] - my[iy];
                        const double pi_ = 3.14159265358979323846;
                        llh -= 0.5 * log(2.0 * pi_ * sig[iy] * sig[iy]) + 0.5 * r_ * r_ / (sig[iy] * sig[iy]);
                        if (b.res && full()) b.res[(size_t)n * nt * NY + (size_t)it * NY + iy] = res[iy];
                        chi2 += res[iy] * res[iy];
                    }
                }
            } else {
                const int ip = S.ip_of_col[cc];
                const int il = cc - 1;  // index in plist
                const double pc = S.pcoef[cc];
                double sy[NY], ssig_tot[NY], dsig[NY];
                M::sy(ip, t, e.y, S.p, S.k, S.h, e.spl, e.sspl[c], xc[c], sy);
                M::ssigmay(ip, S.p, S.k, y, sy, ssig_tot);
                M::dsigmaydp(ip, S.p, S.k, y, dsig);
                if (b.sx && full())
                    AB_UNROLL
                    for (int r = 0; r < NX; ++r)
                        b.sx[(((size_t)n * nt + it) * nplist + il) * NX + r] = chain(xc[c][r], pc);
                AB_UNROLL
                for (int iy = 0; iy < NY; ++iy) {
                    // sigma given by the data overrides the model sigma: dsigmaydp := 0 (src/model.cpp:2199-2209)
                    const double dsp = sig_set[iy] ? 0.0 : dsig[iy];
                    const double ssig = dsp + (ssig_tot[iy] - dsig[iy]);
                    const size_t idx = (((size_t)n * nt + it) * nplist + il) * NY + iy;
                    if (b.sy && full()) b.sy[idx] = chain(sy[iy], pc);
                    if (b.ssigmay && full()) b.ssigmay[idx] = chain(ssig, pc);
                    if (my_set[iy]) {
                        const double r_ = y[iy] - my[iy];
                        const double s2 = sig[iy] * sig[iy];
                        const double dJydy = r_ / s2;
                        const double dJyds = 1.0 / sig[iy] - r_ * r_ / (s2 * sig[iy]);
                        sllh[c] -= dJydy * sy[iy] + dJyds * dsp;
                        const double sr = (-sy[iy] - ssig * res[iy]) / sig[iy];
                        if (b.sres && full())
                            b.sres[((size_t)n * nt * NY + (size_t)it * NY + iy) * nplist + il] = chain(sr, pc);
                    }
                }
            }
        }
        // FIM += sres sres^T per set measurement (reference src/rdata.cpp:1106-1205)
        if (b.FIM && full() && nplist > 0) {
            AB_UNROLL
            for (int iy = 0; iy < NY; ++iy) {
                if (!my_set[iy]) continue;
                double mine[C];
                g.sync();
                AB_UNROLL
                for (int c = 0; c < C; ++c) {
                    mine[c] = 0.0;
                    if (!bdf.active(c) || bdf.col(c) == 0) continue;
                    const int cc = bdf.col(c);
                    double sy[NY], ssig_tot[NY], dsig[NY];
                    M::sy(S.ip_of_col[cc], t, e.y, S.p, S.k, S.h, e.spl, e.sspl[c], xc[c], sy);
                    M::ssigmay(S.ip_of_col[cc], S.p, S.k, y, sy, ssig_tot);
                    M::dsigmaydp(S.ip_of_col[cc], S.p, S.k, y, dsig);
                    const double dsp = sig_set[iy] ? 0.0 : dsig[iy];
                    const double ssig = dsp + (ssig_tot[iy] - dsig[iy]);
                    mine[c] = chain((-sy[iy] - ssig * res[iy]) / sig[iy], S.pcoef[cc]);
                    S.scratch[cc] = mine[c];
                }
                g.sync();
                AB_UNROLL
                for (int c = 0; c < C; ++c) {
                    if (!bdf.active(c) || bdf.col(c) == 0) continue;
                    const int il = bdf.col(c) - 1;
                    double* row = b.FIM + ((size_t)n * nplist + il) * nplist;
                    for (int jl = 0; jl < nplist; ++jl) row[jl] += mine[c] * S.scratch[jl + 1];
                }
            }
        }
        if (g.lane == 0) {
            const size_t idx = (size_t)n * nt + it;
            if (b.numsteps) b.numsteps[idx] = bdf.nst;
            if (b.numrhsevals) b.numrhsevals[idx] = bdf.nfe;
            if (b.numerrtestfails) b.numerrtestfails[idx] = bdf.netf;
            if (b.numnonlinsolvconvfails) b.numnonlinsolvconvfails[idx] = bdf.ncfn;
            if (b.order) b.order[idx] = bdf.qu;
        }
    }

    // ReturnData::invalidate (reference src/rdata.cpp:684-735) with nan_ = NaN.  A failed pre-equilibration instead
    // leaves every per-timepoint field and the objective at its initial value 0 (process_forward_problem returns
    // early, reference src/rdata.cpp:320-323, :105-190): nan_ = 0.
    AB_HD void invalidate(int it_start, const double nan_) {
        const int nt = b.nt;
        llh = nan_;
        chi2 = nan_;
        AB_UNROLL
        for (int c = 0; c < C; ++c) sllh[c] = nan_;
        if (!full()) return;
        AB_UNROLL
        for (int c = 0; c < C; ++c) {
            if (!bdf.active(c)) continue;
            const int cc = bdf.col(c);
            for (int it = it_start; it < nt; ++it) {
                if (cc == 0) {
                    for (int r = 0; r < NX; ++r)
                        if (b.x) b.x[((size_t)n * nt + it) * NX + r] = nan_;
                    for (int iy = 0; iy < NY; ++iy) {
                        if (b.y) b.y[((size_t)n * nt + it) * NY + iy] = nan_;
                        if (b.sigmay && nan_ == 0.0) b.sigmay[((size_t)n * nt + it) * NY + iy] = nan_;
                        if (b.res) b.res[(size_t)n * nt * NY + (size_t)it * NY + iy] = nan_;
                    }
                } else {
                    const int il = cc - 1;
                    for (int r = 0; r < NX; ++r)
                        if (b.sx) b.sx[(((size_t)n * nt + it) * nplist + il) * NX + r] = nan_;
                    for (int iy = 0; iy < NY; ++iy) {
                        if (b.sy) b.sy[(((size_t)n * nt + it) * nplist + il) * NY + iy] = nan_;
                        if (b.ssigmay && nan_ == 0.0) b.ssigmay[(((size_t)n * nt + it) * nplist + il) * NY + iy] = nan_;
                        if (b.sres)
                            b.sres[((size_t)n * nt * NY + (size_t)it * NY + iy) * nplist + il] = nan_;
                    }
                }
            }
            if (cc > 0 && b.FIM)
                for (int jl = 0; jl < nplist; ++jl) b.FIM[((size_t)n * nplist + (cc - 1)) * nplist + jl] = nan_;
        }
    }

    // x0 / sx0 outputs: the state the main simulation starts from (reference src/rdata.cpp:326-334)
    AB_HD void write_initial_state(const bool valid = true) {
        AB_UNROLL
        for (int c = 0; c < C; ++c) {
            if (!bdf.active(c)) continue;
            const int cc = bdf.col(c);
            if (cc == 0) {
                AB_UNROLL
                for (int r = 0; r < NX; ++r) S.xfin[r] = S.zn[0][r][0];
            }
            if (!full()) continue;
            AB_UNROLL
            for (int r = 0; r < NX; ++r) {
                const double v = valid ? S.zn[0][r][cc] : ab_nan();
                if (cc == 0) { if (b.x0) b.x0[(size_t)n * NX + r] = v; }
                else if (b.sx0) b.sx0[((size_t)n * nplist + (cc - 1)) * NX + r] = chain(v, S.pcoef[cc]);
            }
        }
    }

    // x_ss / sx_ss outputs (ReturnData::process_pre_equilibration, reference src/rdata.cpp:261-286, chain rule :874)
    AB_HD void write_steady_state(const bool valid = true) {
        AB_UNROLL
        for (int c = 0; c < C; ++c) {
            if (!bdf.active(c) || !full()) continue;
            const int cc = bdf.col(c);
            AB_UNROLL
            for (int r = 0; r < NX; ++r) {
                const double v = valid ? S.zn[0][r][cc] : ab_nan();
                if (cc == 0) { if (b.x_ss) b.x_ss[(size_t)n * NX + r] = v; }
                else if (b.sx_ss) b.sx_ss[((size_t)n * nplist + (cc - 1)) * NX + r] = chain(v, S.pcoef[cc]);
            }
        }
    }

    // SteadyStateProblem::run for pre-equilibration (reference src/forwardproblem.cpp:812-870, :1107-1206), part 1:
    // Newton's method (stage 0 or 2). Returns true if the steady state was found.
    AB_HD bool preeq_newton(int stage, int (&pe_status)[3], int (&pe_steps)[3], double& pe_wrms) {
        const int rc = cold_newton<M, G, C>(S, o, INFINITY, &pe_steps[stage], &pe_wrms);
        pe_status[stage] = rc == AB_SUCCESS ? SS_SUCCESS
                           : rc == AB_TOO_MUCH_WORK ? SS_FAILED_CONVERGENCE
                           : rc == AB_SINGULAR_JACOBIAN ? SS_FAILED_FACTORIZATION
                           : rc == AB_DAMPING_FACTOR_ERROR ? SS_FAILED_DAMPING : SS_FAILED;
        return rc == AB_SUCCESS;
    }

    AB_HD void run() {
        auto& g = bdf.g;
        const int nt = b.nt;
        llh = 0.0;
        chi2 = 0.0;
        AB_UNROLL
        for (int c = 0; c < C; ++c) sllh[c] = 0.0;
        const bool has_preeq = PREEQ && b.k_preeq != nullptr;
        load_condition(has_preeq);
        double t = o.t0;
        int status = AB_SUCCESS;
        int it = 0;
        // ---- pre-equilibration state (SteadyStateProblem, reference src/forwardproblem.cpp:755-1223) ----
        const int ncols_main = bdf.ncols;
        int pe_status[3] = {SS_NOT_RUN, SS_NOT_RUN, SS_NOT_RUN}, pe_steps[3] = {0, 0, 0};
        double pe_wrms = ab_nan(), pe_t = ab_nan();
        // phase: 0 = main simulation, 1 = steady-state simulation, 2 = steady state found (finish), 3 = simulation failed
        int phase = 0;
        bool newton_on = false, sim_on = false, preeq_failed = false;
        if (PREEQ && has_preeq) {
            const bool sens = ncols_main > 1;
            newton_on = !(o.ss_mode == SS_INTEGRATION_ONLY || o.newton_maxsteps == 0 ||
                          (sens && o.ss_sensi_mode == SS_INTEGRATION_ONLY));
            sim_on = o.ss_mode != SS_NEWTON_ONLY;
            phase = 3;
            if (newton_on && preeq_newton(0, pe_status, pe_steps, pe_wrms)) phase = 2;
            else if (sim_on) {
                // find_steady_state_by_simulation: restart from the initial state with a fresh solver; sensitivities
                // are integrated along unless they come from the Newton solve afterwards (:872-920, :1023-1033)
                if (newton_on && g.lane == 0) {
                    double v[NX];
                    M::x0(S.p, S.k, v);
                    for (int r = 0; r < NX; ++r) S.zn[0][r][0] = v[r];
                }
                g.sync();
                bdf.ncols = (sens && o.ss_sensi_mode != SS_NEWTON_ONLY) ? ncols_main : 1;
                bdf.init(o.t0);
                phase = 1;
            }
        }
        if (phase == 0) {
            write_initial_state();
            bdf.init(o.t0);
        }
        // ---- condition driver: the one call site of the integrator serves both phases ----
        for (;;) {
            double tout = 0.0;
            bool one_step = false;
            if (PREEQ && phase != 0) {
                if (phase == 1) {
                    // EventHandlingSimulator::run_steady_state (reference src/forwardproblem.cpp:153-243)
                    const bool check_sens = bdf.ncols > 1 && o.sensi_ss_check != 0;
                    if (cold_ss_converged<M, G, C>(S, o, bdf.ncols, t, check_sens, &pe_wrms)) {
                        pe_status[1] = SS_SUCCESS;
                        pe_t = t;
                        phase = 2;
                    } else if (pe_steps[1] >= o.maxsteps) {
                        pe_status[1] = SS_FAILED_CONVERGENCE;
                        phase = 3;
                    } else {
                        pe_steps[1]++;
                        tout = ab_max(t, 1.0) * 10.0;
                        one_step = true;
                        if (!isfinite(tout)) {
                            pe_status[1] = SS_FAILED_TOO_LONG;
                            phase = 3;
                        }
                    }
                }
                if (phase == 3) {
                    // simulation did not work (or was not allowed): retry Newton from the last state (:1141-1150)
                    if (newton_on && sim_on && pe_status[1] != SS_NOT_RUN && preeq_newton(2, pe_status, pe_steps, pe_wrms)) {
                        phase = 2;
                    } else {
                        status = AB_ERROR;   // handle_steady_state_failure -> AmiException (src/amici.cpp:155-170)
                        preeq_failed = true;
                        break;
                    }
                }
                if (phase == 2) {
                    // steady-state sensitivities by one linear solve per parameter unless the simulation carried them
                    const bool sens = ncols_main > 1;
                    const bool fwd_done = pe_status[1] == SS_SUCCESS && o.ss_sensi_mode != SS_NEWTON_ONLY;
                    const bool started_in_ss = pe_status[0] == SS_SUCCESS && pe_steps[0] == 0;
                    if (sens && !fwd_done && !started_in_ss) {
                        const double tss = pe_status[1] == SS_SUCCESS ? t : INFINITY;
                        if (!cold_newton_sensis<M, G, C>(S, ncols_main, tss)) {
                            status = AB_ERROR;
                            preeq_failed = true;
                            break;
                        }
                    }
                    bdf.ncols = ncols_main;
                    write_steady_state();
                    // back to the simulation condition (ConditionContext restore) and on to the main simulation
                    g.sync();
                    if (g.lane == 0)
                        for (int i = 0; i < M::NK; ++i) S.k[i] = b.k[(size_t)n * M::NK + i];
                    g.sync();
                    write_initial_state();
                    bdf.init(o.t0);
                    t = o.t0;
                    phase = 0;
                }
            }
            if (!PREEQ || phase == 0) {
                if (it >= nt) break;
                tout = b.ts[it];
                if (!(tout > o.t0) || !(t < tout) || next_t_too_close(t, tout)) {
                    datapoint(it);
                    ++it;
                    continue;
                }
            }
            const int flag = bdf.solve_to(tout, one_step);
            if (flag >= 0 || (PREEQ && phase == 1)) t = bdf.tretlast;   // on failure the reference keeps the last returned solution
            if (flag < 0) {
                if (PREEQ && phase == 1) {
                    pe_status[1] = flag == AB_TOO_MUCH_WORK ? SS_FAILED_CONVERGENCE : SS_FAILED;
                    phase = 3;
                    continue;
                }
                status = flag;
                break;
            }
            if (!PREEQ || phase == 0) {
                g.sync();
                if (g.lane == 0) {
                    AB_UNROLL
                    for (int r = 0; r < NX; ++r) S.xfin[r] = S.zn[0][r][0];
                }
            }
        }
        if (PREEQ && preeq_failed) {
            // the steady-state problem's last state is reported as x_ss / sx_ss and used for xdot / J; the forward
            // problem never started: x0 / sx0 stay NaN (reference src/rdata.cpp:233-252, :261-273, :320-323)
            bdf.ncols = ncols_main;
            write_steady_state();
            write_initial_state(false);
            t = INFINITY;
        } else if (!has_preeq) {
            write_steady_state(false);
        }
        if (PREEQ && has_preeq && g.lane == 0) {
            for (int i = 0; i < 3; ++i) {
                if (b.preeq_status) b.preeq_status[(size_t)n * 3 + i] = pe_status[i];
                if (b.preeq_numsteps) b.preeq_numsteps[(size_t)n * 3 + i] = pe_steps[i];
            }
            if (b.preeq_wrms) b.preeq_wrms[n] = pe_wrms;
            if (b.preeq_t) b.preeq_t[n] = pe_t;
        }
        if (status != AB_SUCCESS) invalidate(it, (PREEQ && preeq_failed) ? 0.0 : ab_nan());
        // xdot and J at the final state = the last solution the solver returned (ReturnData::store_jacobian_and_derivative,
        // reference rdata.h:707-728; FinalStateStorer, src/forwardproblem.cpp:245-256)
        if (full() && (b.xdot || b.J)) {
            g.sync();
            if (g.lane == 0) {
                double xf[NX], f[NX], Jv[M::NNZ_J > 0 ? M::NNZ_J : 1], spl[M::NSPL > 0 ? M::NSPL : 1];
                AB_UNROLL
                for (int r = 0; r < NX; ++r) xf[r] = S.xfin[r];
                spline_values_at<M, G, C>(S, t, spl);
                M::xdot(t, xf, S.p, S.k, S.h, spl, f);
                M::jac(t, xf, S.p, S.k, S.h, spl, Jv);
                if (b.xdot)
                    for (int r = 0; r < NX; ++r) b.xdot[(size_t)n * NX + r] = f[r];
                if (b.J) M::jac_to_dense(Jv, b.J + (size_t)n * NX * NX);
            }
        }
        AB_UNROLL
        for (int c = 0; c < C; ++c) {
            if (!bdf.active(c)) continue;
            const int cc = bdf.col(c);
            if (cc == 0) {
                if (b.status) b.status[n] = status;
                if (b.llh) b.llh[n] = llh;
                if (b.chi2) b.chi2[n] = chi2;
            } else if (b.sllh) {
                b.sllh[(size_t)n * nplist + (cc - 1)] = chain(sllh[c], S.pcoef[cc]);
            }
        }
    }
};

}  // namespace amici_b200
#endif
