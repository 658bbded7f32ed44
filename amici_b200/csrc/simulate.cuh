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
    using Solver = Bdf<M, G, C, PREEQ>;
    static constexpr int NX = M::NX, NY = M::NY, NCOLS = Solver::NCOLS;
    using Shared = typename Solver::Shared;
    // Output columns: 0 = state, 1 + il = entry il of plist.  Normally they are the integrated columns; with a COMPACT layout
    // (bdf_core.cuh Layout) the parameters whose state sensitivity is identically zero have no integrated column, and a
    // lane serves the output columns lane, lane + G, ... (C_OUT slots) with sx = 0 for those.
    static constexpr bool COMPACT = Layout<M, G, C>::COMPACT;
    static_assert(!(COMPACT && M::NE > 0), "compact layouts are for event-free models");
    static constexpr int C_OUT = COMPACT ? (M::NP + 1 + G - 1) / G : C;
    static constexpr int NSPLV = M::NSPL > 0 ? M::NSPL : 1;

    Shared& S;
    const SolverOptions& o;
    const BatchView& b;
    Solver bdf;
    int n;       // trajectory index
    int nplist;
    double llh, chi2;
    double sllh[C_OUT];
    // event bookkeeping (FwdSimWorkspace, reference include/amici/forwardproblem.h:109-175): group-uniform registers
    static constexpr int NEV = M::NE > 0 ? M::NE : 1, NZV = M::NZ > 0 ? M::NZ : 1;
    int roots_found[NEV], nroots[NEV];
    double rootvals[NEV], rval_tmp[NEV];
    double ev_x_old[NX], ev_xdot_old[NX];
    double stau[C];
    double tlastroot, t_trig_done;
    bool recording = true;   // false during pre-equilibration and presimulation: events act on the state, nothing is stored (edata == nullptr / PeriodResult of another simulator)
    bool clock_started = false;   // Solver::set_max_time: the clock starts with the first phase of the condition

    AB_HD Simulation(Shared& s, const SolverOptions& opt, const BatchView& bv, int n_)
        : S(s), o(opt), b(bv), bdf(s, opt, integrated_columns(opt, bv)), n(n_),
          nplist(opt.sensi >= 1 ? bv.nplist : 0) {}

    // 1 + the number of plist entries that get an integrated sensitivity column
    AB_HD static int integrated_columns(const SolverOptions& opt, const BatchView& bv) {
        if (opt.sensi < 1) return 1;
        if (!COMPACT) return 1 + bv.nplist;
        unsigned long long dyn = 0ull;   // one integrated column per DISTINCT parameter with a non-zero state sensitivity
        for (int il = 0; il < bv.nplist; ++il) dyn |= 1ull << bv.plist[il];
        dyn &= ~M::SX_STATIC_MASK;
        int nc = 1;
        for (; dyn; dyn &= dyn - 1) ++nc;
        return nc;
    }

    AB_HD bool full() const { return o.reporting != 2; }   // RDataReporting::likelihood = 2: objective only; the other modes are selected by which output pointers are given
    // output slot j of this lane
    AB_HD int out_col(const int j) const { return bdf.g.lane + G * j; }
    AB_HD bool out_active(const int j) const { return out_col(j) <= nplist; }
    AB_HD int int_col(const int oc) const { return COMPACT ? S.col_of_out[oc] : oc; }

    // ConditionContext + Model::initialize (reference src/edata.cpp:438-556, src/model.cpp:251-278):
    // unscale parameters (reference include/amici/misc.h:164-195), x0, sx0, spline tables.
    AB_HD void load_condition(const double* ksrc) {
        auto& g = bdf.g;
        g.sync();
        if (g.lane == 0) {
            for (int i = 0; i < M::NP; ++i) {
                const double v = b.p[(size_t)n * M::NP + i];
                const int sc = b.pscale ? b.pscale[i] : AB_PSCALE_NONE;
                S.p[i] = sc == AB_PSCALE_LOG10 ? pow(10.0, v) : sc == AB_PSCALE_LN ? exp(v) : v;
            }
            // ConditionContext with FixedParameterContext::simulation / preequilibration / presimulation (reference src/edata.cpp:438-556)
            for (int i = 0; i < M::NK; ++i) S.k[i] = ksrc[(size_t)n * M::NK + i];
            M::derived(S.p, S.k);
            for (int i = 0; i < M::NH; ++i) S.h[i] = 0.0;
            S.xpos_mask = o.x_nonneg;
#if AB_FULL && defined(__CUDA_ARCH__)
            if (o.maxtime > 0.0 && !clock_started) {   // the guard covers the whole simulation of this condition (all phases)
                long long now;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                S.t_start_ns = now;
            }
#endif
            clock_started = true;
            for (int cc = 0; cc < NCOLS; ++cc) {
                S.ip_of_col[cc] = -1;
                S.il_of_col[cc] = -1;
                S.pcoef[cc] = 1.0;
            }
            S.col_of_out[0] = 0;
            S.pcoef_out[0] = 1.0;
            int cc = 1;
            for (int il = 0; il < nplist; ++il) {
                const int ip = b.plist[il];
                const int sc = b.pscale ? b.pscale[ip] : AB_PSCALE_NONE;
                // pcoefficient of ReturnData::apply_chain_rule_factor_to_simulation_results (rdata.cpp:783-796)
                const double pc = sc == AB_PSCALE_LOG10 ? S.p[ip] * log(10.0) : sc == AB_PSCALE_LN ? S.p[ip] : 1.0;
                S.pcoef_out[il + 1] = pc;
                int have = -1;   // COMPACT: a parameter listed twice shares its column (identical vectors, same max-norms)
                if (COMPACT)
                    for (int c2 = 1; c2 < cc; ++c2)
                        if (S.ip_of_col[c2] == ip) have = c2;
                if (COMPACT && (M::SX_STATIC_MASK >> ip & 1ull)) {
                    S.col_of_out[il + 1] = -1;   // sx == 0 identically: no integrated column
                } else if (have > 0) {
                    S.col_of_out[il + 1] = have;
                } else {
                    S.col_of_out[il + 1] = cc;
                    S.ip_of_col[cc] = ip;
                    S.il_of_col[cc] = il;
                    S.pcoef[cc] = pc;
                    ++cc;
                }
            }
        }
        g.sync();
        AB_UNROLL
        for (int c = 0; c < C; ++c) {
            if (!bdf.active(c)) continue;
            const int cc = bdf.col(c);
            double v[NX];
            initial_column(cc, v);
            AB_UNROLL
            for (int r = 0; r < NX; ++r) S.zn[0][r][cc] = v[r];
            AB_UNROLL
            for (int is = 0; is < M::NSPL; ++is)
                spline_build_column<M, NCOLS>(S.spl, is, cc, cc == 0 ? -1 : S.ip_of_col[cc], S.p, S.k);
        }
        g.sync();
    }

    // Model::fx0 / fsx0 with the ExpData overrides (reference src/model.cpp:1907-1966; ConditionContext src/edata.cpp:472-490):
    // a given sx0 is on the parameter scale of pscale and is un-scaled with the chain-rule factor (src/model.cpp:986-1024)
    AB_HD void initial_column(const int cc, double (&v)[NX]) const {
        if (cc == 0) {
            M::x0(S.p, S.k, v);
            if (b.x0_in)
                AB_UNROLL
                for (int r = 0; r < NX; ++r) v[r] = b.x0_in[(size_t)n * NX + r];
        } else {
            M::sx0(S.ip_of_col[cc], S.p, S.k, v);
            if (b.sx0_in)
                AB_UNROLL
                for (int r = 0; r < NX; ++r) v[r] = b.sx0_in[((size_t)n * nplist + S.il_of_col[cc]) * NX + r] / S.pcoef[cc];
        }
    }

    // Model::reinitialize (reference src/model.cpp:280-295; generated x0_fixedParameters.cpp / sx0_fixedParameters.cpp): the
    // selected states whose initial value is a function of the fixed parameters are reset to it, their sensitivities to 0
    AB_HD void reinitialize_states(const unsigned long long idxs) {
        const unsigned long long m = idxs & M::X0_FIXED_MASK;
        if (m == 0) return;
        auto& g = bdf.g;
        g.sync();
        AB_UNROLL
        for (int c = 0; c < C; ++c) {
            if (!bdf.active(c)) continue;
            const int cc = bdf.col(c);
            double v[NX];
            AB_UNROLL
            for (int r = 0; r < NX; ++r) v[r] = 0.0;
            if (cc == 0) M::x0(S.p, S.k, v);
            AB_UNROLL
            for (int r = 0; r < NX; ++r)
                if (m >> r & 1) S.zn[0][r][cc] = v[r];
        }
        g.sync();
    }

    AB_HD double chain(double v, double pc) const { return v != 0.0 ? v * pc : v; }

    // ReturnData::get_data_output + get_data_sensis_fsa for timepoint `it` from the solver state
    // (reference src/rdata.cpp:368-432; observables/likelihood: src/model.cpp:1112-1259, :2140-2413;
    //  residuals/FIM: src/rdata.cpp:985-1205).  Gaussian noise model (generated Jy.cpp/dJydy.cpp/dJydsigma.cpp).
    AB_HD void datapoint(int it, const bool posteq_point = false) {
        auto& g = bdf.g;
        const int nt = b.nt;
        const double t = b.ts[it];
        double xs[NX], spl[NSPLV];
        SplineLoc loc[NSPLV];
        g.sync();
        AB_UNROLL
        for (int r = 0; r < NX; ++r) xs[r] = S.zn[0][r][0];
        x_pos<M>(S, xs);   // observables and the reported state see the non-negative state (Model::fx_rdata, reference src/model.cpp:1930)
        AB_UNROLL
        for (int is = 0; is < M::NSPL; ++is) {
            loc[is] = spline_locate<M>(is, t);
            const double v = spline_eval_at<M, NCOLS>(S.spl, is, 0, loc[is], true);
            spl[is] = spline_is_log<M>(is) ? exp(v) : v;
        }
        double y[NY], sig[NY], my[NY];
        bool my_set[NY], sig_set[NY];
        M::y(t, xs, S.p, S.k, S.h, spl, y);
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
        double srv[C_OUT][NY];   // this lane's rows of sres (chain rule applied), for the FIM
        AB_UNROLL
        for (int j = 0; j < C_OUT; ++j) {
            AB_UNROLL
            for (int iy = 0; iy < NY; ++iy) srv[j][iy] = 0.0;
            if (!out_active(j)) continue;
            const int oc = out_col(j);
            if (oc == 0) {
                if (b.x && full())
                    AB_UNROLL
                    for (int r = 0; r < NX; ++r) b.x[((size_t)n * nt + it) * NX + r] = xs[r];
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
                const int il = oc - 1;  // index in plist
                const int cc = int_col(oc);   // integrated column, or -1: sx == 0
                const int ip = COMPACT ? b.plist[il] : S.ip_of_col[oc];
                const double pc = S.pcoef_out[oc];
                double sxv[NX], sspl[NSPLV], sy[NY], ssig_tot[NY];
                AB_UNROLL
                for (int r = 0; r < NX; ++r) sxv[r] = (!COMPACT || cc >= 0) ? S.zn[0][r][cc >= 0 ? cc : 0] : 0.0;
                AB_UNROLL
                for (int is = 0; is < M::NSPL; ++is) {
                    sspl[is] = 0.0;
                    if (COMPACT && cc < 0) continue;
                    const double sv = spline_eval_at<M, NCOLS>(S.spl, is, cc, loc[is], false);
                    sspl[is] = spline_is_log<M>(is) ? sv * spl[is] : sv;
                }
                M::sy(ip, t, xs, S.p, S.k, S.h, spl, sspl, sxv, sy);
                M::ssigmay(ip, S.p, S.k, y, sy, ssig_tot);   // dsigmay/dp + dsigmay/dy * sy
                if (b.sx && full())
                    AB_UNROLL
                    for (int r = 0; r < NX; ++r)
                        b.sx[(((size_t)n * nt + it) * nplist + il) * NX + r] = chain(sxv[r], pc);
                AB_UNROLL
                for (int iy = 0; iy < NY; ++iy) {
                    // sigma given by the data overrides the model sigma: dsigmaydp := 0 and dsigmaydy := 0
                    // (reference src/model.cpp:2199-2209, :2231-2238)
                    const double ssig = sig_set[iy] ? 0.0 : ssig_tot[iy];
                    const size_t idx = (((size_t)n * nt + it) * nplist + il) * NY + iy;
                    if (b.sy && full()) b.sy[idx] = chain(sy[iy], pc);
                    if (b.ssigmay && full()) b.ssigmay[idx] = chain(ssig, pc);
                    if (my_set[iy]) {
                        const double r_ = y[iy] - my[iy];
                        const double s2 = sig[iy] * sig[iy];
                        const double dJydy = r_ / s2;
                        const double dJyds = 1.0 / sig[iy] - r_ * r_ / (s2 * sig[iy]);
                        // dJydy of the reference carries dJydsigma * dsigmaydy (src/model.cpp:2278-2290): together
                        // -(dJydy * sy + dJydsigma * (dsigmaydp + dsigmaydy * sy))
                        sllh[j] -= dJydy * sy[iy] + dJyds * ssig;
                        const double sr = (-sy[iy] - ssig * res[iy]) / sig[iy];
                        srv[j][iy] = chain(sr, pc);
                        if (b.sres && full())
                            b.sres[((size_t)n * nt * NY + (size_t)it * NY + iy) * nplist + il] = srv[j][iy];
                    }
                }
            }
        }
        // FIM += sres sres^T per set measurement (reference src/rdata.cpp:1106-1205)
        if (b.FIM && nplist > 0) {   // every reporting mode with sensitivities has the FIM (reference src/rdata.cpp:104-119)
            AB_UNROLL
            for (int iy = 0; iy < NY; ++iy) {
                if (!my_set[iy]) continue;
                g.sync();
                AB_UNROLL
                for (int j = 0; j < C_OUT; ++j)
                    if (out_active(j) && out_col(j) > 0) S.scratch[out_col(j)] = srv[j][iy];
                g.sync();
                AB_UNROLL
                for (int j = 0; j < C_OUT; ++j) {
                    if (!out_active(j) || out_col(j) == 0) continue;
                    double* row = b.FIM + ((size_t)n * nplist + (out_col(j) - 1)) * nplist;
                    for (int jl = 0; jl < nplist; ++jl) row[jl] += srv[j][iy] * S.scratch[jl + 1];
                }
            }
        }
        if (g.lane == 0) {
            const size_t idx = (size_t)n * nt + it;
            // the per-timepoint solver diagnostics are not written for t = inf points (they stay 0, src/rdata.cpp:288-300)
            if (b.numsteps) b.numsteps[idx] = posteq_point ? 0 : bdf.nst;
            if (b.numrhsevals) b.numrhsevals[idx] = posteq_point ? 0 : bdf.nfe;
            if (b.numerrtestfails) b.numerrtestfails[idx] = posteq_point ? 0 : bdf.netf;
            if (b.numnonlinsolvconvfails) b.numnonlinsolvconvfails[idx] = posteq_point ? 0 : bdf.ncfn;
            if (b.order) b.order[idx] = posteq_point ? 0 : bdf.qu;
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
        for (int j = 0; j < C_OUT; ++j) sllh[j] = nan_;
        if (!full()) return;
        AB_UNROLL
        for (int j = 0; j < C_OUT; ++j) {
            if (!out_active(j)) continue;
            const int oc = out_col(j);
            for (int it = it_start; it < nt; ++it) {
                if (oc == 0) {
                    for (int r = 0; r < NX; ++r)
                        if (b.x) b.x[((size_t)n * nt + it) * NX + r] = nan_;
                    for (int iy = 0; iy < NY; ++iy) {
                        if (b.y) b.y[((size_t)n * nt + it) * NY + iy] = nan_;
                        if (b.sigmay && nan_ == 0.0) b.sigmay[((size_t)n * nt + it) * NY + iy] = nan_;
                        if (b.res) b.res[(size_t)n * nt * NY + (size_t)it * NY + iy] = nan_;
                    }
                } else {
                    const int il = oc - 1;
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
            if (oc > 0 && b.FIM)
                for (int jl = 0; jl < nplist; ++jl) b.FIM[((size_t)n * nplist + (oc - 1)) * nplist + jl] = nan_;
        }
    }

    // x0 / sx0 outputs: the state the main simulation starts from (reference src/rdata.cpp:326-334)
    AB_HD void write_initial_state(const bool valid = true) {
        auto& g = bdf.g;
        g.sync();
        if (g.lane == 0) {
            AB_UNROLL
            for (int r = 0; r < NX; ++r) S.xfin[r] = S.zn[0][r][0];
        }
        if (!full()) return;
        write_state_columns(b.x0, b.sx0, valid);
    }

    // x_ss / sx_ss outputs (ReturnData::process_pre_equilibration, reference src/rdata.cpp:261-286, chain rule :874)
    AB_HD void write_steady_state(const bool valid = true) {
        if (!full()) return;
        write_state_columns(b.x_ss, b.sx_ss, valid);
    }

    // zn[0] of every output column -> xout[n][nx], sxout[n][nplist][nx] (0 for a column that is not integrated)
    AB_HD void write_state_columns(double* xout, double* sxout, const bool valid) {
        AB_UNROLL
        for (int j = 0; j < C_OUT; ++j) {
            if (!out_active(j)) continue;
            const int oc = out_col(j), cc = int_col(oc);
            AB_UNROLL
            for (int r = 0; r < NX; ++r) {
                const double v = !valid ? ab_nan() : (!COMPACT || cc >= 0) ? S.zn[0][r][cc >= 0 ? cc : 0] : 0.0;
                if (oc == 0) { if (xout) xout[(size_t)n * NX + r] = v; }
                else if (sxout) sxout[((size_t)n * nplist + (oc - 1)) * NX + r] = chain(v, S.pcoef_out[oc]);
            }
        }
    }

    // ---------------------------------------------------------------------------------------------
    // events (EventHandlingSimulator, reference src/forwardproblem.cpp:389-735; Model event functions,
    // reference src/model.cpp:407-472, :1464-1534, :1588-1592; event outputs src/rdata.cpp:434-512)
    // ---------------------------------------------------------------------------------------------
    AB_HD void state_col(double (&x)[NX]) const {
        AB_UNROLL
        for (int r = 0; r < NX; ++r) x[r] = S.zn[0][r][0];
        x_pos<M>(S, x);   // event functions see it too; add_state_event_update writes it back (reference src/model.cpp:1473-1503)
    }
    AB_HD void eval_roots(double t, const double (&x)[NX], double (&rv)[NEV]) const {
        double spl[M::NSPL > 0 ? M::NSPL : 1];
        spline_values_at<M, G, C>(S, t, spl);
        M::root(t, x, S.p, S.k, S.h, spl, rv);
    }
    AB_HD void eval_xdot(double t, const double (&x)[NX], double (&f)[NX]) const {
        double spl[M::NSPL > 0 ? M::NSPL : 1];
        spline_values_at<M, G, C>(S, t, spl);
        M::xdot(t, x, S.p, S.k, S.h, spl, f);
    }
    // h += roots_found (Model::update_heaviside)
    AB_HD void update_heaviside() {
        auto& g = bdf.g;
        g.sync();
        if (g.lane == 0) {
            AB_UNROLL
            for (int ie = 0; ie < M::NE; ++ie) S.h[ie] += roots_found[ie];
        }
        g.sync();
    }
    // Model::initialize_events / reinit_events (reference src/model.cpp:407-442).  from_model_state: the "old" Heaviside
    // vector is the model's current one (re-initialisation after a pre-equilibration / presimulation, reference
    // src/forwardproblem.cpp:305-311, :353-360) instead of the events' initial values; the workspace's root values
    // (FwdSimWorkspace::rootvals) are kept then.
    AB_HD void initialize_events(double t, const bool from_model_state = false) {
        auto& g = bdf.g;
        double x[NX], rv[NEV];
        g.sync();
        state_col(x);
        eval_roots(t, x, rv);
        AB_UNROLL
        for (int ie = 0; ie < M::NE; ++ie) {
            const bool h_old = from_model_state ? (S.h[ie] > 0.0) : M::event_initial_value(ie);
            roots_found[ie] = (!(rv[ie] < 0.0) && !h_old) ? 1 : 0;   // only false -> true triggers an event
            nroots[ie] = 0;
            if (!from_model_state) rootvals[ie] = rval_tmp[ie] = 0.0;
        }
        g.sync();
        if (g.lane == 0) {
            AB_UNROLL
            for (int ie = 0; ie < M::NE; ++ie) S.h[ie] = (rv[ie] < 0.0) ? 0.0 : 1.0;
        }
        g.sync();
        AB_UNROLL
        for (int c = 0; c < C; ++c) stau[c] = 0.0;
        tlastroot = t;
        t_trig_done = t;
    }
    // next explicit trigger time in (t_trig_done, t_end] (Model::get_trigger_timepoints, filtered as in
    // reference src/forwardproblem.cpp:70-79)
    AB_HD double next_trigger(double t_end) const {
        double best = INFINITY;
        AB_UNROLL
        for (int ie = M::NE_SOLVER; ie < M::NE; ++ie) {
            const double te = M::explicit_time(ie, S.p, S.k);
            if (te > t_trig_done && te <= t_end && te < best) best = te;
        }
        return best;
    }

    // ReturnData::get_event_output + get_event_sensis_fsa for the events flagged in roots_found, from the current
    // (pre-event) state; EventHandlingSimulator::store_event's counter logic (reference src/forwardproblem.cpp:524-574)
    AB_HD void store_event(double t) {
        auto& g = bdf.g;
        const int nt = b.nt, nme = o.nmaxevent;
        const bool is_last = nt > 0 && t == b.ts[nt - 1];
        double x[NX];
        g.sync();
        state_col(x);
        for (;;) {
        if (is_last) {
            eval_roots(t, x, rootvals);
            AB_UNROLL
            for (int ie = 0; ie < M::NE; ++ie) roots_found[ie] = nroots[ie] < nme ? 1 : 0;
        }
        double spl[M::NSPL > 0 ? M::NSPL : 1];
        spline_values_at<M, G, C>(S, t, spl);
        AB_UNROLL
        for (int ie = 0; ie < M::NE; ++ie) {
            if (nroots[ie] >= nme) continue;
            if (roots_found[ie] != 1) continue;
            const int slot = nroots[ie];
            const size_t zrow = ((size_t)n * nme + slot) * M::NZ;
            double zv[NZV], rzv[NZV], sig[NZV], mz[NZV];
            bool mz_set[NZV], sig_set[NZV];
            AB_UNROLL
            for (int iz = 0; iz < M::NZ; ++iz) zv[iz] = rzv[iz] = 0.0;
            M::z(ie, t, x, S.p, S.k, S.h, spl, zv);
            M::rz(ie, t, x, S.p, S.k, S.h, spl, rzv);
            M::sigmaz(S.p, S.k, sig);
            AB_UNROLL
            for (int iz = 0; iz < M::NZ; ++iz) {
                mz[iz] = b.mz ? b.mz[zrow + iz] : ab_nan();
                mz_set[iz] = (mz[iz] == mz[iz]) && M::z2event(iz) == ie;
                const double se = b.sigmaz_in ? b.sigmaz_in[zrow + iz] : ab_nan();
                sig_set[iz] = (se == se);
                if (sig_set[iz]) sig[iz] = se;
            }
            AB_UNROLL
            for (int c = 0; c < C; ++c) {
                if (!bdf.active(c)) continue;
                const int cc = bdf.col(c);
                if (cc == 0) {
                    AB_UNROLL
                    for (int iz = 0; iz < M::NZ; ++iz) {
                        if (M::z2event(iz) != ie) continue;
                        if (full()) {
                            if (b.z) b.z[zrow + iz] = zv[iz];
                            if (b.rz && is_last) b.rz[zrow + iz] = rzv[iz];
                            if (b.sigmaz) b.sigmaz[zrow + iz] = sig[iz];
                        }
                        if (mz_set[iz]) {
                            const double pi_ = 3.14159265358979323846;
                            const double s2 = sig[iz] * sig[iz], r_ = zv[iz] - mz[iz];
                            llh -= 0.5 * log(2.0 * pi_ * s2) + 0.5 * r_ * r_ / s2;                      // Jz
                            if (is_last) llh -= 0.5 * log(2.0 * pi_ * s2) + 0.5 * rzv[iz] * rzv[iz] / s2;  // Jrz
                        }
                    }
                } else {
                    const int ip = S.ip_of_col[cc], il = cc - 1;
                    const double pc = S.pcoef[cc];
                    double sxc[NX], szv[NZV], srzv[NZV], dsig[NZV];
                    AB_UNROLL
                    for (int r = 0; r < NX; ++r) sxc[r] = S.zn[0][r][cc];
                    AB_UNROLL
                    for (int iz = 0; iz < M::NZ; ++iz) szv[iz] = srzv[iz] = 0.0;
                    if (!is_last) M::sz(ie, ip, t, x, S.p, S.k, S.h, spl, sxc, szv);
                    else M::srz(ie, ip, t, x, S.p, S.k, S.h, spl, sxc, srzv);
                    M::dsigmazdp(ip, S.p, S.k, dsig);
                    const size_t srow = (((size_t)n * nme + slot) * nplist + il) * M::NZ;
                    AB_UNROLL
                    for (int iz = 0; iz < M::NZ; ++iz) {
                        if (M::z2event(iz) != ie) continue;
                        if (full()) {
                            if (b.sz) b.sz[srow + iz] = chain(szv[iz], pc);
                            if (b.srz && is_last) b.srz[srow + iz] = chain(srzv[iz], pc);
                        }
                        if (mz_set[iz]) {
                            // Model::add_event_objective_sensitivity (reference src/model.cpp:1411-1450, :2643-2760)
                            const double s2 = sig[iz] * sig[iz], r_ = zv[iz] - mz[iz];
                            const double dsp = sig_set[iz] ? 0.0 : dsig[iz];
                            const double dJzds = 1.0 / sig[iz] - r_ * r_ / (s2 * sig[iz]);
                            if (!is_last) {
                                sllh[c] -= r_ / s2 * szv[iz] + dJzds * dsp;
                            } else {
                                const double dJrzds = 1.0 / sig[iz] - rzv[iz] * rzv[iz] / (s2 * sig[iz]);
                                sllh[c] -= rzv[iz] / s2 * srzv[iz] + (dJrzds + dJzds) * dsp;
                            }
                        }
                    }
                }
            }
            nroots[ie]++;
        }
        // at the last timepoint fill_events repeats until every slot is used (reference include/amici/forwardproblem.h:352-360)
        bool more = false;
        AB_UNROLL
        for (int ie = 0; ie < M::NE; ++ie) more = more || nroots[ie] < nme;
        if (!(is_last && more)) break;
        }
    }

    // store_pre_event_info + store_pre_event_state (reference src/forwardproblem.cpp:408-426, :576-613)
    AB_HD void store_pre_event(double t, bool seflag, bool initial_event) {
        auto& g = bdf.g;
        double x[NX];
        g.sync();
        state_col(x);
        eval_roots(t, x, rootvals);
        AB_UNROLL
        for (int ie = 0; ie < M::NE; ++ie) rval_tmp[ie] = rootvals[ie];
        if (M::NZ > 0 && recording) store_event(t);
        if (nplist > 0 || o.sensi >= 1) {
            AB_UNROLL
            for (int r = 0; r < NX; ++r) ev_x_old[r] = x[r];
            eval_xdot(t, x, ev_xdot_old);
        }
        if (nplist > 0) {
            if (!seflag && !initial_event) {
                double spl[M::NSPL > 0 ? M::NSPL : 1];
                spline_values_at<M, G, C>(S, t, spl);
                AB_UNROLL
                for (int ie = 0; ie < M::NE; ++ie) {
                    if (roots_found[ie] != 1) continue;
                    AB_UNROLL
                    for (int c = 0; c < C; ++c) {
                        const int cc = bdf.col(c);
                        stau[c] = 0.0;
                        if (!bdf.active(c) || cc == 0) continue;
                        double sxc[NX];
                        AB_UNROLL
                        for (int r = 0; r < NX; ++r) sxc[r] = S.zn[0][r][cc];
                        stau[c] = M::stau(ie, S.ip_of_col[cc], t, x, S.p, S.k, S.h, spl, sxc);
                    }
                }
            }
            if (initial_event) {
                AB_UNROLL
                for (int c = 0; c < C; ++c) stau[c] = 0.0;
            }
        }
    }

    // EventHandlingSimulator::handle_events (reference src/forwardproblem.cpp:389-514). The solver state zn[0] holds
    // x, sx at the event time t. Returns AB_SUCCESS or AB_ERROR ("stuck in an event").
    AB_HD int handle_events(double t, bool initial_event) {
        auto& g = bdf.g;
        if (!initial_event && t == tlastroot) return AB_ERROR;
        tlastroot = t;
        store_pre_event(t, false, initial_event);
        if (!initial_event) update_heaviside();
        // pending events, first in first out (EventQueue with undefined priorities, reference include/amici/event.h:136-215)
        unsigned long long queue = 0ull;
        int qn = 0;
        AB_UNROLL
        for (int ie = 0; ie < M::NE; ++ie)
            if (roots_found[ie] == 1 && qn < 16) queue |= (unsigned long long)ie << (4 * qn++);
        while (qn > 0) {
            const int ie = (int)(queue & 15ull);
            queue >>= 4;
            qn--;
            double x[NX], dx[NX], f[NX];
            g.sync();
            state_col(x);
            // Model::add_state_event_update (reference src/model.cpp:1481-1503)
            M::deltax(ie, t, x, S.p, S.k, S.h, x, dx);
            AB_UNROLL
            for (int r = 0; r < NX; ++r) x[r] += dx[r];
            g.sync();
            if (g.lane == 0) {
                AB_UNROLL
                for (int r = 0; r < NX; ++r) S.zn[0][r][0] = x[r];
            }
            g.sync();
            if (nplist > 0) {
                // Model::add_state_sensitivity_event_update (reference src/model.cpp:1505-1534)
                eval_xdot(t, x, f);
                double spl[M::NSPL > 0 ? M::NSPL : 1];
                spline_values_at<M, G, C>(S, t, spl);
                AB_UNROLL
                for (int c = 0; c < C; ++c) {
                    const int cc = bdf.col(c);
                    if (!bdf.active(c) || cc == 0) continue;
                    double sxc[NX], dsx[NX];
                    AB_UNROLL
                    for (int r = 0; r < NX; ++r) sxc[r] = S.zn[0][r][cc];
                    M::deltasx(ie, S.ip_of_col[cc], t, x, S.p, S.k, S.h, spl, f, ev_xdot_old, sxc, stau[c], ev_x_old, dsx);
                    AB_UNROLL
                    for (int r = 0; r < NX; ++r) S.zn[0][r][cc] = sxc[r] + dsx[r];
                }
            }
            // detect_secondary_events (reference src/forwardproblem.cpp:615-665)
            eval_roots(t, x, rootvals);
            int secondevent = 0;
            AB_UNROLL
            for (int ie2 = 0; ie2 < M::NE; ++ie2) {
                if (roots_found[ie2] == 0) {
                    const bool h_a = !(rval_tmp[ie2] < 0.0), h_b = !(rootvals[ie2] < 0.0);   // heaviside_differs
                    if (h_a != h_b) {
                        if (rval_tmp[ie2] < rootvals[ie2]) {
                            roots_found[ie2] = 1;
                            if (qn < 16) queue |= (unsigned long long)ie2 << (4 * qn++);
                        } else {
                            roots_found[ie2] = -1;
                        }
                        secondevent++;
                    }
                } else {
                    roots_found[ie2] = 0;
                }
            }
            if (secondevent > 0) {
                store_pre_event(t, true, initial_event);
                update_heaviside();
            }
        }
        // Solver::reinit + sens_reinit
        bdf.amici_reinit(t);
        return AB_SUCCESS;
    }

    // roots reported at time t: the solver's root directions (Solver::get_root_info) and the explicit triggers scheduled at t,
    // whose direction comes from the root values of the previous evaluation: copysign(1, -rootvals), i.e. -1 for +0.0
    // (reference src/forwardproblem.cpp:116-131, :214-231)
    AB_HD void collect_roots(const double t, const double next_t_event) {
        AB_UNROLL
        for (int ie = 0; ie < M::NE_SOLVER; ++ie) roots_found[ie] = bdf.iroots[ie];
        if (t == next_t_event) {
            AB_UNROLL
            for (int ie = M::NE_SOLVER; ie < M::NE; ++ie)
                if (M::explicit_time(ie, S.p, S.k) == t) roots_found[ie] = (copysign(1.0, rootvals[ie]) < 0.0) ? 1 : -1;
            t_trig_done = t;
        }
    }
    // EventHandlingSimulator::handle_initial_events (reference src/forwardproblem.cpp:516-522)
    AB_HD void handle_initial_events(const double t) {
        bool any = false;
        AB_UNROLL
        for (int ie = 0; ie < M::NE; ++ie) any = any || roots_found[ie] == 1;
        if (any) handle_events(t, true);
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
        reset_objective();
        const bool has_preeq = PREEQ && b.k_preeq != nullptr;
        // presimulation (reference src/forwardproblem.cpp:284-321): integrate [t0 - t_presim, t0] with the presimulation
        // fixed parameters (the simulation's if none are given, src/edata.cpp:531-548), nothing is recorded
        const bool has_presim = PREEQ && b.t_presim != nullptr && b.t_presim[n] > 0.0;
        const double* const k_presim = (PREEQ && b.k_presim) ? b.k_presim : b.k;
        load_condition(has_preeq ? b.k_preeq : has_presim ? k_presim : b.k);
        double t = o.t0;
        int status = AB_SUCCESS;
        int it = 0;
        // ---- pre-equilibration state (SteadyStateProblem, reference src/forwardproblem.cpp:755-1223) ----
        const int ncols_main = bdf.ncols;
        int pe_status[3] = {SS_NOT_RUN, SS_NOT_RUN, SS_NOT_RUN}, pe_steps[3] = {0, 0, 0};
        double pe_wrms = ab_nan(), pe_t = ab_nan();
        // phase: 0 = main simulation, 1 = steady-state simulation, 2 = steady state found (finish), 3 = simulation failed,
        //        4 = presimulation
        int phase = 0;
        bool newton_on = false, sim_on = false, preeq_failed = false;
        // post-equilibration (t = inf output points, reference src/forwardproblem.cpp:376-387): the same steady-state machinery,
        // started from the solver state at the last finite timepoint.  pre_* keep the pre-equilibration results meanwhile.
        bool posteq = false, posteq_failed = false;
        int pre_status[3] = {SS_NOT_RUN, SS_NOT_RUN, SS_NOT_RUN}, pre_steps[3] = {0, 0, 0};
        double pre_wrms = ab_nan(), pre_t = ab_nan();
        if (PREEQ && has_preeq) {
            const bool sens = ncols_main > 1;
            newton_on = !(o.ss_mode == SS_INTEGRATION_ONLY || o.newton_maxsteps == 0 ||
                          (sens && o.ss_sensi_mode == SS_INTEGRATION_ONLY));
            sim_on = o.ss_mode != SS_NEWTON_ONLY;
            phase = 3;
            if (M::NE > 0) {
                // Model::initialize at t0 under the pre-equilibration condition (src/forwardproblem.cpp:265-269); events of this
                // period act on the state but store nothing (EventHandlingSimulator with edata == nullptr)
                recording = false;
                initialize_events(o.t0);
                if (newton_on) handle_initial_events(o.t0);   // find_steady_state_by_newtons_method (:1160-1164)
            }
            if (newton_on && preeq_newton(0, pe_status, pe_steps, pe_wrms)) phase = 2;
            else if (sim_on) {
                // find_steady_state_by_simulation: restart from the initial state with a fresh solver; sensitivities
                // are integrated along unless they come from the Newton solve afterwards (:872-920, :1023-1033)
                if (newton_on && M::NE == 0 && g.lane == 0) {
                    double v[NX];
                    initial_column(0, v);
                    for (int r = 0; r < NX; ++r) S.zn[0][r][0] = v[r];
                }
                if (newton_on && M::NE > 0) {
                    // the stored initial state (all columns, Heaviside variables) comes back; the workspace's root flags do
                    // not (they were cleared while the Newton attempt handled the initial events), so the fresh simulator
                    // finds nothing to handle at t0
                    g.sync();
                    AB_UNROLL
                    for (int c = 0; c < C; ++c) {
                        if (!bdf.active(c)) continue;
                        double v[NX];
                        initial_column(bdf.col(c), v);
                        AB_UNROLL
                        for (int r = 0; r < NX; ++r) S.zn[0][r][bdf.col(c)] = v[r];
                    }
                    g.sync();
                    double rv_keep[NEV], rt_keep[NEV];
                    AB_UNROLL
                    for (int ie = 0; ie < M::NE; ++ie) { rv_keep[ie] = rootvals[ie]; rt_keep[ie] = rval_tmp[ie]; }
                    initialize_events(o.t0);
                    AB_UNROLL
                    for (int ie = 0; ie < M::NE; ++ie) { roots_found[ie] = 0; rootvals[ie] = rv_keep[ie]; rval_tmp[ie] = rt_keep[ie]; }
                }
                g.sync();
                bdf.ncols = (sens && o.ss_sensi_mode != SS_NEWTON_ONLY) ? ncols_main : 1;
                bdf.init(o.t0);   // the fresh solver is set up first (:902-904); initial events then leave a re-initialisation pending
                if (M::NE > 0) handle_initial_events(o.t0);   // run_steadystate_simulation_fwd (:1088)
                phase = 1;
            }
        }
        if (PREEQ && phase == 0 && has_presim) {
            // Model::initialize at t0 - t_presim under the presimulation context, Solver::setup (src/forwardproblem.cpp:299-316)
            t = o.t0 - b.t_presim[n];
            if (M::NE > 0) {
                recording = false;   // the presimulation's events belong to another simulator's record
                initialize_events(t);
            }
            reinitialize_states(o.reinit_states_presim);   // update_and_reinit_states_and_sensitivities (:316)
            bdf.init(t);
            if (M::NE > 0) {
                handle_initial_events(t);
                double x[NX];
                g.sync();
                state_col(x);
                eval_roots(t, x, rootvals);
            }
            phase = 4;
        } else if (phase == 0) {
            if (M::NE > 0) initialize_events(o.t0);
            // Solver::setup comes before the simulator handles the initial events (src/forwardproblem.cpp:341-345, :62): an
            // initial event leaves a solver re-initialisation pending
            bdf.init(o.t0);
            if (M::NE > 0) handle_initial_events(o.t0);   // EventHandlingSimulator::handle_initial_events (:516-522)
            write_initial_state();
            if (M::NE > 0) {   // root information at t0 (src/forwardproblem.cpp:66-67)
                double x[NX];
                g.sync();
                state_col(x);
                eval_roots(o.t0, x, rootvals);
            }
        }
        // ---- condition driver: the one call site of the integrator serves both phases ----
        for (;;) {
            double tout = 0.0;
            bool one_step = false;
            double next_t_event = INFINITY;
            if (PREEQ && phase != 0) {
                if (phase == 4) {
                    if (t < o.t0 && !next_t_too_close(t, o.t0)) {
                        tout = o.t0;
                        if (M::NE > M::NE_SOLVER) {   // EventHandlingSimulator::run with timepoints = {t0} (src/forwardproblem.cpp:319, :70-106)
                            next_t_event = next_trigger(o.t0);
                            tout = ab_min(tout, next_t_event);
                        }
                    } else {
                        // handle_main_simulation after a presimulation (src/forwardproblem.cpp:323-372): back to the simulation
                        // condition, x0 / sx0 = the presimulated state, Solver::reinit + sens_reinit at t0
                        g.sync();
                        if (g.lane == 0) {
                            for (int i = 0; i < M::NK; ++i) S.k[i] = b.k[(size_t)n * M::NK + i];
                            M::derived(S.p, S.k);
                        }
                        g.sync();
                        // the root flags for the main simulation come from the presimulated state, before the state
                        // re-initialisation reaches the workspace (src/forwardproblem.cpp:350-365)
                        if (M::NE > 0) initialize_events(o.t0, true);
                        reinitialize_states(o.reinit_states_sim);
                        if (M::NE > 0) {
                            recording = true;
                            handle_initial_events(o.t0);
                        }
                        write_initial_state();
                        bdf.amici_reinit(o.t0);   // the solver is not set up again: counters and step-size history carry on
                        t = o.t0;
                        phase = 0;
                        if (M::NE > 0) {   // root information at t0 (src/forwardproblem.cpp:66-67)
                            double x[NX];
                            g.sync();
                            state_col(x);
                            eval_roots(o.t0, x, rootvals);
                        }
                    }
                }
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
                        if (M::NE > M::NE_SOLVER) {   // run_steady_state: the next time-triggered event is the target (src/forwardproblem.cpp:166-176, :203-205)
                            next_t_event = next_trigger(INFINITY);
                            if (isfinite(next_t_event)) tout = next_t_event;
                        }
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
                        if (posteq) posteq_failed = true;
                        else preeq_failed = true;
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
                            if (posteq) posteq_failed = true;
                            else preeq_failed = true;
                            break;
                        }
                    }
                    bdf.ncols = ncols_main;
                    if (posteq) break;   // the t = inf outputs are written after the loop
                    write_steady_state();
                    // back to the simulation condition (ConditionContext restore) and on to the main simulation
                    g.sync();
                    if (g.lane == 0) {
                        for (int i = 0; i < M::NK; ++i) S.k[i] = (has_presim ? k_presim : b.k)[(size_t)n * M::NK + i];
                        M::derived(S.p, S.k);
                    }
                    g.sync();
                    if (has_presim) {   // the presimulation starts from the steady state (src/forwardproblem.cpp:301-316)
                        t = o.t0 - b.t_presim[n];
                        if (M::NE > 0) initialize_events(t, true);   // Model::reinit_events with the Heaviside state of the pre-equilibration (:305-311)
                        reinitialize_states(o.reinit_states_presim);
                        bdf.init(t);
                        if (M::NE > 0) handle_initial_events(t);    // pre_simulator_.run (:319): still nothing is stored
                        phase = 4;
                    } else {
                        t = o.t0;
                        if (M::NE > 0) initialize_events(t, true);   // :353-360, from the steady state before its re-initialisation
                        reinitialize_states(o.reinit_states_sim);   // src/forwardproblem.cpp:350-352
                        bdf.init(t);
                        if (M::NE > 0) {
                            recording = true;
                            handle_initial_events(t);
                        }
                        write_initial_state();
                        phase = 0;
                    }
                    if (M::NE > 0) {   // root information at the start of the period (src/forwardproblem.cpp:66-67)
                        double x[NX];
                        g.sync();
                        state_col(x);
                        eval_roots(t, x, rootvals);
                    }
                    AB_UNROLL
                    for (int i = 0; i < 3; ++i) {   // keep the pre-equilibration record; pe_* are re-used by a post-equilibration
                        pre_status[i] = pe_status[i];
                        pre_steps[i] = pe_steps[i];
                        pe_status[i] = SS_NOT_RUN;
                        pe_steps[i] = 0;
                    }
                    pre_wrms = pe_wrms;
                    pre_t = pe_t;
                    pe_wrms = pe_t = ab_nan();
                    if (phase == 4) continue;   // next: the presimulation
                }
            }
            if (!PREEQ || phase == 0) {
                if (it >= nt) break;
                tout = b.ts[it];
                if (PREEQ && tout == INFINITY) {
                    // SteadyStateProblem(is_preeq = false)::run from the current solver state (src/forwardproblem.cpp:812-870)
                    posteq = true;
                    const bool sens = ncols_main > 1;
                    newton_on = !(o.ss_mode == SS_INTEGRATION_ONLY || o.newton_maxsteps == 0 ||
                                  (sens && o.ss_sensi_mode == SS_INTEGRATION_ONLY));
                    sim_on = o.ss_mode != SS_NEWTON_ONLY;
                    g.sync();
                    AB_UNROLL
                    for (int c = 0; c < C; ++c) {   // Newton's method works on zn[0] in place: keep the solver's state (acorbuf is free here)
                        if (!bdf.active(c)) continue;
                        AB_UNROLL
                        for (int r = 0; r < NX; ++r) S.acorbuf[r][bdf.col(c)] = S.zn[0][r][bdf.col(c)];
                    }
                    g.sync();
                    phase = 3;
                    if (newton_on && preeq_newton(0, pe_status, pe_steps, pe_wrms)) phase = 2;
                    else if (sim_on) {
                        g.sync();
                        AB_UNROLL
                        for (int c = 0; c < C; ++c) {
                            if (!bdf.active(c)) continue;
                            AB_UNROLL
                            for (int r = 0; r < NX; ++r) S.zn[0][r][bdf.col(c)] = S.acorbuf[r][bdf.col(c)];
                        }
                        g.sync();
                        // "Postequilibration -> Solver was already created, use that one" (src/forwardproblem.cpp:877-878): the
                        // integrator continues with its history; only steady-state sensitivity mode newtonOnly switches its
                        // forward sensitivities off (switch_forward_sensis_off, :1022-1026)
                        bdf.ncols = (sens && o.ss_sensi_mode != SS_NEWTON_ONLY) ? ncols_main : 1;
                        phase = 1;
                    }
                    continue;
                }
                if (!(tout > o.t0) || !(t < tout) || next_t_too_close(t, tout)) {
                    datapoint(it);
                    ++it;
                    continue;
                }
                if (M::NE > M::NE_SOLVER) {   // next stop: output time or time-triggered event (src/forwardproblem.cpp:100-106)
                    next_t_event = next_trigger(b.ts[nt - 1]);
                    tout = ab_min(tout, next_t_event);
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
            if (PREEQ && M::NE > 0 && phase == 1) {
                // run_steady_state (src/forwardproblem.cpp:208-235): Solver::step may have over-stepped the next time-triggered
                // event; the solution is read back there (Solver::write_solution(t, sol)) and the solver restarts from it
                if (flag == AB_ROOT_RETURN || t >= next_t_event) {
                    if (t >= next_t_event) {
                        if (t > next_t_event) bdf.collapse_to(next_t_event);
                        t = next_t_event;
                    }
                    collect_roots(t, next_t_event);
                    if (handle_events(t, false) != AB_SUCCESS) {   // AmiException inside the steady-state simulation (:938-945)
                        pe_status[1] = SS_FAILED;
                        phase = 3;
                    }
                }
                continue;
            }
            if (!PREEQ || phase == 0 || (M::NE > 0 && phase == 4)) {
                if (!PREEQ || phase == 0) {
                    g.sync();
                    if (g.lane == 0) {
                        AB_UNROLL
                        for (int r = 0; r < NX; ++r) S.xfin[r] = S.zn[0][r][0];
                    }
                }
                if (M::NE > 0 && (flag == AB_ROOT_RETURN || t == next_t_event)) {
                    // solver-tracked or time-triggered event (src/forwardproblem.cpp:116-134)
                    collect_roots(t, next_t_event);
                    const int rc = handle_events(t, false);
                    if (rc != AB_SUCCESS) {
                        status = rc;
                        break;
                    }
                }
            }
        }
        // fill unused event-observable slots at the final timepoint (src/forwardproblem.cpp:144-148)
        if (M::NZ > 0 && nt > 0 && status == AB_SUCCESS && !(PREEQ && preeq_failed)) {
            bool more = false;
            AB_UNROLL
            for (int ie = 0; ie < M::NE; ++ie) more = more || nroots[ie] < o.nmaxevent;
            if (more) {
                AB_UNROLL
                for (int ie = 0; ie < M::NE; ++ie) roots_found[ie] = 0;
                store_event(t);
            }
        }
        if (PREEQ && posteq) {
            // ReturnData::process_post_equilibration (reference src/rdata.cpp:288-312): the t = inf outputs come from the final
            // state of the steady-state problem, also when it failed (status is AMICI_ERROR then, nothing is invalidated)
            bdf.ncols = ncols_main;
            g.sync();
            if (g.lane == 0) {
                AB_UNROLL
                for (int r = 0; r < NX; ++r) S.xfin[r] = S.zn[0][r][0];
            }
            g.sync();
            for (; it < nt; ++it) datapoint(it, true);
            if (pe_status[1] != SS_SUCCESS || posteq_failed) t = INFINITY;
            if (g.lane == 0) {
                for (int i = 0; i < 3; ++i) {
                    if (b.posteq_status) b.posteq_status[(size_t)n * 3 + i] = pe_status[i];
                    if (b.posteq_numsteps) b.posteq_numsteps[(size_t)n * 3 + i] = pe_steps[i];
                }
                if (b.posteq_wrms) b.posteq_wrms[n] = pe_wrms;
                if (b.posteq_t) b.posteq_t[n] = pe_t;
            }
        }
        if (PREEQ && o.posteq && !posteq && g.lane == 0) {
            // the post-equilibration never started (failure before the t = inf points): ReturnData defaults (src/rdata.cpp:172-175)
            if (b.posteq_wrms) b.posteq_wrms[n] = ab_nan();
            if (b.posteq_t) b.posteq_t[n] = ab_nan();
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
            const bool done = !preeq_failed;   // a finished pre-equilibration moved its record to pre_*
            for (int i = 0; i < 3; ++i) {
                if (b.preeq_status) b.preeq_status[(size_t)n * 3 + i] = done ? pre_status[i] : pe_status[i];
                if (b.preeq_numsteps) b.preeq_numsteps[(size_t)n * 3 + i] = done ? pre_steps[i] : pe_steps[i];
            }
            if (b.preeq_wrms) b.preeq_wrms[n] = done ? pre_wrms : pe_wrms;
            if (b.preeq_t) b.preeq_t[n] = done ? pre_t : pe_t;
        }
        if (status != AB_SUCCESS && !(PREEQ && posteq_failed)) invalidate(it, (PREEQ && preeq_failed) ? 0.0 : ab_nan());
        finish(status, t);
    }

    AB_HD void reset_objective() {
        llh = 0.0;
        chi2 = 0.0;
        AB_UNROLL
        for (int j = 0; j < C_OUT; ++j) sllh[j] = 0.0;
    }

    // end of a condition: xdot / J at the final state, status and objective
    AB_HD void finish(const int status, const double t) {
        auto& g = bdf.g;
        // xdot and J at the final state = the last solution the solver returned (ReturnData::store_jacobian_and_derivative,
        // reference rdata.h:707-728; FinalStateStorer, src/forwardproblem.cpp:245-256)
        if (full() && (b.xdot || b.J)) {
            g.sync();
            if (g.lane == 0) {
                double xf[NX], f[NX], Jv[M::NNZ_J > 0 ? M::NNZ_J : 1], spl[M::NSPL > 0 ? M::NSPL : 1];
                AB_UNROLL
                for (int r = 0; r < NX; ++r) xf[r] = S.xfin[r];
                x_pos<M>(S, xf);
                spline_values_at<M, G, C>(S, t, spl);
                M::xdot(t, xf, S.p, S.k, S.h, spl, f);
                M::jac(t, xf, S.p, S.k, S.h, spl, Jv);
                if (b.xdot)
                    for (int r = 0; r < NX; ++r) b.xdot[(size_t)n * NX + r] = f[r];
                if (b.J) M::jac_to_dense(Jv, b.J + (size_t)n * NX * NX);
            }
        }
        AB_UNROLL
        for (int j = 0; j < C_OUT; ++j) {
            if (!out_active(j)) continue;
            const int oc = out_col(j);
            if (oc == 0) {
                if (b.status) b.status[n] = status;
                if (b.llh) b.llh[n] = llh;
                if (b.chi2) b.chi2[n] = chi2;
            } else if (b.sllh) {
                b.sllh[(size_t)n * nplist + (oc - 1)] = chain(sllh[j], S.pcoef_out[oc]);
            }
        }
    }
};

}  // namespace amici_b200
#endif
