// One condition end-to-end on one lane-group: the batched counterpart of the reference's
// run_simulation -> ForwardProblem -> EventHandlingSimulator::run -> ReturnData assembly
// (reference src/amici.cpp:65-218, src/forwardproblem.cpp:53-151, src/rdata.cpp:213-432, :749-875).
#ifndef AMICI_B200_SIMULATE_CUH
#define AMICI_B200_SIMULATE_CUH
#include "bdf_core.cuh"

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

template <class M, int G, int C>
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
    AB_HD void load_condition() {
        auto& g = bdf.g;
        g.sync();
        if (g.lane == 0) {
            for (int i = 0; i < M::NP; ++i) {
                const double v = b.p[(size_t)n * M::NP + i];
                const int sc = b.pscale ? b.pscale[i] : AB_PSCALE_NONE;
                S.p[i] = sc == AB_PSCALE_LOG10 ? pow(10.0, v) : sc == AB_PSCALE_LN ? exp(v) : v;
            }
            for (int i = 0; i < M::NK; ++i) S.k[i] = b.k[(size_t)n * M::NK + i];
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

    // ReturnData::invalidate (reference src/rdata.cpp:684-735)
    AB_HD void invalidate(int it_start) {
        const int nt = b.nt;
        const double nan_ = ab_nan();
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
                        if (b.res) b.res[(size_t)n * nt * NY + (size_t)it * NY + iy] = nan_;
                    }
                } else {
                    const int il = cc - 1;
                    for (int r = 0; r < NX; ++r)
                        if (b.sx) b.sx[(((size_t)n * nt + it) * nplist + il) * NX + r] = nan_;
                    for (int iy = 0; iy < NY; ++iy) {
                        if (b.sy) b.sy[(((size_t)n * nt + it) * nplist + il) * NY + iy] = nan_;
                        if (b.sres)
                            b.sres[((size_t)n * nt * NY + (size_t)it * NY + iy) * nplist + il] = nan_;
                    }
                }
            }
            if (cc > 0 && b.FIM)
                for (int jl = 0; jl < nplist; ++jl) b.FIM[((size_t)n * nplist + (cc - 1)) * nplist + jl] = nan_;
        }
    }

    AB_HD void run() {
        auto& g = bdf.g;
        const int nt = b.nt;
        llh = 0.0;
        chi2 = 0.0;
        AB_UNROLL
        for (int c = 0; c < C; ++c) sllh[c] = 0.0;
        load_condition();
        // x0 / sx0 outputs
        AB_UNROLL
        for (int c = 0; c < C; ++c) {
            if (!bdf.active(c) || !full()) continue;
            const int cc = bdf.col(c);
            AB_UNROLL
            for (int r = 0; r < NX; ++r) {
                const double v = S.zn[0][r][cc];
                if (cc == 0) { if (b.x0) b.x0[(size_t)n * NX + r] = v; }
                else if (b.sx0) b.sx0[((size_t)n * nplist + (cc - 1)) * NX + r] = chain(v, S.pcoef[cc]);
            }
        }
        bdf.init(o.t0);
        double t = o.t0;
        int status = AB_SUCCESS;
        int it = 0;
        for (; it < nt; ++it) {
            const double tout = b.ts[it];
            if (tout > o.t0) {
                while (t < tout) {
                    if (next_t_too_close(t, tout)) break;
                    const int flag = bdf.solve_to(tout);
                    t = bdf.tretlast;
                    if (flag < 0) { status = flag; break; }
                }
                if (status != AB_SUCCESS) break;
            }
            datapoint(it);
        }
        if (status != AB_SUCCESS) invalidate(it);
        // xdot and J at the final state (ReturnData::store_jacobian_and_derivative, reference rdata.h:707-728)
        if (full() && (b.xdot || b.J)) {
            typename Solver::Eval e;
            double xc[C][NX];
            AB_UNROLL
            for (int c = 0; c < C; ++c) {
                AB_UNROLL
                for (int r = 0; r < NX; ++r) xc[c][r] = bdf.active(c) ? S.zn[0][r][bdf.col(c)] : 0.0;
            }
            bcast_state<M, G, C>(S, g, xc, e);
            eval_splines<M, G, C>(S, g, bdf.ncols, t, e);
            if (g.lane == 0) {
                double f[NX], Jv[M::NNZ_J > 0 ? M::NNZ_J : 1];
                M::xdot(t, e.y, S.p, S.k, S.h, e.spl, f);
                M::jac(t, e.y, S.p, S.k, S.h, e.spl, Jv);
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
