// dae.cuh — index-1 DAE models: forward simulation of M x' = f(t, x, p, k) with the mass matrix
// M = diag(state is differential), the form of the reference's DAE models (reference src/model_dae.cpp:31-64: residual
// F = xdot(x) - M dx, Jacobian dF/dx - cj M; the shipped model_robertson_py: two differential states and the algebraic
// equation 0 = x1 + x2 + x3 - 1).  ONE THREAD PER CONDITION on ScalarBdf (adjoint.cuh).
//
// The reference integrates these models with IDAS (variable-coefficient BDF in fixed-leading-coefficient form,
// src/solver_idas.cpp).  This path is NOT a restatement of IDAS: it is the engine's CVODES-equivalent Nordsieck BDF with a
// mass matrix — same formulas of order 1..5, same step-size / order control (BdfControl), corrector
// M (rl1 zn[1] + acor) - gamma f = 0 with the Newton matrix M - gamma J, algebraic states left out of the error test as
// AMICI configures IDAS (IDASetSuppressAlg(true), src/solver.cpp:252).  Consequences of that choice handled in ScalarBdf: the
// history of the algebraic states is kept at order 1 (linear extrapolation as predictor; a Nordsieck array of order q fed with
// corrections that no error test bounds drifts until the corrector stalls), every step estimates its own Newton convergence
// rate, the corrector tolerance is ten times tighter, and a step size that underflows to zero ends the solve.  Step sequences
// therefore differ from IDAS'; results are compared with the reference's IDAS runs by the parity protocol (error against a tight-tolerance IDAS run no larger
// than IDAS' own, tests/parity.py).
//
//   consistent start   IDACalcIC(IDA_YA_YDP_INIT) (src/solver_idas.cpp:831-840): differential states fixed, Newton on the
//                      algebraic equations for the algebraic states at t0 (the shipped Robertson model starts from the
//                      inconsistent x3 = 0); the reported x0 is the corrected one, as in the reference's ReturnData
//   events             time-triggered Heaviside switches only (NE_SOLVER == 0, no state updates; model_calvetti_py's forcing):
//                      the trigger times are stop times (src/forwardproblem.cpp:100-106), h += direction of the crossing
//                      (Model::update_heaviside), the solver restarts at order 1 as after every AMICI event; the algebraic
//                      states are NOT re-initialised there — the first step after the event pulls them onto the new
//                      constraint, exactly what IDAS does after IDAReInit with the error test blind to them
//   sensitivities      first-order forward sensitivities of event-free models: M s' = J s + df/dp column by column, corrected after
//                      the states by a Newton iteration on the states' factors (the staggered corrector; ScalarBdf::sens_correct),
//                      error test over states and sensitivities (errconS), algebraic rows of sx(t0) from the differentiated
//                      constraints; sx, sy, ssigmay, sllh, sres and the FIM reported
//   outputs            x, y, sigmay, llh, res, chi2, counters, status (AdjointSimulation's forward outputs)
//   not implemented    adjoint / second-order sensitivities, sensitivities of models with switches, state-triggered or
//                      state-changing events, pre-/post-equilibration of DAE models
#ifndef AMICI_B200_DAE_CUH
#define AMICI_B200_DAE_CUH
#include "adjoint.cuh"

namespace amici_b200 {

// S: with forward sensitivities (the NQ rows of ScalarBdf's second vector are the columns s_il = dx / dp_plist[il])
template <class M, bool S = false>
struct DaeProblem {
    static constexpr int NX = M::NX, NQ = S ? (M::NP > 0 ? M::NP : 1) * M::NX : 0;
    static constexpr bool HIST = false, SENS = S;
    const double* s_now = nullptr;   // the solver's current sensitivities (for the derivative at a restart and cvHin)
    static constexpr unsigned long long ALG_MASK = alg_mask_of<M>::value;
    const AdjCond<M>& c;
    double rtol, atol, rtolq, atolq;
    AB_NI bool rhs(double t, const double* y, double* f) const {
        double spl[AdjCond<M>::NSPLV];
        c.splines(t, spl);
        M::xdot(t, y, c.p, c.k, c.h, spl, f);
        bool ok = true;
        for (int r = 0; r < NX; ++r) ok = ok && isfinite(f[r]);
        return ok;
    }
    AB_HD void jac(double t, const double* y, double (&J)[NX][NX]) const {
        double spl[AdjCond<M>::NSPLV], Jv[M::NNZ_J > 0 ? M::NNZ_J : 1];
        c.splines(t, spl);
        M::jac(t, y, c.p, c.k, c.h, spl, Jv);
        M::jac_to_dense(Jv, &J[0][0]);
    }
    // out <- J(t, y) s + dxdot/dp for every column (Model_DAE::fsxdot, reference src/model_dae.cpp; algebraic rows: dg/dx s + dg/dp)
    AB_NI bool sens_rhs(double t, const double* y, const double* s, double* out) const {
        double spl[AdjCond<M>::NSPLV], sspl[AdjCond<M>::NSPLV], Jv[M::NNZ_J > 0 ? M::NNZ_J : 1];
        c.splines(t, spl);
        M::jac(t, y, c.p, c.k, c.h, spl, Jv);
        bool ok = true;
        for (int il = 0; il < c.nplist; ++il) {
            double col[NX], sc[NX];
            for (int r = 0; r < NX; ++r) sc[r] = s[il * NX + r];
            c.spline_sens(t, il, spl, sspl);
            M::dxdotdp(c.plist[il], t, y, c.p, c.k, c.h, spl, sspl, col);
            M::jac_mul_add(Jv, sc, col);
            for (int r = 0; r < NX; ++r) {
                out[il * NX + r] = col[r];
                ok = ok && isfinite(col[r]);
            }
        }
        return ok;
    }
    AB_HD bool quad(double t, const double* y, double* fq) const { return S && s_now ? sens_rhs(t, y, s_now, fq) : true; }
    AB_HD void store_point(double, int, const double*, const double*) {}
};

template <class M>
struct DaeSimulation {
    static constexpr int NX = M::NX;
    static constexpr unsigned long long AM = alg_mask_of<M>::value;
    const SolverOptions& o;
    const BatchView& b;
    int n;
    AB_HD DaeSimulation(const SolverOptions& opt, const BatchView& bv, int n_) : o(opt), b(bv), n(n_) {}
    AB_HD static bool alg(int r) { return ((AM >> r) & 1ull) != 0ull; }

    // IDACalcIC(IDA_YA_YDP_INIT): Newton on the algebraic rows for the algebraic states, differential states fixed.
    // Convergence as IDAS tests it (idas_ic.c: WRMS norm of the update below epiccon = 0.01 * 0.33), at most 10 iterations
    // with a fresh Jacobian each (maxnj x maxnit of IDAS are 4 x 10).
    template <class DP>
    AB_NI bool consistent_start(const DP& dp, double t0, double* x) const {
        int ia[NX], na = 0;
        for (int r = 0; r < NX; ++r)
            if (alg(r)) ia[na++] = r;
        if (na == 0) return true;
        for (int it = 0; it < 10; ++it) {
            double f[NX], J[NX][NX], A[NX][NX], rhs[NX];
            if (!dp.rhs(t0, x, f)) return false;
            dp.jac(t0, x, J);
            for (int i = 0; i < na; ++i) {
                rhs[i] = -f[ia[i]];
                for (int j = 0; j < na; ++j) A[i][j] = J[ia[i]][ia[j]];
            }
            for (int kk = 0; kk < na; ++kk) {   // Gaussian elimination with partial pivoting on the na x na block
                int pr = kk;
                for (int r = kk + 1; r < na; ++r)
                    if (fabs(A[r][kk]) > fabs(A[pr][kk])) pr = r;
                if (A[pr][kk] == 0.0) return false;
                if (pr != kk) {
                    for (int cc = 0; cc < na; ++cc) { const double tmp = A[kk][cc]; A[kk][cc] = A[pr][cc]; A[pr][cc] = tmp; }
                    const double tmp = rhs[kk]; rhs[kk] = rhs[pr]; rhs[pr] = tmp;
                }
                for (int r = kk + 1; r < na; ++r) {
                    const double mlt = A[r][kk] / A[kk][kk];
                    for (int cc = kk; cc < na; ++cc) A[r][cc] -= mlt * A[kk][cc];
                    rhs[r] -= mlt * rhs[kk];
                }
            }
            double nrm = 0.0;
            for (int i = na - 1; i >= 0; --i) {
                double s = rhs[i];
                for (int j = i + 1; j < na; ++j) s -= A[i][j] * rhs[j];
                rhs[i] = s / A[i][i];
            }
            for (int i = 0; i < na; ++i) {
                const double w = 1.0 / (dp.rtol * fabs(x[ia[i]]) + dp.atol);
                nrm += rhs[i] * w * rhs[i] * w;
                x[ia[i]] += rhs[i];
            }
            if (!(nrm == nrm)) return false;
            if (sqrt(nrm / NX) <= 0.0033) return true;
        }
        return false;
    }

    // next explicit trigger time in (t_done, t_end] (Simulation::next_trigger)
    AB_HD double next_trigger(const AdjCond<M>& c, double t_done, double t_end) const {
        double best = INFINITY;
        for (int ie = 0; ie < M::NE; ++ie) {
            const double te = M::explicit_time(ie, c.p, c.k);
            if (te > t_done && te <= t_end && te < best) best = te;
        }
        return best;
    }
    AB_HD void roots_at(const AdjCond<M>& c, double t, const double* x, double* rv) const {
        double spl[AdjCond<M>::NSPLV];
        c.splines(t, spl);
        M::root(t, x, c.p, c.k, c.h, spl, rv);
    }

    // the algebraic rows of the initial sensitivities from the differentiated constraints, dg/dx_a s_a = -(dg/dx_d s_d + dg/dp)
    // (IDACalcIC with sensitivities): sens_rhs with s_a = 0 gives the right-hand side
    template <class DP>
    AB_NI bool consistent_start_sens(const DP& dp, double t0, const double* x, double* s, int nplist) const {
        int ia[NX], na = 0;
        for (int r = 0; r < NX; ++r)
            if (alg(r)) ia[na++] = r;
        if (na == 0 || nplist == 0) return true;
        double J[NX][NX], g[(M::NP > 0 ? M::NP : 1) * NX];
        for (int il = 0; il < nplist; ++il)
            for (int i = 0; i < na; ++i) s[il * NX + ia[i]] = 0.0;
        if (!dp.sens_rhs(t0, x, s, g)) return false;
        dp.jac(t0, x, J);
        for (int il = 0; il < nplist; ++il) {
            double A[NX][NX], rhs[NX];
            for (int i = 0; i < na; ++i) {
                rhs[i] = -g[il * NX + ia[i]];
                for (int j = 0; j < na; ++j) A[i][j] = J[ia[i]][ia[j]];
            }
            for (int kk = 0; kk < na; ++kk) {
                int pr = kk;
                for (int r = kk + 1; r < na; ++r)
                    if (fabs(A[r][kk]) > fabs(A[pr][kk])) pr = r;
                if (A[pr][kk] == 0.0) return false;
                if (pr != kk) {
                    for (int cc = 0; cc < na; ++cc) { const double tmp = A[kk][cc]; A[kk][cc] = A[pr][cc]; A[pr][cc] = tmp; }
                    const double tmp = rhs[kk]; rhs[kk] = rhs[pr]; rhs[pr] = tmp;
                }
                for (int r = kk + 1; r < na; ++r) {
                    const double mlt = A[r][kk] / A[kk][kk];
                    for (int cc = kk; cc < na; ++cc) A[r][cc] -= mlt * A[kk][cc];
                    rhs[r] -= mlt * rhs[kk];
                }
            }
            for (int i = na - 1; i >= 0; --i) {
                double v = rhs[i];
                for (int j = i + 1; j < na; ++j) v -= A[i][j] * rhs[j];
                rhs[i] = v / A[i][i];
                s[il * NX + ia[i]] = rhs[i];
            }
        }
        return true;
    }

    AB_HD void run() {
        if (o.sensi >= 1 && b.nplist > 0 && M::NP > 0) run_impl<true>();
        else run_impl<false>();
    }

    template <bool S>
    AB_HD void run_impl() {
        FwdHistory<NX> none;
        none.t = none.y = nullptr;
        none.ord = nullptr;
        none.stride = 1;
        none.cap = 0;
        AdjointSimulation<M> out(o, b, none, nullptr, n);   // load / datapoint / invalidate / finish of the forward outputs
        out.load();
        out.fsa = S;
        if (!S) out.c.nplist = 0;
        const int nt = b.nt, nplist = out.c.nplist;
        const double rtol_s = o.rtol_fsa == o.rtol_fsa && o.rtol_fsa > 0.0 ? o.rtol_fsa : o.rtol;
        const double atol_s = o.atol_fsa == o.atol_fsa && o.atol_fsa > 0.0 ? o.atol_fsa : o.atol;
        DaeProblem<M, S> dp{nullptr, out.c, o.rtol, o.atol, rtol_s, atol_s};
        // IDAS allows four Newton iterations per step (MAXIT, idas_impl.h) where CVODES allows three: the jump of the algebraic
        // states after a switch is absorbed by the first corrector solve after the restart, whatever the step size
        SolverOptions od = o;
        od.max_nonlin_iters = o.max_nonlin_iters < 4 ? 4 : o.max_nonlin_iters;
        ScalarBdf<DaeProblem<M, S>> bdf(dp, od, S ? nplist * NX : 0);
        if (S) dp.s_now = bdf.znQ[0];
        double x0[NX];
        M::x0(out.c.p, out.c.k, x0);
        if (b.x0_in)
            for (int r = 0; r < NX; ++r) x0[r] = b.x0_in[(size_t)n * NX + r];
        int status = AB_SUCCESS, it = 0;
        if (!consistent_start(dp, o.t0, x0)) status = AB_IC_FAIL;
        if (b.x0 && out.full())
            for (int r = 0; r < NX; ++r) b.x0[(size_t)n * NX + r] = x0[r];
        constexpr int NEV = M::NE > 0 ? M::NE : 1;
        double rootvals[NEV];
        if (M::NE > 0) {   // Model::initialize_events (reference src/model.cpp:407-442): the Heaviside state of the roots at t0
            roots_at(out.c, o.t0, x0, rootvals);
            for (int ie = 0; ie < M::NE; ++ie) out.c.h[ie] = rootvals[ie] < 0.0 ? 0.0 : 1.0;
        }
        if (S && status == AB_SUCCESS) {   // sx(t0): Model::fsx0 (+ ExpData::sx0 on the parameter scale), algebraic rows made consistent
            for (int il = 0; il < nplist; ++il) {
                double col[NX];
                M::sx0(out.c.plist[il], out.c.p, out.c.k, col);
                if (b.sx0_in)
                    for (int r = 0; r < NX; ++r) col[r] = b.sx0_in[((size_t)n * nplist + il) * NX + r] / out.pcoef[il];
                for (int r = 0; r < NX; ++r) bdf.znQ[0][il * NX + r] = col[r];
            }
            if (!consistent_start_sens(dp, o.t0, x0, bdf.znQ[0], nplist)) status = AB_IC_FAIL;
            if (b.sx0 && out.full())
                for (int il = 0; il < nplist; ++il)
                    for (int r = 0; r < NX; ++r) {
                        const double v = bdf.znQ[0][il * NX + r];
                        b.sx0[((size_t)n * nplist + il) * NX + r] = v != 0.0 ? v * out.pcoef[il] : v;
                    }
        }
        if (status == AB_SUCCESS) {
            for (int r = 0; r < NX; ++r) bdf.zn[0][r] = x0[r];
            bdf.init(o.t0);
            double t = o.t0, t_done = o.t0;
            while (it < nt) {
                double tout = b.ts[it];
                if (!(tout > o.t0) || !(t < tout) || adj_too_close(t, tout)) {
                    const int counters[5] = {bdf.nst, bdf.nfe, bdf.netf, bdf.ncfn, bdf.qu};
                    out.datapoint(it, bdf.zn[0], counters, S ? bdf.znQ[0] : nullptr);
                    ++it;
                    continue;
                }
                double t_event = INFINITY;
                if (M::NE > 0) {
                    t_event = next_trigger(out.c, t_done, b.ts[nt - 1]);
                    tout = ab_min(tout, t_event);
                }
                const int flag = bdf.solve_to(tout, false);
                if (flag < 0) { status = flag; break; }
                t = bdf.tretlast;
                if (M::NE > 0 && t == t_event) {
                    // every switch whose trigger time is t flips in the direction of its crossing (Simulation::collect_roots:
                    // the sign of the root before the trigger), h += roots_found, CVodeSolver::reinit semantics for the solver
                    for (int ie = 0; ie < M::NE; ++ie)
                        if (M::explicit_time(ie, out.c.p, out.c.k) == t) out.c.h[ie] += (copysign(1.0, rootvals[ie]) < 0.0) ? 1.0 : -1.0;
                    roots_at(out.c, t, bdf.zn[0], rootvals);
                    for (int ie = 0; ie < M::NE; ++ie)   // a root that is exactly zero at its trigger keeps the sign it is heading to
                        if (rootvals[ie] == 0.0) rootvals[ie] = out.c.h[ie] > 0.0 ? 0.0 : -0.0;
                    t_done = t;
                    bdf.flush_pending();
                    bdf.amici_reinit(t);
                }
            }
        }
        if (status != AB_SUCCESS) out.invalidate(it);
        out.finish(status);
    }
};

}  // namespace amici_b200
#endif
