// adjoint.cuh — adjoint sensitivity analysis (SensitivityMethod::adjoint, reference src/backwardproblem.cpp,
// src/model_ode.cpp:323-374, ThirdParty/sundials/src/cvodes/cvodea.c) for models without events: ONE THREAD PER CONDITION.
//
//   forward pass   the state alone through the CVODES-equivalent BDF (BdfControl, bdf_core.cuh); every internal step is
//                  stored as a data point (t, order, y) of the forward history (CVodeF + CVApolynomialStorePnt), the
//                  objective and dJ/dx are formed at the output times, sllh collects the direct parameter dependence;
//   backward pass  xB' = -J(t, x(t))^T xB with the quadratures qB' = -xB^T dxdot/dp, integrated from the last output
//                  time to t0 by the SAME step control (negative step sizes; quadratures with their own error test),
//                  x(t) from the variable-degree Newton interpolation of the stored points (CVApolynomialGetY), jumps
//                  xB += dJ/dx at the output times with AMICI's solver re-initialisation;
//   gradient       sllh = -(xQB + sx0^T xB) + direct terms, chain rule for the parameter scale.
//
// The work per condition is nx + nplist scalars (JAK-STAT: 26) against nx * (1 + nplist) = 162 of the forward method, and
// the conditions of a warp share nothing, so the lane-group layout has nothing to offer here: all vectors are thread-private
// (local memory, L1-resident), the history is in HBM, interleaved over the thread slots of the grid.
// ScalarBdf, the one-thread solver defined here, also carries the DAE path (dae.cuh): mass matrix, algebraic-state handling and
// the staggered sensitivity corrector are compiled in only for problems that declare ALG_MASK / SENS.
#ifndef AMICI_B200_ADJOINT_CUH
#define AMICI_B200_ADJOINT_CUH
#include "bdf_core.cuh"

namespace amici_b200 {

// is_next_t_too_close (reference src/forwardproblem.cpp:26-38)
AB_HD bool adj_too_close(double cur_t, double t_next) {
    const double tdiff = t_next - cur_t;
    if (tdiff == 0.0) return true;
    return fabs(tdiff) < 2.0 * DBL_EPSILON * ab_max(fabs(cur_t), fabs(t_next));
}

// the large pieces are real functions: the forward and the backward solver instantiate the step control twice, and a thread's
// state is in local memory anyway
#if defined(AB_NI)
// a translation unit with one small problem (tu/robertson.cu: 3 states) may ask for inlined pieces: loops over NX unroll, the
// solver state moves from local memory into registers
#elif defined(__CUDACC__)
#define AB_NI __host__ __device__ __noinline__
#else
#define AB_NI __attribute__((noinline))
#endif

// DAE models (emitter: ALG_MASK): bit r set = state r is algebraic, row r of the right-hand side is 0 = g_r(t, x)
template <class M, class = void>
struct alg_mask_of { static constexpr unsigned long long value = 0ull; };
template <class M>
struct alg_mask_of<M, decltype((void)M::ALG_MASK)> { static constexpr unsigned long long value = M::ALG_MASK; };

struct ScalarShared {
    double ctl[16];
    double ssdat[5][8];
};

// ---------------------------------------------------------------------------------------------------------------------
// BdfControl's vector primitives for one thread and NX ODE rows + NQ quadrature rows.  P supplies
//   bool rhs(t, y, ydot)   void jac(t, y, J[NX][NX])   bool quad(t, y, qdot)   double rtol, atol, rtolq, atolq
// P::ALG_MASK != 0 (dae.cuh): index-1 DAE M y' = f(t, y) with M = diag(state is differential).  The BDF formula and the step
// control are unchanged; the corrector solves M (rl1 zn[1] + acor) - gamma f(zn[0] + acor) = 0 with the Newton matrix
// M - gamma J (algebraic rows divided by gamma: 0 = g with the matrix row -dg/dx), the algebraic states carry no derivative at the start and are left out of every norm of the local error test
// and the order / step-size selection (IDASetSuppressAlg(true), reference src/solver.cpp:252) but not of the Newton convergence test.
// ---------------------------------------------------------------------------------------------------------------------
template <class P>
struct ScalarBdf : BdfControl<ScalarBdf<P>, ScalarShared> {
    static constexpr int NX = P::NX, NQ = P::NQ, NQA = P::NQ > 0 ? P::NQ : 1;
    static constexpr bool ROOTS = false, EVENTS = true, STG = false, QUAD = P::NQ > 0, HIST = P::HIST;
    static constexpr unsigned long long AM = P::ALG_MASK;
    // P::SENS (dae.cuh): the NQ "quadrature" rows are forward sensitivities, column il in rows il * NX .. il * NX + NX - 1; they
    // are corrected after the states by a Newton iteration on their own linear system with the states' factors (the staggered
    // corrector, CV_STAGGERED) instead of cvQuadNls' explicit formula, and take the error test of the quadratures (errconS)
    static constexpr bool SENS = P::SENS;
    AB_HD static bool alg(int r) { return AM != 0ull && ((AM >> r) & 1ull) != 0ull; }
    AB_HD static bool algq(int r) { return SENS && alg(r % NX); }
    using Base = BdfControl<ScalarBdf<P>, ScalarShared>;
    using Base::o; using Base::tn; using Base::h; using Base::l; using Base::tq; using Base::rl1; using Base::gamma;
    using Base::gammap; using Base::gamrat; using Base::crate; using Base::delp; using Base::acnrm; using Base::q; using Base::L;
    using Base::eta; using Base::nst; using Base::nstlp; using Base::nstlj; using Base::nfe; using Base::nni; using Base::nsetups;
    using Base::nje; using Base::convfail; using Base::jcur; using Base::pending;

    ScalarShared sh;
    P& prob;
    int nq;   // active quadrature rows (nplist)
    double zn[LMAX][NX], acor[NX], ewt[NX];
    double ewt_d[AM != 0ull ? NX : 1];   // DAE: ewt with zeros for the algebraic states (error test, order selection)
    AB_HD const double* we() const { return AM != 0ull ? ewt_d : ewt; }
    double znQ[LMAX][NQA], acorQ[NQA], ewtQ[NQA];
    double lu[NX][NX], savedJ[NX][NX];
    int piv[NX];

    AB_HD ScalarBdf(P& p, const SolverOptions& opt, int nq_) : Base(sh, opt), prob(p), nq(nq_) {}

    AB_HD int nx() const { return NX; }
    AB_HD int sldet() const { return Base::sldet_serial(); }
    AB_HD void gsync() const {}
    AB_HD bool leader() const { return true; }
    AB_HD bool time_exceeded() const { return false; }
    AB_HD void step_sync_join() {}
    AB_HD void step_sync_leave() {}
    AB_HD void step_sync_arrive() {}
    AB_HD void on_step() { prob.store_point(tn, Base::qu, zn[0], pending ? acor : nullptr); }

    AB_HD static double upd(double m, double v) { return (v > m || v != v) ? v : m; }
    AB_NI double wrms2(const double* a, const double* b, double fb, const double* w, int n) const {
        double s = 0.0;
        for (int r = 0; r < n; ++r) {
            const double v = (b ? a[r] + fb * b[r] : a[r]) * w[r];
            s += v * v;
        }
        return n > 0 ? sqrt(s / n) : 0.0;
    }
    AB_HD double wrms_max_zn(int j, const double lj = 0.0) const {
        double m = wrms2(zn[j], pending ? acor : nullptr, lj, we(), NX);
        if (QUAD) m = upd(m, wrms2(znQ[j], pending ? acorQ : nullptr, lj, ewtQ, nq));   // cvQuadUpdateNorm
        return m;
    }
    AB_HD double wrms_max_up(const double cquot) const {
        double m = wrms2(acor, zn[QMAX], -cquot, we(), NX);
        if (QUAD) m = upd(m, wrms2(acorQ, znQ[QMAX], -cquot, ewtQ, nq));
        return m;
    }
    AB_NI void set_ewt() {
        for (int r = 0; r < NX; ++r) ewt[r] = 1.0 / (prob.rtol * fabs(pending ? zn[0][r] + acor[r] : zn[0][r]) + prob.atol);
        if (AM != 0ull)
            for (int r = 0; r < NX; ++r) ewt_d[r] = alg(r) ? 0.0 : ewt[r];
        if (QUAD)
            for (int r = 0; r < nq; ++r) ewtQ[r] = algq(r) ? 0.0 : 1.0 / (prob.rtolq * fabs(pending ? znQ[0][r] + acorQ[r] : znQ[0][r]) + prob.atolq);
    }
    AB_HD void stab_sums(const double lq, const double lqm1, double& s1, double& s2) const {
        for (int r = 0; r < NX; ++r) {
            const double a = (zn[q][r] + lq * acor[r]) * we()[r], b = (zn[q - 1][r] + lqm1 * acor[r]) * we()[r];
            s1 += a * a;
            s2 += b * b;
        }
    }
    // cvCompleteStep / cvRescale / cvPredict / cvRestore in one pass per row (see Bdf::history_pass)
    template <int PRED, int N>
    AB_NI void pass(double (&z)[LMAX][N], const double* ac, int n, const bool cmpl, const bool resc) {
        double f[QMAX + 1], lc[QMAX + 1];
        f[0] = 1.0;
        AB_UNROLL
        for (int j = 1; j <= QMAX; ++j) f[j] = resc ? f[j - 1] * eta : 1.0;
        AB_UNROLL
        for (int j = 0; j <= QMAX; ++j) lc[j] = (cmpl && j <= q && j < LMAX) ? l[j < LMAX ? j : 0] : 0.0;
        AB_NOUNROLL
        for (int r = 0; r < n; ++r) {
            double v[QMAX + 1];
            const double a = cmpl ? ac[r] : 0.0;
            // DAE: the algebraic states are outside the error test, so nothing bounds their corrections, and a Nordsieck array
            // of order q fed with unbounded corrections drifts (observed: predictions of the algebraic states oscillating with
            // growing amplitude until the corrector stalls).  Their history is therefore kept at order 1 whatever the order of
            // the method — zn[1] = last backward difference, i.e. l = (1, 1, 0, ...): a linear extrapolation as predictor; the
            // corrector puts them on the constraint.
            const bool arow = ((const void*)&z[0][0] == (const void*)&zn[0][0]) ? alg(r) : algq(r);
            AB_UNROLL
            for (int j = 0; j <= QMAX; ++j)
                v[j] = arow ? (j <= 1 ? z[j][r] + a : 0.0) : ((j <= 1 || j <= q) ? z[j][r] : 0.0) + lc[j] * a;
            AB_UNROLL
            for (int j = 1; j <= QMAX; ++j) v[j] *= f[j];
            if (PRED != 0) {
                AB_UNROLL
                for (int k = 1; k <= QMAX; ++k) {
                    AB_UNROLL
                    for (int j = QMAX; j >= k; --j) v[j - 1] = PRED > 0 ? v[j - 1] + v[j] : v[j - 1] - v[j];
                }
            }
            AB_UNROLL
            for (int j = 0; j <= QMAX; ++j)
                if (j <= 1 || j <= q) z[j][r] = v[j];
        }
    }
    template <int PRED>
    AB_HD void history_pass(const bool cmpl, const bool resc) {
        pass<PRED, NX>(zn, acor, NX, cmpl, resc);
        if (QUAD) pass<PRED, NQA>(znQ, acorQ, nq, cmpl, resc);
    }
    template <int N>
    AB_HD void up(double (&z)[LMAX][N], int n, const double A1, const double (&ll)[LMAX + 1]) {
        for (int r = 0; r < n; ++r) {
            const double zL = A1 * z[QMAX][r];
            for (int j = 2; j <= QMAX; ++j) {
                if (j == L) z[j][r] = zL;
                else if (j <= q) z[j][r] += ll[j] * zL;
            }
        }
    }
    template <int N>
    AB_HD void down(double (&z)[LMAX][N], int n, const double (&ll)[LMAX + 1]) {
        for (int r = 0; r < n; ++r) {
            const double v = z[q][r];
            for (int j = 2; j < q; ++j) z[j][r] -= ll[j] * v;
        }
    }
    AB_HD void order_up(const double A1, const double (&ll)[LMAX + 1]) {
        up<NX>(zn, NX, A1, ll);
        if (QUAD) up<NQA>(znQ, nq, A1, ll);
    }
    AB_HD void order_down(const double (&ll)[LMAX + 1]) {
        down<NX>(zn, NX, ll);
        if (QUAD) down<NQA>(znQ, nq, ll);
    }
    AB_HD void store_correction(const bool save) {
        if (save) store_correction_high();
    }
    AB_HD void store_correction_high() {
        for (int r = 0; r < NX; ++r) zn[QMAX][r] = acor[r];
        if (QUAD)
            for (int r = 0; r < nq; ++r) znQ[QMAX][r] = acorQ[r];
    }
    AB_HD void scale_zn1(double f) {
        for (int r = 0; r < NX; ++r) zn[1][r] *= f;
        if (QUAD)
            for (int r = 0; r < nq; ++r) znQ[1][r] *= f;
    }
    AB_HD void init_vectors() {
        for (int r = 0; r < NX; ++r) acor[r] = ewt[r] = 0.0;
        for (int r = 0; r < NQA; ++r) acorQ[r] = ewtQ[r] = 0.0;
        for (int j = 1; j < LMAX; ++j) {
            for (int r = 0; r < NX; ++r) zn[j][r] = 0.0;
            for (int r = 0; r < NQA; ++r) znQ[j][r] = 0.0;
        }
    }
    AB_HD bool rhs_to_zn1(double scale) {
        double f[NX];
        bool ok = prob.rhs(tn, zn[0], f);
        for (int r = 0; r < NX; ++r) zn[1][r] = alg(r) ? 0.0 : scale * f[r];
        if (QUAD) {
            double fq[NQA];
            ok = prob.quad(tn, zn[0], fq) && ok;
            for (int r = 0; r < nq; ++r) znQ[1][r] = algq(r) ? 0.0 : scale * fq[r];
        }
        return ok;
    }
    // cvHin / cvUpperBoundH0 / cvYddNorm (cvodes.c:5436-5798) with the quadrature terms of errconQ; zn[1] = y'(tn) unscaled
    AB_NI int initial_step(double tout, double* h_out, int* nfe_hin) {
        const double tdiff = tout - tn;
        if (tdiff == 0.0) return AB_TOO_CLOSE;
        const double sign = tdiff > 0.0 ? 1.0 : -1.0;
        const double tdist = fabs(tdiff);
        const double tround = DBL_EPSILON * ab_max(fabs(tn), fabs(tout));
        if (tdist < 2.0 * tround) return AB_TOO_CLOSE;
        const double hlb = 100.0 * tround;
        double hub_inv = 0.0;
        for (int r = 0; r < NX; ++r)
            if (!alg(r)) hub_inv = upd(hub_inv, fabs(zn[1][r]) / (0.1 * fabs(zn[0][r]) + 1.0 / ewt[r]));
        if (QUAD)
            for (int r = 0; r < nq; ++r) hub_inv = upd(hub_inv, fabs(znQ[1][r]) / (0.1 * fabs(znQ[0][r]) + 1.0 / ewtQ[r]));
        double hub = 0.1 * tdist;
        if (hub * hub_inv > 1.0) hub = 1.0 / hub_inv;
        double hg = sqrt(hlb * hub);
        if (hub < hlb) {
            *h_out = sign * hg;
            return AB_SUCCESS;
        }
        double hs = hg, hnew = hg;
        for (int count1 = 1; count1 <= 4; count1++) {
            bool hgOK = false;
            double yddnrm = 0.0;
            for (int count2 = 1; count2 <= 4; count2++) {
                const double hgs = hg * sign;
                double y[NX], f[NX], fq[NQA];
                for (int r = 0; r < NX; ++r) y[r] = hgs * zn[1][r] + zn[0][r];
                bool ok = prob.rhs(tn + hgs, y, f);
                (*nfe_hin)++;
                if (QUAD) ok = prob.quad(tn + hgs, y, fq) && ok;
                if (ok) {
                    for (int r = 0; r < NX; ++r) f[r] = (1.0 / hgs) * f[r] - (1.0 / hgs) * zn[1][r];
                    yddnrm = wrms2(f, nullptr, 0.0, we(), NX);
                    if (QUAD) {
                        for (int r = 0; r < nq; ++r) fq[r] = (1.0 / hgs) * fq[r] - (1.0 / hgs) * znQ[1][r];
                        yddnrm = upd(yddnrm, wrms2(fq, nullptr, 0.0, ewtQ, nq));
                    }
                    hgOK = true;
                    break;
                }
                hg *= 0.2;
            }
            if (!hgOK) {
                if (count1 <= 2) return AB_REPTD_RHSFUNC_ERR;
                hnew = hs;
                break;
            }
            hs = hg;
            hnew = (yddnrm * hub * hub > 2.0) ? sqrt(2.0 / yddnrm) : sqrt(hg * hub);
            if (count1 == 4) break;
            const double hrat = hnew / hg;
            if (hrat > 0.5 && hrat < 2.0) break;
            if (count1 > 1 && hrat > 2.0) {
                hnew = hg;
                break;
            }
            hg = hnew;
        }
        double h0 = 0.5 * hnew;
        if (h0 < hlb) h0 = hlb;
        if (h0 > hub) h0 = hub;
        *h_out = sign * h0;
        return AB_SUCCESS;
    }
    // cvLsSetup + cvLsLinSys (cvodes_ls.c:1611-1710, :1386-1470) with a dense LU with partial pivoting (SUNLinSol_Dense semantics)
    AB_NI int lsetup(bool jbad_nls, const double* ycur) {
        if (jbad_nls) convfail = CONVFAIL_BAD_J;
        const double dgamma = fabs(gamma / gammap - 1.0);
        const bool jbad = (nst == 0) || (nst >= nstlj + 51) || (convfail == CONVFAIL_BAD_J && dgamma < 0.2) || (convfail == CONVFAIL_OTHER);
        if (jbad) {
            jcur = true;
            prob.jac(tn, ycur, savedJ);
            nje++;
            nstlj = nst;
        } else {
            jcur = false;
        }
        bool ok = true;
        for (int r = 0; r < NX; ++r)
            for (int c = 0; c < NX; ++c) {
                ok = ok && isfinite(savedJ[r][c]);
                // algebraic rows without gamma (as in IDAS, where cj only multiplies dF/dy'): a stale gamma does not touch them
                lu[r][c] = alg(r) ? -savedJ[r][c] : (r == c ? 1.0 : 0.0) - gamma * savedJ[r][c];
            }
        for (int kk = 0; kk < NX && ok; ++kk) {
            int pr = kk;
            double best = fabs(lu[kk][kk]);
            for (int r = kk + 1; r < NX; ++r)
                if (fabs(lu[r][kk]) > best) { best = fabs(lu[r][kk]); pr = r; }
            piv[kk] = pr;
            if (best == 0.0) { ok = false; break; }
            if (pr != kk)
                for (int c = 0; c < NX; ++c) { const double tmp = lu[kk][c]; lu[kk][c] = lu[pr][c]; lu[pr][c] = tmp; }
            const double inv = 1.0 / lu[kk][kk];
            for (int r = kk + 1; r < NX; ++r) {
                const double m = lu[r][kk] * inv;
                lu[r][kk] = m;
                if (m != 0.0)
                    for (int c = kk + 1; c < NX; ++c) lu[r][c] -= m * lu[kk][c];
            }
        }
        nsetups++;
        gamrat = 1.0;
        gammap = gamma;
        crate = 1.0;
        nstlp = nst;
        return ok ? AB_SUCCESS : NLS_CONV_RECVR;
    }
    AB_HD void lu_solve(double* b) const {
        // SUNDIALS' denseGETRS: ALL row interchanges first (lsetup swaps whole rows, so the multipliers of a column sit where
        // the later interchanges have moved them), then L y = b, then U x = y
        AB_NOUNROLL
        for (int kk = 0; kk < NX; ++kk) {
            const int pr = piv[kk];
            if (pr != kk) { const double tmp = b[kk]; b[kk] = b[pr]; b[pr] = tmp; }
        }
        AB_NOUNROLL
        for (int kk = 0; kk < NX; ++kk) {
            const double bk = b[kk];
            for (int r = kk + 1; r < NX; ++r) b[r] -= lu[r][kk] * bk;
        }
        AB_NOUNROLL
        for (int r = NX - 1; r >= 0; --r) {
            double s = b[r];
            for (int c = r + 1; c < NX; ++c) s -= lu[r][c] * b[c];
            b[r] = s / lu[r][r];
        }
    }
    // cvNls + SUNNonlinSolSolve_Newton (see Bdf::nls_impl)
    AB_NI int nls(int nflag) {
        convfail = (nflag == FIRST_CALL || nflag == PREV_ERR_FAIL) ? CONVFAIL_NONE : CONVFAIL_OTHER;
        bool callSetup = (nflag == PREV_CONV_FAIL) || (nflag == PREV_ERR_FAIL) || (nst == 0) || (nst >= nstlp + 20) ||
                         (fabs(gamrat - 1.0) > 0.3);
        for (int r = 0; r < NX; ++r) acor[r] = 0.0;
        bool jbad = false;
        int ret = AB_SUCCESS;
        // DAE: the convergence rate of earlier steps says nothing about the first update of the algebraic states (linear
        // extrapolation as predictor): every step earns its own estimate, i.e. a first update above the tolerance is followed
        // by a second iteration
        if (AM != 0ull) crate = 1.0;
        for (;;) {
            int m = 0;
            for (;;) {
                double ycur[NX], delta[NX];
                for (int r = 0; r < NX; ++r) ycur[r] = zn[0][r] + acor[r];
                const bool ok = prob.rhs(tn, ycur, delta);
                nfe++;
                if (!ok) { ret = NLS_RHS_RECVR; break; }
                if (m == 0 && callSetup) {
                    ret = lsetup(jbad, ycur);
                    if (ret != AB_SUCCESS) break;
                }
                nni++;
                for (int r = 0; r < NX; ++r) delta[r] = alg(r) ? delta[r] : -(rl1 * zn[1][r] + acor[r] - gamma * delta[r]);
                lu_solve(delta);
                const double sc = (gamrat != 1.0) ? 2.0 / (1.0 + gamrat) : 1.0;
                for (int r = 0; r < NX; ++r) {
                    if (gamrat != 1.0) delta[r] *= sc;
                    acor[r] += delta[r];
                }
                const double Del = wrms2(delta, nullptr, 0.0, ewt, NX);
                if (m > 0) crate = ab_max(0.3 * crate, Del / delp);
                // DAE: ten times the ODE's corrector accuracy — what is left of the first update of an algebraic state is an error
                // no error test sees (measured on 512 Robertson draws: constraint violation 2.1e-8 -> 2.2e-9, IDAS 2.0e-9)
                const double dcon = Del * ab_min(1.0, crate) / (AM != 0ull ? 0.1 * tq[4] : tq[4]);
                if (dcon <= 1.0) {
                    acnrm = (m == 0 && AM == 0ull) ? Del : wrms2(acor, nullptr, 0.0, we(), NX);
                    jcur = false;
                    return AB_SUCCESS;
                }
                if (!(Del == Del) || (m >= 1 && Del > 2.0 * delp)) { ret = NLS_CONV_RECVR; break; }
                delp = Del;
                m++;
                if (m >= o.max_nonlin_iters) { ret = NLS_CONV_RECVR; break; }
            }
            if (ret > 0 && !jcur) {
                callSetup = true;
                jbad = true;
                for (int r = 0; r < NX; ++r) acor[r] = 0.0;
                continue;
            }
            break;
        }
        return ret;
    }
    // cvQuadNls (cvodes.c:7236-7262)
    AB_NI bool quad_correct() {
        if (SENS) return sens_correct();
        double y[NX], fq[NQA];
        for (int r = 0; r < NX; ++r) y[r] = zn[0][r] + acor[r];
        const bool ok = prob.quad(tn, y, fq);
        for (int r = 0; r < nq; ++r) acorQ[r] = rl1 * (h * fq[r] - znQ[1][r]);
        return ok;
    }
    AB_HD double quad_norm() const { return wrms2(acorQ, nullptr, 0.0, ewtQ, nq); }
    // Staggered corrector of the sensitivities (cvStgrNls, cvodes_nls_stg.c) at the converged state: the system is linear in the
    // sensitivities, M (rl1 znS[1] + acorS) - gamma (J s + df/dp) = 0 column by column, solved with the factors of the states'
    // last setup (M - gammap J): a fixed-point iteration that contracts with |1 - gamma / gammap| and the drift of J
    AB_NI bool sens_correct() {
        double y[NX], sd[NQA];
        for (int r = 0; r < NX; ++r) y[r] = zn[0][r] + acor[r];
        for (int r = 0; r < nq; ++r) acorQ[r] = 0.0;
        double delp_s = 0.0, crate_s = 1.0;
        for (int m = 0;; ++m) {
            for (int r = 0; r < nq; ++r) sd[r] = znQ[0][r] + acorQ[r];
            if (!prob.sens_rhs(tn, y, sd, sd)) return false;   // sd <- J s + df/dp (algebraic rows: dg/dx s + dg/dp)
            double Del = 0.0;
            for (int il = 0; il * NX < nq; ++il) {
                double d[NX];
                for (int r = 0; r < NX; ++r) {
                    const int i = il * NX + r;
                    d[r] = alg(r) ? sd[i] : -(rl1 * znQ[1][i] + acorQ[i] - gamma * sd[i]);
                }
                lu_solve(d);
                const double sc = (gamrat != 1.0) ? 2.0 / (1.0 + gamrat) : 1.0;
                double s2 = 0.0;
                for (int r = 0; r < NX; ++r) {
                    const int i = il * NX + r;
                    d[r] *= sc;
                    acorQ[i] += d[r];
                    const double w = d[r] / (prob.rtolq * fabs(znQ[0][i]) + prob.atolq);
                    s2 += w * w;
                }
                Del = upd(Del, sqrt(s2 / NX));
            }
            if (m > 0) crate_s = ab_max(0.3 * crate_s, Del / delp_s);
            if (Del * ab_min(1.0, crate_s) / (0.1 * tq[4]) <= 1.0) return true;
            if (!(Del == Del) || (m >= 1 && Del > 2.0 * delp_s) || m + 1 >= 4) return false;
            delp_s = Del;
        }
    }
    // CVodeGetDky(t, 0) / CVodeGetQuadDky(t, 0) on the completed history
    AB_NI void interpolate(double t, double* y, double* yq) {
        Base::flush_pending();
        const double s = (t - tn) / h;
        for (int r = 0; r < NX; ++r) {
            double v = 0.0;
            for (int j = q; j >= 0; --j) v = v * s + zn[j][r];
            y[r] = v;
        }
        for (int r = 0; r < (QUAD ? nq : 0); ++r) {
            double v = 0.0;
            for (int j = q; j >= 0; --j) v = v * s + znQ[j][r];
            yq[r] = v;
        }
    }
};

// ---------------------------------------------------------------------------------------------------------------------
// forward history: the data points of CVodeF and the interpolation of CVApolynomialGetY (cvodea.c:2548-2640, :3236-3400)
// ---------------------------------------------------------------------------------------------------------------------
template <int NX>
struct FwdHistory {
    double* t;       // t[i * stride]
    double* y;       // y[(i * NX + r) * stride]
    int* ord;        // ord[i * stride]: order used by the step that ended at point i (CVApolynomialStorePnt: cv_qu)
    size_t stride;   // element stride: the slots of all threads are interleaved
    int cap, np;
    bool overflow;
    // interpolation state
    int ilast, order;
    bool new_data;
    double dt, T[QMAX + 1], Y[QMAX + 1][NX];

    AB_HD void reset() { np = 0; overflow = false; new_data = true; ilast = -1; order = 0; }
    AB_HD double tp(int i) const { return t[(size_t)i * stride]; }
    AB_HD void store(double tt, int qu, const double* z0, const double* ac) {
        if (np >= cap) { overflow = true; return; }
        t[(size_t)np * stride] = tt;
        ord[(size_t)np * stride] = qu;
        for (int r = 0; r < NX; ++r) y[((size_t)np * NX + r) * stride] = ac ? z0[r] + ac[r] : z0[r];
        np++;
        new_data = true;
    }
    // y(t) by the Newton polynomial through order + 1 stored points ending at the right end of the step that contains t
    AB_NI bool get(double tt, double* out) {
        bool newpoint = false;
        if (new_data) {
            ilast = np - 1;
            newpoint = true;
            new_data = false;
        }
        if (np < 2 || ilast < 1) {
            for (int r = 0; r < NX; ++r) out[r] = y[(size_t)r * stride];
            return np >= 1;
        }
        int indx;
        const bool to_left = (tt - tp(ilast - 1)) < 0.0, to_right = (tt - tp(ilast)) > 0.0;
        if (to_left) {
            newpoint = true;
            indx = ilast;
            for (;;) {
                if (indx == 0) break;
                if (tt - tp(indx - 1) <= 0.0) indx--;
                else break;
            }
            ilast = indx == 0 ? 1 : indx;
            if (indx == 0 && fabs(tt - tp(0)) > 1e6 * DBL_EPSILON) return false;   // CV_GETY_BADT
        } else if (to_right) {
            newpoint = true;
            indx = ilast;
            for (;;) {
                if (indx < np - 1 && tt - tp(indx) > 0.0) indx++;
                else break;
            }
            if (tt - tp(indx) > 1e6 * DBL_EPSILON * (1.0 + fabs(tt))) return false;
            ilast = indx;
        } else {
            indx = ilast;
        }
        if (indx == 0) {
            for (int r = 0; r < NX; ++r) out[r] = y[(size_t)r * stride];
            return true;
        }
        if (newpoint) {
            dt = fabs(tp(indx) - tp(indx - 1));
            int base = indx;
            order = ord[(size_t)base * stride];
            if (order > QMAX) order = QMAX;
            if (indx < order) base += order - indx;
            if (base > np - 1) { order -= base - (np - 1); base = np - 1; }
            // the divided differences row by row in registers; the factors dt / (T[j] - T[j - i]) are the same for every row
            double Tl[QMAX + 1], fac[QMAX + 1][QMAX + 1];
            AB_UNROLL
            for (int j = 0; j <= QMAX; ++j) Tl[j] = j <= order ? tp(base - j) : 0.0;
            AB_UNROLL
            for (int j = 0; j <= QMAX; ++j) T[j] = Tl[j];
            AB_UNROLL
            for (int i = 1; i <= QMAX; ++i) {
                AB_UNROLL
                for (int j = QMAX; j >= i; --j) fac[i][j] = j <= order ? dt / (Tl[j] - Tl[j - i]) : 0.0;
            }
            AB_NOUNROLL
            for (int r = 0; r < NX; ++r) {
                double yy[QMAX + 1];
                AB_UNROLL
                for (int j = 0; j <= QMAX; ++j) yy[j] = j <= order ? y[((size_t)(base - j) * NX + r) * stride] : 0.0;
                AB_UNROLL
                for (int i = 1; i <= QMAX; ++i) {
                    AB_UNROLL
                    for (int j = QMAX; j >= i; --j)
                        if (j <= order) yy[j] = fac[i][j] * yy[j] - fac[i][j] * yy[j - 1];
                }
                AB_UNROLL
                for (int j = 0; j <= QMAX; ++j) Y[j][r] = yy[j];
            }
        }
        double cv[QMAX + 1];
        cv[0] = 1.0;
        AB_UNROLL
        for (int i = 0; i < QMAX; ++i) cv[i + 1] = i < order ? cv[i] * (tt - T[i]) / dt : 0.0;
        AB_UNROLL
        for (int r = 0; r < NX; ++r) {
            double v = 0.0;
            AB_UNROLL
            for (int i = 0; i <= QMAX; ++i) v += cv[i] * Y[i][r];   // rows above the order hold zeros
            out[r] = v;
        }
        return true;
    }
};

// what both problems of a condition see
template <class M>
struct AdjCond {
    static constexpr int NSPLV = M::NSPL > 0 ? M::NSPL : 1;
    double p[M::NP > 0 ? M::NP : 1];
    double k[M::NK_EXT > 0 ? M::NK_EXT : 1];
    double h[M::NH > 0 ? M::NH : 1];
    SplineTable<M, M::NP + 1> spl;   // column 0: value polynomial, column 1 + il: sensitivity w.r.t. plist[il]
    const int* plist;
    int nplist;
    AB_HD void splines(double t, double* v) const {
        for (int is = 0; is < M::NSPL; ++is) {
            const SplineLoc loc = spline_locate<M>(is, t);
            const double u = spline_eval_at<M, M::NP + 1>(spl, is, 0, loc, true);
            v[is] = spline_is_log<M>(is) ? exp(u) : u;
        }
    }
    AB_HD void spline_sens(double t, int il, const double* v, double* sv) const {
        for (int is = 0; is < M::NSPL; ++is) {
            const SplineLoc loc = spline_locate<M>(is, t);
            const double u = spline_eval_at<M, M::NP + 1>(spl, is, 1 + il, loc, false);
            sv[is] = spline_is_log<M>(is) ? u * v[is] : u;
        }
    }
    // the same with the interval located once for all parameters
    AB_HD void spline_sens_at(const SplineLoc* loc, int il, const double* v, double* sv) const {
        for (int is = 0; is < M::NSPL; ++is) {
            const double u = spline_eval_at<M, M::NP + 1>(spl, is, 1 + il, loc[is], false);
            sv[is] = spline_is_log<M>(is) ? u * v[is] : u;
        }
    }
};

template <class M>
struct FwdProblem {
    static constexpr int NX = M::NX, NQ = 0;
    static constexpr bool HIST = true, SENS = false;
    AB_HD bool sens_rhs(double, const double*, const double*, double*) const { return true; }
    static constexpr unsigned long long ALG_MASK = 0ull;
    const AdjCond<M>& c;
    FwdHistory<M::NX>& hist;
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
    AB_HD bool quad(double, const double*, double*) const { return true; }
    AB_HD void store_point(double t, int qu, const double* z0, const double* ac) { hist.store(t, qu, z0, ac); }
};

// xB' = -J(t, x(t))^T xB, qB'[il] = -xB^T dxdot/dp_plist[il] (Model_ODE::fxBdot / fqBdot / fJSparseB, reference src/model_ode.cpp:323-374)
template <class M>
struct BwdProblem {
    static constexpr int NX = M::NX, NQ = M::NP > 0 ? M::NP : 1;
    static constexpr bool HIST = false, SENS = false;
    AB_HD bool sens_rhs(double, const double*, const double*, double*) const { return true; }
    static constexpr unsigned long long ALG_MASK = 0ull;
    const AdjCond<M>& c;
    FwdHistory<M::NX>& hist;
    double rtol, atol, rtolq, atolq;
    double tc, x[NX], Jd[NX * NX], spl[AdjCond<M>::NSPLV];
    bool have;
    AB_NI bool at(double t) {   // CVArhs / CVArhsQ / the Jacobian wrapper: forward solution from the interpolation
        if (have && t == tc) return true;
        if (!hist.get(t, x)) return false;
        double Jv[M::NNZ_J > 0 ? M::NNZ_J : 1];
        c.splines(t, spl);
        M::jac(t, x, c.p, c.k, c.h, spl, Jv);
        M::jac_to_dense(Jv, Jd);
        tc = t;
        have = true;
        return true;
    }
    AB_HD bool rhs(double t, const double* xB, double* f) {
        if (!at(t)) return false;
        bool ok = true;
        for (int cc = 0; cc < NX; ++cc) {
            double s = 0.0;
            for (int r = 0; r < NX; ++r) s += Jd[r * NX + cc] * xB[r];
            f[cc] = -s;
            ok = ok && isfinite(f[cc]);
        }
        return ok;
    }
    AB_HD void jac(double t, const double*, double (&J)[NX][NX]) {
        at(t);
        for (int r = 0; r < NX; ++r)
            for (int cc = 0; cc < NX; ++cc) J[r][cc] = -Jd[cc * NX + r];
    }
    AB_NI bool quad(double t, const double* xB, double* fq) {
        if (!at(t)) return false;
        bool ok = true;
        SplineLoc loc[AdjCond<M>::NSPLV];
        for (int is = 0; is < M::NSPL; ++is) loc[is] = spline_locate<M>(is, t);
        for (int il = 0; il < c.nplist; ++il) {
            double sspl[AdjCond<M>::NSPLV], col[NX];
            c.spline_sens_at(loc, il, spl, sspl);
            M::dxdotdp(c.plist[il], t, x, c.p, c.k, c.h, spl, sspl, col);
            double s = 0.0;
            for (int r = 0; r < NX; ++r) s += col[r] * xB[r];
            fq[il] = -s;
            ok = ok && isfinite(s);
        }
        return ok;
    }
};

// ---------------------------------------------------------------------------------------------------------------------
// one condition: ForwardProblem::run + BackwardProblem::workBackwardProblem + ReturnData::process_backward_problem
// (reference src/forwardproblem.cpp:53-151, src/backwardproblem.cpp:27-65, :215-330, src/rdata.cpp:514-629)
// ---------------------------------------------------------------------------------------------------------------------
template <class M>
struct AdjointSimulation {
    static constexpr int NX = M::NX, NY = M::NY, NPA = M::NP > 0 ? M::NP : 1;
    const SolverOptions& o;
    const BatchView& b;
    FwdHistory<NX>& hist;
    double* dJydx;   // scratch of this thread: [nt][NX], element stride hist.stride
    int n;
    AdjCond<M> c;
    double pcoef[NPA], sllh[NPA], llh, chi2;
    bool fsa = false;   // forward sensitivities are reported (dae.cuh): a failure also invalidates sres and the FIM

    AB_HD AdjointSimulation(const SolverOptions& opt, const BatchView& bv, FwdHistory<NX>& hs, double* dj, int n_)
        : o(opt), b(bv), hist(hs), dJydx(dj), n(n_) {}
    AB_HD bool full() const { return o.reporting != 2; }
    AB_HD static double nan_() {
#ifdef __CUDA_ARCH__
        return __longlong_as_double(0x7ff8000000000000LL);
#else
        return NAN;
#endif
    }
    AB_HD double& dj(int it, int r) { return dJydx[((size_t)it * NX + r) * hist.stride]; }

    AB_NI void load() {
        for (int i = 0; i < M::NP; ++i) {
            const double v = b.p[(size_t)n * M::NP + i];
            const int sc = b.pscale ? b.pscale[i] : AB_PSCALE_NONE;
            c.p[i] = sc == AB_PSCALE_LOG10 ? pow(10.0, v) : sc == AB_PSCALE_LN ? exp(v) : v;
        }
        for (int i = 0; i < M::NK; ++i) c.k[i] = b.k[(size_t)n * M::NK + i];
        M::derived(c.p, c.k);
        for (int i = 0; i < M::NH; ++i) c.h[i] = 0.0;
        c.plist = b.plist;
        c.nplist = b.nplist;
        for (int is = 0; is < M::NSPL; ++is) spline_build_column<M, M::NP + 1>(c.spl, is, 0, -1, c.p, c.k);
        for (int il = 0; il < c.nplist; ++il) {
            const int ip = b.plist[il];
            const int sc = b.pscale ? b.pscale[ip] : AB_PSCALE_NONE;
            pcoef[il] = sc == AB_PSCALE_LOG10 ? c.p[ip] * log(10.0) : sc == AB_PSCALE_LN ? c.p[ip] : 1.0;
            sllh[il] = 0.0;
            for (int is = 0; is < M::NSPL; ++is) spline_build_column<M, M::NP + 1>(c.spl, is, 1 + il, ip, c.p, c.k);
        }
        llh = chi2 = 0.0;
    }

    // observables, objective, dJ/dx and the direct parameter dependence of the objective at output time it
    // sx != nullptr: forward sensitivities of the state (column il in sx[il * NX ..]): sx / sy / ssigmay are reported and sllh is
    // completed here (Model::get_observable_sensitivity + add_observable_objective_sensitivity), nothing is kept for a backward pass
    AB_NI void datapoint(int it, const double* xs, const int* counters, const double* sx = nullptr) {
        const int nt = b.nt;
        const double t = b.ts[it];
        double spl[AdjCond<M>::NSPLV], y[NY], sig[NY], dJy[NY], dJs[NY];
        bool sig_set[NY];
        c.splines(t, spl);
        M::y(t, xs, c.p, c.k, c.h, spl, y);
        M::sigmay(c.p, c.k, y, sig);
        if (b.x && full())
            for (int r = 0; r < NX; ++r) b.x[((size_t)n * nt + it) * NX + r] = xs[r];
        for (int iy = 0; iy < NY; ++iy) {
            const size_t idx = ((size_t)n * nt + it) * NY + iy;
            const double my = b.my ? b.my[idx] : nan_();
            const double se = b.sigmay_in ? b.sigmay_in[idx] : nan_();
            sig_set[iy] = (se == se);
            if (sig_set[iy]) sig[iy] = se;
            if (b.y && full()) b.y[idx] = y[iy];
            if (b.sigmay && full()) b.sigmay[idx] = sig[iy];
            dJy[iy] = dJs[iy] = 0.0;
            if (my == my) {
                const double r_ = y[iy] - my, s2 = sig[iy] * sig[iy];
                const double pi_ = 3.14159265358979323846;
                llh -= 0.5 * log(2.0 * pi_ * s2) + 0.5 * r_ * r_ / s2;
                const double res = -r_ / sig[iy];
                if (b.res && full()) b.res[(size_t)n * nt * NY + (size_t)it * NY + iy] = res;
                chi2 += res * res;
                dJy[iy] = r_ / s2;
                dJs[iy] = sig_set[iy] ? 0.0 : 1.0 / sig[iy] - r_ * r_ / (s2 * sig[iy]);
            }
        }
        if (counters) {
            const size_t idx = (size_t)n * nt + it;
            if (b.numsteps) b.numsteps[idx] = counters[0];
            if (b.numrhsevals) b.numrhsevals[idx] = counters[1];
            if (b.numerrtestfails) b.numerrtestfails[idx] = counters[2];
            if (b.numnonlinsolvconvfails) b.numnonlinsolvconvfails[idx] = counters[3];
            if (b.order) b.order[idx] = counters[4];
        }
        if (c.nplist == 0) return;
        if (sx) {
            double srv[NPA][NY];   // sres of this timepoint (chain rule applied), for the FIM
            for (int il = 0; il < c.nplist; ++il) {
                const int ip = c.plist[il];
                double sspl[AdjCond<M>::NSPLV], syv[NY], ssv[NY];
                c.spline_sens(t, il, spl, sspl);
                M::sy(ip, t, xs, c.p, c.k, c.h, spl, sspl, sx + il * NX, syv);
                M::ssigmay(ip, c.p, c.k, y, syv, ssv);
                const double pc = pcoef[il];
                if (b.sx && full())
                    for (int r = 0; r < NX; ++r) {
                        const double v = sx[il * NX + r];
                        b.sx[(((size_t)n * nt + it) * c.nplist + il) * NX + r] = v != 0.0 ? v * pc : v;
                    }
                double s = 0.0;
                for (int iy = 0; iy < NY; ++iy) {
                    const double ssig = sig_set[iy] ? 0.0 : ssv[iy];
                    const size_t idx = (((size_t)n * nt + it) * c.nplist + il) * NY + iy;
                    if (b.sy && full()) b.sy[idx] = syv[iy] != 0.0 ? syv[iy] * pc : syv[iy];
                    if (b.ssigmay && full()) b.ssigmay[idx] = ssig != 0.0 ? ssig * pc : ssig;
                    s += dJy[iy] * syv[iy] + dJs[iy] * ssig;
                    // sres = (-sy - ssigmay res) / sigma for set measurements (Simulation::datapoint; reference src/rdata.cpp:1060-1104)
                    const double my = b.my ? b.my[((size_t)n * nt + it) * NY + iy] : nan_();
                    srv[il][iy] = 0.0;
                    if (my == my) {
                        const double res = -(y[iy] - my) / sig[iy];
                        const double sr = (-syv[iy] - ssig * res) / sig[iy];
                        srv[il][iy] = sr != 0.0 ? sr * pc : sr;
                        if (b.sres && full()) b.sres[((size_t)n * nt * NY + (size_t)it * NY + iy) * c.nplist + il] = srv[il][iy];
                    }
                }
                sllh[il] -= s;
            }
            if (b.FIM)   // FIM += sres sres^T per set measurement (reference src/rdata.cpp:1106-1205)
                for (int iy = 0; iy < NY; ++iy) {
                    const double my = b.my ? b.my[((size_t)n * nt + it) * NY + iy] : nan_();
                    if (!(my == my)) continue;
                    for (int il = 0; il < c.nplist; ++il)
                        for (int jl = 0; jl < c.nplist; ++jl) b.FIM[((size_t)n * c.nplist + il) * c.nplist + jl] += srv[il][iy] * srv[jl][iy];
                }
            return;
        }
        // dy/dx column by column from the emitted sy = dy/dx sx + dy/dp (linear in sx), dsigmay/dy likewise from ssigmay
        double zero[NX], sy0[NY], ss0[NY], sspl[AdjCond<M>::NSPLV];
        for (int r = 0; r < NX; ++r) zero[r] = 0.0;
        for (int il = 0; il < c.nplist; ++il) {   // Model::fdJydp (reference src/model.cpp:2340-2380): sllh -= dJydp
            const int ip = c.plist[il];
            c.spline_sens(t, il, spl, sspl);
            M::sy(ip, t, xs, c.p, c.k, c.h, spl, sspl, zero, sy0);
            M::ssigmay(ip, c.p, c.k, y, sy0, ss0);
            double s = 0.0;
            for (int iy = 0; iy < NY; ++iy) s += dJy[iy] * sy0[iy] + dJs[iy] * ss0[iy];
            sllh[il] -= s;
        }
        {   // Model::fdJydx (:2382-2413)
            const int ip = c.plist[0];
            c.spline_sens(t, 0, spl, sspl);
            M::sy(ip, t, xs, c.p, c.k, c.h, spl, sspl, zero, sy0);
            M::ssigmay(ip, c.p, c.k, y, sy0, ss0);
            for (int ix = 0; ix < NX; ++ix) {
                double e[NX], sy1[NY], ss1[NY];
                for (int r = 0; r < NX; ++r) e[r] = r == ix ? 1.0 : 0.0;
                M::sy(ip, t, xs, c.p, c.k, c.h, spl, sspl, e, sy1);
                M::ssigmay(ip, c.p, c.k, y, sy1, ss1);
                double s = 0.0;
                for (int iy = 0; iy < NY; ++iy) s += dJy[iy] * (sy1[iy] - sy0[iy]) + dJs[iy] * (ss1[iy] - ss0[iy]);
                dj(it, ix) = s;
            }
        }
    }

    AB_HD void invalidate(int it_start) {
        const int nt = b.nt;
        for (int it = it_start; it < nt; ++it) {
            for (int il = 0; il < (fsa ? c.nplist : 0); ++il) {   // forward sensitivities (dae.cuh), as Simulation::invalidate
                if (b.sx && full()) for (int r = 0; r < NX; ++r) b.sx[(((size_t)n * nt + it) * c.nplist + il) * NX + r] = nan_();
                for (int iy = 0; iy < NY; ++iy) {
                    if (b.sy && full()) b.sy[(((size_t)n * nt + it) * c.nplist + il) * NY + iy] = nan_();
                    if (b.sres && full()) b.sres[((size_t)n * nt * NY + (size_t)it * NY + iy) * c.nplist + il] = nan_();
                }
            }
            if (b.x && full()) for (int r = 0; r < NX; ++r) b.x[((size_t)n * nt + it) * NX + r] = nan_();
            for (int iy = 0; iy < NY; ++iy) {
                const size_t idx = ((size_t)n * nt + it) * NY + iy;
                if (b.y && full()) b.y[idx] = nan_();
                if (b.sigmay && full()) b.sigmay[idx] = nan_();
                if (b.res && full()) b.res[(size_t)n * nt * NY + (size_t)it * NY + iy] = nan_();
            }
        }
        if (fsa && b.FIM)
            for (int i = 0; i < c.nplist * c.nplist; ++i) b.FIM[(size_t)n * c.nplist * c.nplist + i] = nan_();
        llh = chi2 = nan_();
    }

    AB_HD void finish(int status) {
        if (b.status) b.status[n] = status;
        if (b.llh) b.llh[n] = llh;
        if (b.chi2) b.chi2[n] = chi2;
        if (b.sllh)
            for (int il = 0; il < c.nplist; ++il) {
                const double v = status == AB_SUCCESS ? sllh[il] : nan_();
                b.sllh[(size_t)n * c.nplist + il] = v != 0.0 ? v * pcoef[il] : v;
            }
    }

    AB_HD void run() {
        const int nt = b.nt;
        load();
        hist.reset();
        // ---- forward ----
        FwdProblem<M> fp{c, hist, o.rtol, o.atol, o.rtol, o.atol};
        ScalarBdf<FwdProblem<M>> fwd(fp, o, 0);
        double x0[NX];
        M::x0(c.p, c.k, x0);
        if (b.x0_in)
            for (int r = 0; r < NX; ++r) x0[r] = b.x0_in[(size_t)n * NX + r];
        for (int r = 0; r < NX; ++r) fwd.zn[0][r] = x0[r];
        if (b.x0 && full())
            for (int r = 0; r < NX; ++r) b.x0[(size_t)n * NX + r] = x0[r];
        fwd.init(o.t0);
        hist.store(o.t0, 0, x0, nullptr);
        double t = o.t0;
        int status = AB_SUCCESS, it = 0;
        for (; it < nt; ++it) {
            const double tout = b.ts[it];
            if (tout > o.t0 && t < tout && !adj_too_close(t, tout)) {
                const int flag = fwd.solve_to(tout, false);
                if (flag < 0) { status = flag; break; }
                if (hist.overflow) { status = AB_TOO_MUCH_WORK; break; }
                t = fwd.tretlast;
            }
            const int counters[5] = {fwd.nst, fwd.nfe, fwd.netf, fwd.ncfn, fwd.qu};
            datapoint(it, fwd.zn[0], counters);
        }
        if (status != AB_SUCCESS || c.nplist == 0 || nt == 0) {
            if (status != AB_SUCCESS) invalidate(it);
            finish(status);
            return;
        }
        // ---- backward ----
        SolverOptions ob = o;
        ob.maxsteps = o.maxsteps > 20000000 ? 2000000000 : o.maxsteps * 100;   // Solver::apply_max_num_steps_B (reference src/solver.cpp:121-139)
        BwdProblem<M> bp{c, hist, o.rtol, o.atol, o.rtol, o.atol};
        bp.have = false;
        ScalarBdf<BwdProblem<M>> bwd(bp, ob, c.nplist);
        double xB[NX], xQ[NPA];
        for (int r = 0; r < NX; ++r) xB[r] = 0.0;
        for (int r = 0; r < NPA; ++r) xQ[r] = 0.0;
        it = nt - 1;
        bool called = false;
        if (b.ts[it] == t) {   // EventHandlingBwdSimulator::run: data point at t_start
            for (int r = 0; r < NX; ++r) xB[r] += dj(it, r);
            --it;
        }
        for (int r = 0; r < NX; ++r) bwd.zn[0][r] = xB[r];
        for (int r = 0; r < NPA; ++r) bwd.znQ[0][r] = 0.0;
        bwd.init(t);   // Solver::setup_b
        for (;;) {
            const bool last = it < 0;
            const double tnext = last ? o.t0 : b.ts[it];
            if (tnext < t) {
                int flag = AB_SUCCESS;
                if (!called) {
                    // first CVodeB call: CV_NORMAL without a stop time, the solver steps past tnext and interpolates back
                    for (;;) {
                        flag = bwd.solve_to(tnext, true);
                        if (flag < 0 || (bwd.tn - tnext) * bwd.h >= 0.0) break;
                    }
                    if (flag >= 0) bwd.interpolate(tnext, xB, xQ);
                } else {
                    flag = bwd.solve_to(tnext, false);   // after reInit_post_process the stop time tnext is set
                    if (flag >= 0) {
                        for (int r = 0; r < NX; ++r) xB[r] = bwd.zn[0][r];
                        for (int r = 0; r < c.nplist; ++r) xQ[r] = bwd.znQ[0][r];
                    }
                }
                called = true;
                if (flag < 0) { status = flag; break; }
                t = tnext;
            }
            if (last) break;
            for (int r = 0; r < NX; ++r) xB[r] += dj(it, r);   // handle_datapoint_b
            --it;
            // reinit_b + reinit_quad_b
            for (int r = 0; r < NX; ++r) bwd.zn[0][r] = xB[r];
            for (int r = 0; r < c.nplist; ++r) bwd.znQ[0][r] = xQ[r];
            bwd.pending = false;
            if (called) bwd.amici_reinit(t);
            else bwd.init(t);
            if (it < 0 && !(t > o.t0)) break;
        }
        if (status == AB_SUCCESS) {
            // ReturnData::handle_sx0_forward + process_backward_problem: sllh -= sx0^T xB + xQB
            for (int il = 0; il < c.nplist; ++il) {
                double sx0[NX];
                M::sx0(c.plist[il], c.p, c.k, sx0);
                if (b.sx0_in)
                    for (int r = 0; r < NX; ++r) sx0[r] = b.sx0_in[((size_t)n * c.nplist + il) * NX + r] / pcoef[il];
                double s = 0.0;
                for (int r = 0; r < NX; ++r) s += xB[r] * sx0[r];
                sllh[il] -= s + xQ[il];
            }
        }
        finish(status);
    }
};

}  // namespace amici_b200
#endif
