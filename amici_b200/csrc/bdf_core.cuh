// Batched variable-order, variable-step BDF with simultaneous forward sensitivities.
//
// Behavioural specification: SUNDIALS CVODES 7.1.1 as driven by AMICI's CVodeSolver
// (reference ThirdParty/sundials/src/cvodes/cvodes.c; cvodes_nls_sim.c; cvodes_ls.c;
// sunnonlinsol_newton.c; reference src/solver_cvodes.cpp, src/forwardproblem.cpp).  This is not
// a translation of those sources: there are no N_Vector/SUNMatrix/SUNLinearSolver objects.
// The stacked (state | sensitivity) system is stored column-wise, every lane of a lane-group
// (group.cuh) owns C columns, the Nordsieck history of all columns sits in the group's
// shared-memory slice, the per-column work vectors in registers, and all scalar step control
// is computed redundantly per lane.  Function names below cite the CVODES routine whose
// decisions they must reproduce (constants: reference cvodes_impl.h:57-131,202-204;
// cvodes_ls_impl.h:40-41).
//
// Code layout is driven by the instruction cache (profiles/r01_v0_ncu_summary.md): everything on
// the per-step path is inlined exactly once (one call site for the model right-hand side, the
// Pascal-triangle update, the rescale, the triangular solves), everything rare (first step,
// initial step size, order-1 restart, stability-limit detection) is an out-of-line AB_COLD
// function that only touches the shared-memory slice, so the solver object itself never has its
// address taken and stays in registers.
//
// The same source is compiled for the host with G = 1 by tests/cpu_emul (algorithm check
// against the compiled reference without a GPU); the product is the CUDA instantiation.
#ifndef AMICI_B200_BDF_CORE_CUH
#define AMICI_B200_BDF_CORE_CUH
#include "../../include/amici_b200.h"
#include "group.cuh"
#include "model_common.cuh"
#include "spline.cuh"

namespace amici_b200 {

// plain-data interface structs: the C ABI's (include/amici_b200.h)
using SolverOptions = ab_solver_options;
using BatchView = ab_batch;

// ---------------------------------------------------------------------------------------------
// per-trajectory shared storage
// ---------------------------------------------------------------------------------------------
constexpr int QMAX = 5;           // BDF_Q_MAX
#ifndef AB_SLDET_SVC
#define AB_SLDET_SVC 0            // 1: lane groups of 16 hand cvSLdet to a service warp of their block (launch.cuh sldet_service)
#endif
#ifndef AB_SLDET_SVC_SLEEP
#define AB_SLDET_SVC_SLEEP 32
#endif
enum : int { SVC_REQ = 0x10000, SVC_ANS = 0x20000 };
constexpr int LMAX = QMAX + 1;

template <class M, int NCOLS>
struct TrajShared {
    double zn[LMAX][M::NX][NCOLS];   // Nordsieck history, column fastest
    double acorbuf[M::NX][NCOLS];    // accepted correction of the last step while its cvCompleteStep is deferred (Bdf::pending)
    double lus[M::NLU];              // sparse LU of I - gamma*J (static pivot order, emitted per model)
    double lu[M::NX][M::NX];         // dense LU with partial pivoting: fallback when a static pivot is rejected
    double savedJ[M::NNZ_J > 0 ? M::NNZ_J : 1];
    double ybuf[M::NX];              // broadcast buffer for the state column
    double xfin[M::NX];              // state at the last successful solver return (ReturnData::xdot/J are evaluated there)
    double scratch[M::NP + 1];       // exchange buffer over OUTPUT columns (FIM rows)
    double p[M::NP > 0 ? M::NP : 1];
    double k[M::NK_EXT > 0 ? M::NK_EXT : 1];   // fixed parameters, then the model's derived constants (M::derived)
    double h[M::NH > 0 ? M::NH : 1];
    double pcoef[NCOLS];             // chain-rule factor per integrated column
    double pcoef_out[M::NP + 1];     // ... per output column (0 = state, 1 + il = plist entry il)
    double ctl[16];                  // mirrors of tau[0..6] (ctl[0..6]) and l[0..5] (ctl[8..13]) for indexing by the run-time order
    double ssdat[5][8];              // cvBDFStab history: ring buffer of (sqm2^2, sqm1^2, sq^2) in [1..3] and their ratios to
                                     // the next older entry in [5..7] (the rat[][] of cvSLdet, computed once), newest at ss_head
    double sld[40];                  // exchange area of bdf_sldet_xchg (ten rounds of up to four values)
    unsigned long long xpos_mask;    // Model::set_state_is_non_negative as a bit set (AB_FULL builds; 0 = off)
    long long t_start_ns;            // device clock when the condition started (Solver::set_max_time; AB_FULL builds)
#if AB_SLDET_SVC
    int svc;                         // mailbox of the block's cvSLdet service warp (AB_SLDET_SVC, launch.cuh): request / result word
#endif
    int piv[M::NX];
    int ip_of_col[NCOLS];            // parameter index of an integrated column (-1: state / unused)
    int il_of_col[NCOLS];            // its position in plist
    int col_of_out[M::NP + 1];       // integrated column of an output column, -1 if the parameter's sensitivity is identically 0
    SplineTable<M, NCOLS> spl;
};

// COMPACT: the lane group is too small for one column per parameter, but large enough for the parameters whose state
// sensitivity is not identically zero (M::SX_STATIC_MASK, written by the emitter): only those columns are integrated.  The
// dropped columns are exact zeros in CVODES too (zero initial value, zero right-hand side) and a zero column never wins the
// max-norms that drive the step control, so every decision and every integrated value is unchanged; the outputs of the
// dropped parameters are formed from sx = 0 at the output points (Simulation::datapoint).
template <class M, int G, int C>
struct Layout {
    static constexpr int NX = M::NX;
    static constexpr bool COMPACT = M::SX_STATIC_MASK != 0ull && G * C < M::NP + 1 && G * C >= M::NP_DYN + 1;
    static constexpr int NCOLS_MAX = COMPACT ? M::NP_DYN + 1 : M::NP + 1;
    static constexpr int NCOLS = (G * C < NCOLS_MAX) ? G * C : NCOLS_MAX;  // columns that can ever be active
    using Shared = TrajShared<M, NCOLS>;
};

// internal flags of cvStep (cvodes.c local defines)
enum : int { FIRST_CALL = 101, PREV_CONV_FAIL = 102, PREV_ERR_FAIL = 103 };
enum : int { NLS_CONV_RECVR = 902, NLS_RHS_RECVR = 903 };
enum : int { CONVFAIL_NONE = 0, CONVFAIL_BAD_J = 1, CONVFAIL_OTHER = 2 };

template <int N>
AB_HD double pick(const double (&a)[N], int i) {
    double v = a[0];
    AB_UNROLL
    for (int j = 1; j < N; ++j) v = (i == j) ? a[j] : v;
    return v;
}

// ---------------------------------------------------------------------------------------------
// model right-hand side of the stacked system (shared by the hot inlined path and the cold helpers)
// ---------------------------------------------------------------------------------------------
template <class M, int C>
struct ModelEval {
    double y[M::NX];                                   // state column (broadcast)
    double spl[M::NSPL > 0 ? M::NSPL : 1];
    double sspl[C][M::NSPL > 0 ? M::NSPL : 1];
};

template <class M, int G, int C>
AB_HD void eval_splines(const typename Layout<M, G, C>::Shared& S, const Group<G>& g, int ncols, double t,
                        ModelEval<M, C>& e) {
    constexpr int NCOLS = Layout<M, G, C>::NCOLS;
    AB_UNROLL
    for (int is = 0; is < M::NSPL; ++is) {
        const SplineLoc loc = spline_locate<M>(is, t);
        const double v = spline_eval_at<M, NCOLS>(S.spl, is, 0, loc, true);
        e.spl[is] = spline_is_log<M>(is) ? exp(v) : v;
        AB_UNROLL
        for (int c = 0; c < C; ++c) {
            const int cc = g.lane + G * c;
            e.sspl[c][is] = 0.0;
            if (cc >= ncols || cc == 0) continue;
            const double s = spline_eval_at<M, NCOLS>(S.spl, is, cc, loc, false);
            e.sspl[c][is] = spline_is_log<M>(is) ? s * e.spl[is] : s;
        }
    }
}

// Model::compute_x_pos (reference src/model.cpp:3088-3108): the model functions see max(x_i, 0) for the states flagged non-negative
template <class M, class SharedT>
AB_HD void x_pos(const SharedT& S, double (&x)[M::NX]) {
#if AB_FULL
    const unsigned long long m = S.xpos_mask;
    if (m) {
        AB_UNROLL
        for (int r = 0; r < M::NX; ++r)
            if ((m >> r & 1ull) && x[r] < 0.0) x[r] = 0.0;
    }
#else
    (void)S; (void)x;
#endif
}

// Broadcast the state column (valid on the lane that owns column 0) to all lanes of the group.
template <class M, int G, int C>
AB_HD void bcast_state(typename Layout<M, G, C>::Shared& S, const Group<G>& g, const double (&v)[C][M::NX],
                       ModelEval<M, C>& e) {
    if (G == 1) {
        AB_UNROLL
        for (int r = 0; r < M::NX; ++r) e.y[r] = v[0][r];
        x_pos<M>(S, e.y);
        return;
    }
    g.sync();
    if (g.lane == 0) {
        AB_UNROLL
        for (int r = 0; r < M::NX; ++r) S.ybuf[r] = v[0][r];
    }
    g.sync();
    AB_UNROLL
    for (int r = 0; r < M::NX; ++r) e.y[r] = S.ybuf[r];
    x_pos<M>(S, e.y);
}

// f for column 0, J(y)*s_c + dfdp_c for sensitivity columns (Model_ODE::fxdot / fsxdot,
// reference src/model_ode.cpp:105-114, :426-450).  ycols = current values of all own columns.
// Returns false if any output is non-finite (AMICI's check_finite -> recoverable error).
// same_t: e.spl / e.sspl already hold the splines at this t (later Newton iterations of one step evaluate at the same tn).
template <class M, int G, int C>
AB_HD bool model_rhs(typename Layout<M, G, C>::Shared& S, const Group<G>& g, int ncols, double t,
                     const double (&ycols)[C][M::NX], double (&out)[C][M::NX], ModelEval<M, C>& e, const bool same_t = false) {
    bcast_state<M, G, C>(S, g, ycols, e);
    if (M::NSPL > 0 && !same_t) eval_splines<M, G, C>(S, g, ncols, t, e);
    double Jv[M::NNZ_J > 0 ? M::NNZ_J : 1];
    AB_UNROLL
    for (int i = 0; i < (M::NNZ_J > 0 ? M::NNZ_J : 1); ++i) Jv[i] = 0.0;
    bool need_j = false;
    AB_UNROLL
    for (int c = 0; c < C; ++c) need_j = need_j || (g.lane + G * c < ncols && g.lane + G * c > 0);
    if (need_j) M::jac(t, e.y, S.p, S.k, S.h, e.spl, Jv);
    bool ok = true;
    AB_UNROLL
    for (int c = 0; c < C; ++c) {
        const int cc = g.lane + G * c;
        if (cc >= ncols) continue;
        if (cc == 0) {
            M::xdot(t, e.y, S.p, S.k, S.h, e.spl, out[c]);
        } else {
            M::dxdotdp(S.ip_of_col[cc], t, e.y, S.p, S.k, S.h, e.spl, e.sspl[c], out[c]);
            M::jac_mul_add(Jv, ycols[c], out[c]);
        }
        AB_UNROLL
        for (int r = 0; r < M::NX; ++r) ok = ok && isfinite(out[c][r]);
    }
    return !g.any(!ok);
}

// ---------------------------------------------------------------------------------------------
// cold helpers (out of line; touch only the shared-memory slice)
// ---------------------------------------------------------------------------------------------
// zn[dst] <- scale * F(t, zn[0] + a * zn[1]); returns false on a non-finite right-hand side
template <class M, int G, int C>
AB_COLD bool cold_rhs_store(typename Layout<M, G, C>::Shared& S, int ncols, double t, double a, double scale, int dst) {
    constexpr int NX = M::NX;
    Group<G> g;
    ModelEval<M, C> e;
    double yy[C][NX], f[C][NX];
    AB_UNROLL
    for (int c = 0; c < C; ++c) {
        const int cc = g.lane + G * c;
        AB_UNROLL
        for (int r = 0; r < NX; ++r) yy[c][r] = (cc < ncols) ? S.zn[0][r][cc] + (a != 0.0 ? a * S.zn[1][r][cc] : 0.0) : 0.0;
    }
    const bool ok = model_rhs<M, G, C>(S, g, ncols, t, yy, f, e);
    g.sync();
    AB_UNROLL
    for (int c = 0; c < C; ++c) {
        const int cc = g.lane + G * c;
        if (cc >= ncols) continue;
        AB_UNROLL
        for (int r = 0; r < NX; ++r) S.zn[dst][r][cc] = scale * f[c][r];
    }
    g.sync();
    return ok;
}

// cvHin / cvUpperBoundH0 / cvYddNorm (cvodes.c:5436-5798): initial step size. zn[0], zn[1] = y0, f(t0, y0);
// zn[2] is used as scratch (the order is 1 here). Returns AB_SUCCESS / AB_TOO_CLOSE / AB_REPTD_RHSFUNC_ERR.
template <class M, int G, int C>
AB_COLD int cold_hin(typename Layout<M, G, C>::Shared& S, const SolverOptions& o, int ncols, double tn, double tout,
                     double* h_out, int* nfe) {
    constexpr int NX = M::NX;
    Group<G> g;
    const double tdiff = tout - tn;
    if (tdiff == 0.0) return AB_TOO_CLOSE;
    const double sign = tdiff > 0.0 ? 1.0 : -1.0;
    const double tdist = fabs(tdiff);
    const double tround = DBL_EPSILON * ab_max(fabs(tn), fabs(tout));
    if (tdist < 2.0 * tround) return AB_TOO_CLOSE;
    const double hlb = 100.0 * tround;
    double hub_inv = 0.0;
    AB_UNROLL
    for (int c = 0; c < C; ++c) {
        const int cc = g.lane + G * c;
        if (cc >= ncols) continue;
        const double rt = cc == 0 ? o.rtol : o.rtol_fsa, at = cc == 0 ? o.atol : o.atol_fsa;
        AB_UNROLL
        for (int r = 0; r < NX; ++r) {
            const double ew = 1.0 / (rt * fabs(S.zn[0][r][cc]) + at);
            const double den = 0.1 * fabs(S.zn[0][r][cc]) + 1.0 / ew;
            const double v = fabs(S.zn[1][r][cc]) / den;
            hub_inv = (v > hub_inv || v != v) ? v : hub_inv;
        }
    }
    hub_inv = g.max(hub_inv);
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
            const bool ok = cold_rhs_store<M, G, C>(S, ncols, tn + hgs, hgs, 1.0, 2);
            (*nfe)++;
            if (ok) {
                double m = 0.0;
                AB_UNROLL
                for (int c = 0; c < C; ++c) {
                    const int cc = g.lane + G * c;
                    if (cc >= ncols) continue;
                    const double rt = cc == 0 ? o.rtol : o.rtol_fsa, at = cc == 0 ? o.atol : o.atol_fsa;
                    double s = 0.0;
                    AB_UNROLL
                    for (int r = 0; r < NX; ++r) {
                        const double ew = 1.0 / (rt * fabs(S.zn[0][r][cc]) + at);
                        const double ydd = (1.0 / hgs) * S.zn[2][r][cc] - (1.0 / hgs) * S.zn[1][r][cc];
                        const double pr = ydd * ew;
                        s += pr * pr;
                    }
                    const double n = sqrt(s * (1.0 / NX));
                    m = (n > m || n != n) ? n : m;
                }
                yddnrm = g.max(m);
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

// cvSLdet (cvodes.c:8256): stability-limit detection on the 5-deep history of scaled derivative norms.
// s[i][k]: i = 1 (current) .. 5 (oldest), k = 1..3 (orders q-1, q, q+1).  All loops are unrolled with static
// indices so that the work arrays live in registers; divisions whose divisor has been range-checked use fast_div.
#if defined(AB_SLDET_STATS) && !defined(__CUDA_ARCH__)
#define AB_SLDET_TRACE(tag, rr) sldet_trace(tag, rr)
void sldet_trace(int tag, double rr);
#else
#define AB_SLDET_TRACE(tag, rr)
#endif
#if defined(AB_SLDET_RECORD) && !defined(__CUDA_ARCH__)
void sldet_record(const double* ssdat, int head, int q, int rc);
#endif
AB_HD int bdf_sldet(const double (&s)[6][4], const double (&ratc)[5][4], const int q) {
    const double TINY_ = 1e-10;
    double rav[4], qkr[4], sigsq[4], smax[4], ssmax[4], vrat[4], qc[6][4], qco[6][4];
    const double rrcut = 0.98, vrrtol = 1.0e-4, vrrt2 = 5.0e-4, sqtol = 1.0e-3, rrtol = 1.0e-2;
    double rr = 0.0;
    int kflag = 0;
    bool bad = false;
    AB_UNROLL
    for (int kk = 1; kk <= 3; kk++) {
        double smink = s[1][kk], smaxk = 0.0;
        AB_UNROLL
        for (int i = 1; i <= 5; i++) {
            smink = ab_min(smink, s[i][kk]);
            smaxk = ab_max(smaxk, s[i][kk]);
        }
        bad = bad || (smink < TINY_ * smaxk);
        smax[kk] = smaxk;
        ssmax[kk] = smaxk * smaxk;
    }
    if (bad) return -1;
    AB_UNROLL
    for (int kk = 1; kk <= 3; kk++) {
        double sumrat = 0.0, sumrsq = 0.0;
        AB_UNROLL
        for (int i = 1; i <= 4; i++) {
            const double rat = ratc[i][kk];   // s[i][kk] / s[i + 1][kk], cached by bdf_stab
            sumrat += rat;
            sumrsq += rat * rat;
        }
        rav[kk] = 0.25 * sumrat;
        vrat[kk] = fabs(0.25 * sumrsq - rav[kk] * rav[kk]);
        qc[5][kk] = s[1][kk] * s[3][kk] - s[2][kk] * s[2][kk];
        qc[4][kk] = s[2][kk] * s[3][kk] - s[1][kk] * s[4][kk];
        qc[3][kk] = 0.0;
        qc[2][kk] = s[2][kk] * s[5][kk] - s[3][kk] * s[4][kk];
        qc[1][kk] = s[4][kk] * s[4][kk] - s[3][kk] * s[5][kk];
        AB_UNROLL
        for (int i = 1; i <= 5; i++) qco[i][kk] = qc[i][kk];
    }
    const double vmin = ab_min(vrat[1], ab_min(vrat[2], vrat[3]));
    const double vmax = ab_max(vrat[1], ab_max(vrat[2], vrat[3]));
    if (vmin < vrrtol * vrrtol) {
        if (vmax > vrrt2 * vrrt2) return -2;
        rr = (rav[1] + rav[2] + rav[3]) / 3.0;
        double drrmax = 0.0;
        AB_UNROLL
        for (int kk = 1; kk <= 3; kk++) drrmax = ab_max(drrmax, fabs(rav[kk] - rr));
        if (drrmax > vrrt2) return -3;
        kflag = 1;
        AB_SLDET_TRACE(1, rr);
    } else {
        if (fabs(qco[1][1]) < TINY_ * ssmax[1]) return -4;
        double tem = fast_div(qco[1][2], qco[1][1]);
        AB_UNROLL
        for (int i = 2; i <= 5; i++) qco[i][2] -= tem * qco[i][1];
        tem = fast_div(qco[1][3], qco[1][1]);
        AB_UNROLL
        for (int i = 2; i <= 5; i++) qco[i][3] -= tem * qco[i][1];
        if (fabs(qco[2][2]) < TINY_ * ssmax[2]) return -4;
        tem = fast_div(qco[2][3], qco[2][2]);
        AB_UNROLL
        for (int i = 3; i <= 5; i++) qco[i][3] -= tem * qco[i][2];
        if (fabs(qco[4][3]) < TINY_ * ssmax[3]) return -4;
        rr = -fast_div(qco[5][3], qco[4][3]);
        if (rr < TINY_ || rr > 100.0) return -5;
        AB_UNROLL
        for (int kk = 1; kk <= 3; kk++)
            qkr[kk] = qc[5][kk] + rr * (qc[4][kk] + rr * rr * (qc[2][kk] + rr * qc[1][kk]));
        // max_k |qkr_k| / ssmax_k < sqtol without the three quotients (only the comparison is needed)
        bool sq_small = true;
        AB_UNROLL
        for (int kk = 1; kk <= 3; kk++) sq_small = sq_small && (fabs(qkr[kk]) < sqtol * ssmax[kk]);
        AB_SLDET_TRACE(20, rr);
        AB_SLDET_TRACE(sq_small ? 21 : 22, rr);
        if (sq_small) {
            kflag = 2;
        } else {
            // Newton corrections to improve rr.  This is the common branch (JAK-STAT: 59 % of the calls that get here), so
            // the nine quotients |qjk| / ssmax_j per sweep are products with three reciprocals formed once.
            double sqmin = 0.0, inv_ssmax[4];
            AB_UNROLL
            for (int j = 1; j <= 3; j++) inv_ssmax[j] = fast_div(1.0, ssmax[j]);
            AB_NOUNROLL
            for (int it = 1; it <= 3; it++) {
                AB_SLDET_TRACE(23, rr);
                double rrc[4], sqmx[4], qjk[4][4];
                AB_UNROLL
                for (int kk = 1; kk <= 3; kk++) {
                    const double qp = qc[4][kk] + rr * rr * (3.0 * qc[2][kk] + rr * 4.0 * qc[1][kk]);
                    double drr = 0.0;
                    if (fabs(qp) > TINY_ * ssmax[kk]) drr = -fast_div(qkr[kk], qp);
                    rrc[kk] = rr + drr;
                }
                AB_UNROLL
                for (int kk = 1; kk <= 3; kk++) {
                    const double sv = rrc[kk];
                    double sqmaxk = 0.0;
                    AB_UNROLL
                    for (int j = 1; j <= 3; j++) {
                        qjk[j][kk] = qc[5][j] + sv * (qc[4][j] + sv * sv * (qc[2][j] + sv * qc[1][j]));
                        const double saqj = fabs(qjk[j][kk]) * inv_ssmax[j];
                        if (saqj > sqmaxk) sqmaxk = saqj;
                    }
                    sqmx[kk] = sqmaxk;
                }
                sqmin = sqmx[1] + 1.0;
                int kmin = 1;
                AB_UNROLL
                for (int kk = 1; kk <= 3; kk++)
                    if (sqmx[kk] < sqmin) {
                        kmin = kk;
                        sqmin = sqmx[kk];
                    }
                rr = (kmin == 1) ? rrc[1] : (kmin == 2) ? rrc[2] : rrc[3];
                if (sqmin < sqtol) {
                    kflag = 3;
                    break;
                }
                AB_UNROLL
                for (int j = 1; j <= 3; j++) qkr[j] = (kmin == 1) ? qjk[j][1] : (kmin == 2) ? qjk[j][2] : qjk[j][3];
            }
            if (sqmin > sqtol) return -6;
        }
    }
    // Only a result > 3 makes cvBDFStab act, and that needs rr > rrcut (the promotion of kflag at the end): with
    // rr <= rrcut every remaining path returns a value <= 3, so the rest can be skipped.
    if (!(rr > rrcut)) return kflag;
    AB_SLDET_TRACE(30, rr);
    bool bad7 = false;
    AB_UNROLL
    for (int kk = 1; kk <= 3; kk++) {
        const double rsa = s[1][kk], rsb = s[2][kk] * rr, rsc = s[3][kk] * rr * rr;
        const double rsd = s[4][kk] * rr * rr * rr;
        const double rd1a = rsa - rsb, rd1b = rsb - rsc, rd1c = rsc - rsd;
        const double rd2a = rd1a - rd1b, rd2b = rd1b - rd1c, rd3a = rd2a - rd2b;
        // CVODES returns -7 at the first k that fails; later k are not evaluated there, and their values are unused here
        const bool small = fabs(rd1b) < TINY_ * smax[kk];
        const double cest1 = small ? 1.0 : -fast_div(rd3a, rd1b);   // |rd1b| >= TINY * smax here
        bad7 = bad7 || small || (cest1 < TINY_ || cest1 > 4.0);
        const double corr1 = fast_div(rd2b, cest1 * (rr * rr));   // (rd2b / cest1) / rr^2 with one quotient; the value is only used if cest1 is in [TINY, 4]
        sigsq[kk] = s[3][kk] + corr1;
    }
    if (bad7) return -7;
    if (sigsq[2] < TINY_) return -8;
    const double ratp = fast_div(sigsq[3], sigsq[2]), ratm = fast_div(sigsq[1], sigsq[2]);
    const double qfac1 = 0.25 * (q * q - 1.0), qfac2 = (q == 3) ? 1.0 : (q == 4) ? (2.0 / 3.0) : 0.5;   // 2 / (q - 1), q in 3..5
    const double bb = ratp * ratm - 1.0 - qfac1 * ratp;
    const double tem = 1.0 - qfac2 * bb;
    if (fabs(tem) < TINY_) return -8;
    const double rrb = fast_div(1.0, tem);   // |tem| >= TINY
    if (fabs(rrb - rr) > rrtol) return -9;
    if (rr > rrcut) {
        if (kflag == 1) kflag = 4;
        if (kflag == 2) kflag = 5;
        if (kflag == 3) kflag = 6;
    }
    return kflag;
}

// cvSLdet with the three orders k = 1..3 in three LANES of the group instead of three unrolled copies in every lane.
// The stability test is a long chain of dependent scalar FP64 operations that every lane would execute identically; it
// sits on the pace-setting path of the aligned warps (a quarter of the JAK-STAT step time with the serial form).  Lane l
// owns order k = l % 4 + 1 (lanes with l % 4 == 3 duplicate k = 1), lanes 0..2 are the shuffle sources; the arithmetic per
// order is the serial function's, the cross-order steps (elimination, candidate scan, final ratios) exchange the few values
// they need by shuffles.  Same return values as bdf_sldet.
template <int G>
AB_HD int bdf_sldet_lanes(const Group<G>& g, const double (&ssdat)[5][8], const int ss_head, const int q) {
    const double TINY_ = 1e-10;
    const double rrcut = 0.98, vrrtol = 1.0e-4, vrrt2 = 5.0e-4, sqtol = 1.0e-3, rrtol = 1.0e-2;
    const int k = ((g.lane & 3) == 3) ? 1 : (g.lane & 3) + 1;
    double s[6], rat[5];
    AB_UNROLL
    for (int i = 1; i <= 5; ++i) {
        int idx = ss_head + i - 1;
        idx = idx >= 5 ? idx - 5 : idx;
        s[i] = ssdat[idx][k];
        if (i <= 4) rat[i] = ssdat[idx][4 + k];
    }
    double smink = s[1], smaxk = 0.0;
    AB_UNROLL
    for (int i = 1; i <= 5; i++) {
        smink = ab_min(smink, s[i]);
        smaxk = ab_max(smaxk, s[i]);
    }
    if (g.any(smink < TINY_ * smaxk)) return -1;
    const double smax = smaxk, ssmax = smaxk * smaxk;
    double sumrat = 0.0, sumrsq = 0.0;
    AB_UNROLL
    for (int i = 1; i <= 4; i++) {
        sumrat += rat[i];
        sumrsq += rat[i] * rat[i];
    }
    const double rav = 0.25 * sumrat;
    const double vrat = fabs(0.25 * sumrsq - rav * rav);
    const double qc5 = s[1] * s[3] - s[2] * s[2];
    const double qc4 = s[2] * s[3] - s[1] * s[4];
    const double qc2 = s[2] * s[5] - s[3] * s[4];
    const double qc1 = s[4] * s[4] - s[3] * s[5];
    const double vmin = g.min(vrat), vmax = g.max(vrat);
    double rr = 0.0;
    int kflag = 0;
    if (vmin < vrrtol * vrrtol) {
        if (vmax > vrrt2 * vrrt2) return -2;
        rr = (g.bcast(rav, 0) + g.bcast(rav, 1) + g.bcast(rav, 2)) / 3.0;
        if (g.max(fabs(rav - rr)) > vrrt2) return -3;
        kflag = 1;
    } else {
        // eliminate with column 1, then column 2; rr from column 3 (rows 1, 2, 4, 5; row 3 is zero)
        const double a1 = g.bcast(qc1, 0), a2 = g.bcast(qc2, 0), a4 = g.bcast(qc4, 0), a5 = g.bcast(qc5, 0);
        if (fabs(a1) < TINY_ * g.bcast(ssmax, 0)) return -4;
        const double tem1 = fast_div(qc1, a1);
        const double e2 = qc2 - tem1 * a2, e4 = qc4 - tem1 * a4, e5 = qc5 - tem1 * a5;
        const double b2 = g.bcast(e2, 1), b4 = g.bcast(e4, 1), b5 = g.bcast(e5, 1);
        if (fabs(b2) < TINY_ * g.bcast(ssmax, 1)) return -4;
        const double tem2 = fast_div(e2, b2);
        const double f4 = e4 - tem2 * b4, f5 = e5 - tem2 * b5;
        const double c4 = g.bcast(f4, 2), c5 = g.bcast(f5, 2);
        if (fabs(c4) < TINY_ * g.bcast(ssmax, 2)) return -4;
        rr = -fast_div(c5, c4);
        if (rr < TINY_ || rr > 100.0) return -5;
        double qkr = qc5 + rr * (qc4 + rr * rr * (qc2 + rr * qc1));
        if (!g.any(!(fabs(qkr) < sqtol * ssmax))) {
            kflag = 2;
        } else {
            const double inv_ssmax = fast_div(1.0, ssmax);
            double sqmin = 0.0;
            AB_NOUNROLL
            for (int it = 1; it <= 3; it++) {
                const double qp = qc4 + rr * rr * (3.0 * qc2 + rr * 4.0 * qc1);
                double drr = 0.0;
                if (fabs(qp) > TINY_ * ssmax) drr = -fast_div(qkr, qp);
                const double rrc = rr + drr;
                const double sv1 = g.bcast(rrc, 0), sv2 = g.bcast(rrc, 1), sv3 = g.bcast(rrc, 2);
                const double p1 = qc5 + sv1 * (qc4 + sv1 * sv1 * (qc2 + sv1 * qc1));
                const double p2 = qc5 + sv2 * (qc4 + sv2 * sv2 * (qc2 + sv2 * qc1));
                const double p3 = qc5 + sv3 * (qc4 + sv3 * sv3 * (qc2 + sv3 * qc1));
                const double m1 = g.max(fabs(p1) * inv_ssmax), m2 = g.max(fabs(p2) * inv_ssmax), m3 = g.max(fabs(p3) * inv_ssmax);
                sqmin = m1 + 1.0;
                int kmin = 1;
                if (m1 < sqmin) { kmin = 1; sqmin = m1; }
                if (m2 < sqmin) { kmin = 2; sqmin = m2; }
                if (m3 < sqmin) { kmin = 3; sqmin = m3; }
                rr = (kmin == 1) ? sv1 : (kmin == 2) ? sv2 : sv3;
                if (sqmin < sqtol) {
                    kflag = 3;
                    break;
                }
                qkr = (kmin == 1) ? p1 : (kmin == 2) ? p2 : p3;
            }
            if (sqmin > sqtol) return -6;
        }
    }
    if (!(rr > rrcut)) return kflag;   // only a result > 3 acts, and that needs rr > rrcut
    const double rsa = s[1], rsb = s[2] * rr, rsc = s[3] * rr * rr;
    const double rsd = s[4] * rr * rr * rr;
    const double rd1a = rsa - rsb, rd1b = rsb - rsc, rd1c = rsc - rsd;
    const double rd2a = rd1a - rd1b, rd2b = rd1b - rd1c, rd3a = rd2a - rd2b;
    const bool small = fabs(rd1b) < TINY_ * smax;
    const double cest1 = small ? 1.0 : -fast_div(rd3a, rd1b);
    if (g.any(small || cest1 < TINY_ || cest1 > 4.0)) return -7;
    const double sigsq = s[3] + fast_div(rd2b, cest1 * (rr * rr));
    const double sg1 = g.bcast(sigsq, 0), sg2 = g.bcast(sigsq, 1), sg3 = g.bcast(sigsq, 2);
    if (sg2 < TINY_) return -8;
    const double ratp = fast_div(sg3, sg2), ratm = fast_div(sg1, sg2);
    const double qfac1 = 0.25 * (q * q - 1.0), qfac2 = (q == 3) ? 1.0 : (q == 4) ? (2.0 / 3.0) : 0.5;
    const double bb = ratp * ratm - 1.0 - qfac1 * ratp;
    const double tem = 1.0 - qfac2 * bb;
    if (fabs(tem) < TINY_) return -8;
    const double rrb = fast_div(1.0, tem);
    if (fabs(rrb - rr) > rrtol) return -9;
    if (kflag == 1) kflag = 4;
    if (kflag == 2) kflag = 5;
    if (kflag == 3) kflag = 6;
    return kflag;
}

// cvSLdet with the three orders in three lanes and the cross-order values exchanged through the group's shared-memory
// slice (one store + one group barrier + broadcast loads per exchange, each exchange in its own slots).  For lane groups that
// share their warp with other conditions (G < 32): register shuffles / votes with a partial lane mask compile into a
// divergence check plus an out-of-line serialised region per call site, and the all-orders-per-lane form (bdf_sldet) is
// 5.5 KB of straight-line code in a step loop whose speed is set by instruction supply (profiles/r02_*).  Same arithmetic per
// order and same return values as bdf_sldet / bdf_sldet_lanes.
template <int G>
AB_HD int bdf_sldet_xchg(const Group<G>& g, const double (&ssdat)[5][8], double (&xb)[40], const int ss_head, const int q) {
    const double TINY_ = 1e-10;
    const double rrcut = 0.98, vrrtol = 1.0e-4, vrrt2 = 5.0e-4, sqtol = 1.0e-3, rrtol = 1.0e-2;
    const int k = ((g.lane & 3) == 3) ? 1 : (g.lane & 3) + 1;
    const bool pub = g.lane < 3;   // lanes 0..2 publish orders 1..3
    // exchange round `slot`: every lane gets the three lanes' values
#define AB_X3(slot, v, v1, v2, v3)        \
    do {                                  \
        if (pub) xb[4 * (slot) + g.lane] = (v); \
        g.sync();                         \
        v1 = xb[4 * (slot)];              \
        v2 = xb[4 * (slot) + 1];          \
        v3 = xb[4 * (slot) + 2];          \
    } while (0)
    double s[6], rat[5];
    g.sync();   // the previous call's loads are done before this call's stores
    AB_UNROLL
    for (int i = 1; i <= 5; ++i) {
        int idx = ss_head + i - 1;
        idx = idx >= 5 ? idx - 5 : idx;
        s[i] = ssdat[idx][k];
        if (i <= 4) rat[i] = ssdat[idx][4 + k];
    }
    double smink = s[1], smaxk = 0.0;
    AB_UNROLL
    for (int i = 1; i <= 5; i++) {
        smink = ab_min(smink, s[i]);
        smaxk = ab_max(smaxk, s[i]);
    }
    const double smax = smaxk, ssmax = smaxk * smaxk;
    double sumrat = 0.0, sumrsq = 0.0;
    AB_UNROLL
    for (int i = 1; i <= 4; i++) {
        sumrat += rat[i];
        sumrsq += rat[i] * rat[i];
    }
    const double rav = 0.25 * sumrat;
    const double vrat = fabs(0.25 * sumrsq - rav * rav);
    const double qc5 = s[1] * s[3] - s[2] * s[2];
    const double qc4 = s[2] * s[3] - s[1] * s[4];
    const double qc2 = s[2] * s[5] - s[3] * s[4];
    const double qc1 = s[4] * s[4] - s[3] * s[5];
    double b1, b2, b3, v1, v2, v3;
    AB_X3(0, (smink < TINY_ * smaxk) ? 1.0 : 0.0, b1, b2, b3);
    if (b1 + b2 + b3 != 0.0) return -1;
    AB_X3(1, vrat, v1, v2, v3);
    const double vmin = ab_min(v1, ab_min(v2, v3)), vmax = ab_max(v1, ab_max(v2, v3));
    double rr = 0.0;
    int kflag = 0;
    if (vmin < vrrtol * vrrtol) {
        if (vmax > vrrt2 * vrrt2) return -2;
        double r1, r2, r3;
        AB_X3(2, rav, r1, r2, r3);
        rr = (r1 + r2 + r3) / 3.0;
        const double drrmax = ab_max(fabs(r1 - rr), ab_max(fabs(r2 - rr), fabs(r3 - rr)));
        if (drrmax > vrrt2) return -3;
        kflag = 1;
    } else {
        // eliminate with column 1, then column 2; rr from column 3 (rows 1, 2, 4, 5; row 3 is zero)
        if (g.lane == 0) { xb[8] = qc1; xb[9] = qc2; xb[10] = qc4; xb[11] = qc5; xb[12] = ssmax; }
        g.sync();
        const double a1 = xb[8], a2 = xb[9], a4 = xb[10], a5 = xb[11];
        if (fabs(a1) < TINY_ * xb[12]) return -4;
        const double tem1 = fast_div(qc1, a1);
        const double e2 = qc2 - tem1 * a2, e4 = qc4 - tem1 * a4, e5 = qc5 - tem1 * a5;
        if (g.lane == 1) { xb[13] = e2; xb[14] = e4; xb[15] = e5; xb[16] = ssmax; }
        g.sync();
        const double bb2 = xb[13], bb4 = xb[14], bb5 = xb[15];
        if (fabs(bb2) < TINY_ * xb[16]) return -4;
        const double tem2 = fast_div(e2, bb2);
        const double f4 = e4 - tem2 * bb4, f5 = e5 - tem2 * bb5;
        if (g.lane == 2) { xb[17] = f4; xb[18] = f5; xb[19] = ssmax; }
        g.sync();
        const double c4 = xb[17], c5 = xb[18];
        if (fabs(c4) < TINY_ * xb[19]) return -4;
        rr = -fast_div(c5, c4);
        if (rr < TINY_ || rr > 100.0) return -5;
        double qkr = qc5 + rr * (qc4 + rr * rr * (qc2 + rr * qc1));
        double n1, n2, n3;
        AB_X3(5, (fabs(qkr) < sqtol * ssmax) ? 0.0 : 1.0, n1, n2, n3);
        if (n1 + n2 + n3 == 0.0) {
            kflag = 2;
        } else {
            const double inv_ssmax = fast_div(1.0, ssmax);
            double sqmin = 0.0;
            AB_NOUNROLL
            for (int it = 1; it <= 3; it++) {
                const double qp = qc4 + rr * rr * (3.0 * qc2 + rr * 4.0 * qc1);
                double drr = 0.0;
                if (fabs(qp) > TINY_ * ssmax) drr = -fast_div(qkr, qp);
                const double rrc = rr + drr;
                double sv1, sv2, sv3;
                g.sync();   // slots 6..9 are re-used by the iterations
                AB_X3(6, rrc, sv1, sv2, sv3);
                const double p1 = qc5 + sv1 * (qc4 + sv1 * sv1 * (qc2 + sv1 * qc1));
                const double p2 = qc5 + sv2 * (qc4 + sv2 * sv2 * (qc2 + sv2 * qc1));
                const double p3 = qc5 + sv3 * (qc4 + sv3 * sv3 * (qc2 + sv3 * qc1));
                // m_j = max over the orders of |p_j| / ssmax: three values per publishing lane
                if (pub) { xb[28 + g.lane] = fabs(p1) * inv_ssmax; xb[32 + g.lane] = fabs(p2) * inv_ssmax; xb[36 + g.lane] = fabs(p3) * inv_ssmax; }
                g.sync();
                const double m1 = ab_max(xb[28], ab_max(xb[29], xb[30])), m2 = ab_max(xb[32], ab_max(xb[33], xb[34]));
                const double m3 = ab_max(xb[36], ab_max(xb[37], xb[38]));
                sqmin = m1 + 1.0;
                int kmin = 1;
                if (m1 < sqmin) { kmin = 1; sqmin = m1; }
                if (m2 < sqmin) { kmin = 2; sqmin = m2; }
                if (m3 < sqmin) { kmin = 3; sqmin = m3; }
                rr = (kmin == 1) ? sv1 : (kmin == 2) ? sv2 : sv3;
                if (sqmin < sqtol) {
                    kflag = 3;
                    break;
                }
                qkr = (kmin == 1) ? p1 : (kmin == 2) ? p2 : p3;
            }
            if (sqmin > sqtol) return -6;
        }
    }
    if (!(rr > rrcut)) return kflag;   // only a result > 3 acts, and that needs rr > rrcut
    const double rsa = s[1], rsb = s[2] * rr, rsc = s[3] * rr * rr;
    const double rsd = s[4] * rr * rr * rr;
    const double rd1a = rsa - rsb, rd1b = rsb - rsc, rd1c = rsc - rsd;
    const double rd2a = rd1a - rd1b, rd2b = rd1b - rd1c, rd3a = rd2a - rd2b;
    const bool small = fabs(rd1b) < TINY_ * smax;
    const double cest1 = small ? 1.0 : -fast_div(rd3a, rd1b);
    double w1, w2, w3;
    g.sync();
    AB_X3(0, (small || cest1 < TINY_ || cest1 > 4.0) ? 1.0 : 0.0, w1, w2, w3);
    if (w1 + w2 + w3 != 0.0) return -7;
    const double sigsq = s[3] + fast_div(rd2b, cest1 * (rr * rr));
    double sg1, sg2, sg3;
    AB_X3(1, sigsq, sg1, sg2, sg3);
    if (sg2 < TINY_) return -8;
    const double ratp = fast_div(sg3, sg2), ratm = fast_div(sg1, sg2);
    const double qfac1 = 0.25 * (q * q - 1.0), qfac2 = (q == 3) ? 1.0 : (q == 4) ? (2.0 / 3.0) : 0.5;
    const double bb = ratp * ratm - 1.0 - qfac1 * ratp;
    const double tem = 1.0 - qfac2 * bb;
    if (fabs(tem) < TINY_) return -8;
    const double rrb = fast_div(1.0, tem);
    if (fabs(rrb - rr) > rrtol) return -9;
    if (kflag == 1) kflag = 4;
    if (kflag == 2) kflag = 5;
    if (kflag == 3) kflag = 6;
    return kflag;
#undef AB_X3
}

// x^(1/n) for the four step-size formulas (eta = 1/((bias*err)^(1/n) + addon))
AB_HD double nth_root(double x, int n) {
#if defined(__CUDA_ARCH__) && !defined(AB_EXACT_ROOT)
    // Step-size ratios are heuristics that CVODES itself clips and thresholds (eta < 1.5 keeps h, eta <= etamax):
    // single-precision exp2/log2 (relative error ~1e-6 in eta) instead of the ~150-instruction double pow.
    return (double)exp2f(__log2f((float)x) * (1.0f / (float)n));
#else
    return pow(x, 1.0 / n);
#endif
}

// Dense LU with partial pivoting of I - gamma*J (J = S.savedJ) in shared memory: the fallback when the emitted
// static-pivot sparse LU rejects a pivot (SUNLinSol_Dense semantics). Returns false on an exactly singular matrix.
template <class M, int G, int C>
struct DenseRef {
    typename Layout<M, G, C>::Shared& S;
    AB_HD double& operator()(int r, int c) { return S.lu[r][c]; }
};

// In-place LU with partial pivoting of the matrix in S.lu (executed by one lane); false if exactly singular.
template <class M, int G, int C>
AB_HD bool dense_lu_inplace(typename Layout<M, G, C>::Shared& S) {
    constexpr int NX = M::NX;
    for (int kk = 0; kk < NX; ++kk) {
        int pr = kk;
        double best = fabs(S.lu[kk][kk]);
        for (int i = kk + 1; i < NX; ++i) {
            const double a = fabs(S.lu[i][kk]);
            if (a > best) { best = a; pr = i; }
        }
        S.piv[kk] = pr;
        if (!(best > 0.0)) return false;
        if (pr != kk)
            for (int j = 0; j < NX; ++j) {
                const double tmp = S.lu[kk][j];
                S.lu[kk][j] = S.lu[pr][j];
                S.lu[pr][j] = tmp;
            }
        const double inv = 1.0 / S.lu[kk][kk];
        for (int i = kk + 1; i < NX; ++i) {
            const double m = S.lu[i][kk] * inv;
            S.lu[i][kk] = m;
            if (m != 0.0)
                for (int j = kk + 1; j < NX; ++j) S.lu[i][j] -= m * S.lu[kk][j];
        }
    }
    return true;
}

template <class M, int G, int C>
AB_COLD bool cold_dense_factor(typename Layout<M, G, C>::Shared& S, double gamma) {
    Group<G> g;
    g.sync();
    int ok = 1;
    if (g.lane == 0) {
        DenseRef<M, G, C> A{S};
        M::jac_to_A(S.savedJ, gamma, A);
        ok = dense_lu_inplace<M, G, C>(S) ? 1 : 0;
    }
    ok = g.bcast(ok, 0);
    g.sync();
    return ok != 0;
}

// b <- A^{-1} b with the dense factors, one column (b lives in the caller's local memory)
template <class M, int G, int C>
AB_COLD void cold_dense_solve(const typename Layout<M, G, C>::Shared& S, double* b) {
    constexpr int NX = M::NX;
    for (int kk = 0; kk < NX; ++kk) {
        const int pr = S.piv[kk];
        const double tmp = b[kk];
        b[kk] = b[pr];
        b[pr] = tmp;
    }
    for (int i = 1; i < NX; ++i) {
        double s = b[i];
        for (int j = 0; j < i; ++j) s -= S.lu[i][j] * b[j];
        b[i] = s;
    }
    for (int i = NX - 1; i >= 0; --i) {
        double s = b[i];
        for (int j = i + 1; j < NX; ++j) s -= S.lu[i][j] * b[j];
        b[i] = s / S.lu[i][i];
    }
}

// ---------------------------------------------------------------------------------------------
// the solver
// ---------------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------------
// CVODES step control, shared by the lane-group solver (Bdf below: a condition per group of lanes, vectors in registers)
// and the block solver (bdf_block.cuh: a condition per thread block, vectors in global memory).  CRTP: D supplies the
// vector primitives
//   nx() gsync() leader()                       size / barrier of the cooperating threads / one elected thread
//   sldet()                                     cvSLdet on S.ssdat (sldet_serial() or a cooperative version)
//   set_ewt() wrms_max_zn(j, lj) wrms_max_up(cquot) stab_sums(lq, lqm1, s1, s2)
//   history_pass<PRED>(cmpl, resc) order_up(A1, ll) order_down(ll)
//   store_correction(save) store_correction_high() scale_zn1(f) init_vectors()
//   nls(nflag) rhs_to_zn1(scale) initial_step(tout, &h0, &nfe)
//   step_sync_join() step_sync_arrive() step_sync_leave()
//   and, if D::ROOTS: rcheck1() rcheck2() rcheck3(tout, one_step) collapse_to(t), members tlo, irfnd
// and everything here is scalar logic that every thread executes identically.
// ---------------------------------------------------------------------------------------------
// optional features of a derived solver (adjoint.cuh): quadrature variables that are corrected after the states have passed
// their error test (cvQuadNls) and take their own (D::QUAD), and a callback after every internal step (D::HIST: the forward
// pass of the adjoint method stores its solution there).  Solvers that do not declare them compile none of it.
template <class D, class = void> struct HasQuad { static constexpr bool value = false; };
template <class D> struct HasQuad<D, decltype((void)D::QUAD)> { static constexpr bool value = D::QUAD; };
template <class D, class = void> struct HasHist { static constexpr bool value = false; };
template <class D> struct HasHist<D, decltype((void)D::HIST)> { static constexpr bool value = D::HIST; };

template <class D, class SharedT>
struct BdfControl {
    SharedT& S;
    const SolverOptions& o;

    // step control (uniform scalars; CVodeMem fields of the same name). Arrays are only ever
    // indexed with compile-time constants (unrolled, predicated loops) so that they stay in registers.
    double tn, h, hscale, hprime, eta, etamax, etaq, etaqm1, etaqp1, hu;
    double tau[LMAX + 1], l[LMAX], tq[6];
    double rl1, gamma, gammap, gamrat, crate, delp, acnrm, saved_tq5;
    double crateS, acnrmS;   // staggered corrector (D::STG): convergence rate and correction norm of the sensitivity solve
    double tstop, tretlast;
    int q, qprime, qwait, L, qu;
    int nst, nstlp, nstlj, nscon, nfe, netf, ncfn, nni, nsetups, nje;
    int convfail, ss_head;
    bool jcur, tstopset;
    bool pending;          // the accepted correction has not been added to the history yet: see history_pass
    bool reinit_pending;   // AMICI's CVodeSolver::reinit happened: the next call starts with reInit_post_process
    // A step in progress (the locals of cvStep) and a solver call in progress (the locals of CVode): members, so that the
    // nested driver (solve_to / step) and the flat one (flat.cuh: one attempt per iteration of a warp-uniform loop) run the
    // same pieces: solve_begin, step_pre, step_begin, attempt, step_finish, step_post.
    double st_saved_t, st_dsm;
    int st_ncf, st_nef, st_nflag, st_dq;
    int st_ncfS, st_nefS;    // staggered corrector: failure counters of the sensitivity stage (the locals ncfS / nefS of cvStep)
    bool st_restore, st_resc, st_reload;
    int so_nstloc, so_mxstep, so_nst_saved;
    bool so_post, so_one_step;
    double so_tout;
    static constexpr int CONTINUE = 1 << 20;   // "no return code yet" of the pieces

    AB_HD BdfControl(SharedT& s, const SolverOptions& opt) : S(s), o(opt) {}
    AB_HD D& self() { return static_cast<D&>(*this); }
    AB_HD const D& self() const { return static_cast<const D&>(*this); }

    // add a pending correction to the history (anything that reads zn other than through the pending-aware
    // accessors — outputs, root finding, order changes — needs this first)
    AB_HD void flush_pending() {
        if (pending) {
            self().template history_pass<0>(true, false);
            pending = false;
        }
    }

    // cvIncreaseBDF (cvodes.c:6256) for deltaq = +1, cvDecreaseBDF (:6363) for deltaq = -1 (cvAdjustOrder :6134)
    AB_HD void adjust_order(int deltaq) {
        if (q == 2 && deltaq != 1) return;
        double ll[LMAX + 1];
        AB_UNROLL
        for (int i = 0; i <= LMAX; ++i) ll[i] = 0.0;
        ll[2] = 1.0;
        if (deltaq == 1) {
            double alpha0 = -1.0, alpha1 = 1.0, prod = 1.0, xiold = 1.0, hsum = hscale;
            AB_UNROLL
            for (int j = 1; j < QMAX; ++j) {
                if (j < q) {
                    hsum += tau[j + 1];
                    const double xi = fast_div(hsum, hscale);
                    prod *= xi;
                    alpha0 -= 1.0 / (j + 1);
                    alpha1 += fast_div(1.0, xi);
                    AB_UNROLL
                    for (int i = j + 2; i >= 2; --i) ll[i] = ll[i] * xiold + ll[i - 1];
                    xiold = xi;
                }
            }
            const double A1 = fast_div(-alpha0 - alpha1, prod);
            self().order_up(A1, ll);   // zn[L] = A1 * zn[qmax] (saved correction), zn[2..q] += ll[j] * zn[L]
        } else if (deltaq == -1) {
            double hsum = 0.0;
            AB_UNROLL
            for (int j = 1; j <= QMAX - 2; ++j) {
                if (j <= q - 2) {
                    hsum += tau[j];
                    const double xi = fast_div(hsum, hscale);
                    AB_UNROLL
                    for (int i = j + 2; i >= 2; --i) ll[i] = ll[i] * xi + ll[i - 1];
                }
            }
            if (q > 2) self().order_down(ll);   // zn[2..q-1] -= ll[j] * zn[q]
        }
    }

    // cvSetBDF + cvSetTqBDF + cvSet (cvodes.c:6788-6892, :6617)
    AB_HD void set_coeffs() {
        double alpha0 = -1.0, alpha0_hat = -1.0, xi_inv = 1.0, xistar_inv = 1.0, hsum = h;
        l[0] = l[1] = 1.0;
        AB_UNROLL
        for (int i = 2; i < LMAX; ++i) l[i] = 0.0;
        if (q > 1) {
            AB_UNROLL
            for (int j = 2; j < QMAX; ++j) {
                if (j < q) {
                    hsum += tau[j - 1];
                    xi_inv = fast_div(h, hsum);
                    alpha0 -= 1.0 / j;
                    AB_UNROLL
                    for (int i = j; i >= 1; --i) l[i] += l[i - 1] * xi_inv;
                }
            }
            alpha0 -= (q == 2) ? 0.5 : (q == 3) ? (1.0 / 3.0) : (q == 4) ? 0.25 : 0.2;
            xistar_inv = -l[1] - alpha0;
            hsum += tau_at(q - 1);
            xi_inv = fast_div(h, hsum);
            alpha0_hat = -l[1] - xi_inv;
            AB_UNROLL
            for (int i = QMAX; i >= 1; --i)
                if (i <= q) l[i] += l[i - 1] * xistar_inv;
        }
        mirror_l();
        const double lq = l_at(q);
        const double A1 = 1.0 - alpha0_hat + alpha0;
        const double A2 = 1.0 + q * A1;
        tq[2] = fabs(fast_div(A1, alpha0 * A2));
        tq[5] = fabs(fast_div(A2 * xistar_inv, lq * xi_inv));
        if (qwait == 1) {
            if (q > 1) {
                const double Cc = fast_div(xistar_inv, lq);
                const double A3 = alpha0 + 1.0 / q;
                const double A4 = alpha0_hat + xi_inv;
                const double Cpinv = fast_div(1.0 - A4 + A3, A3);
                tq[1] = fabs(Cc * Cpinv);
            } else {
                tq[1] = 1.0;
            }
            hsum += tau_at(q);
            xi_inv = fast_div(h, hsum);
            const double A5 = alpha0 - (1.0 / (q + 1));
            const double A6 = alpha0_hat - xi_inv;
            const double Cppinv = fast_div(1.0 - A6 + A5, A2);
            tq[3] = fabs(fast_div(Cppinv, xi_inv * (q + 2) * A5));
        }
        tq[4] = fast_div(0.1, tq[2]);  // nlscoef = CORTES = 0.1
        rl1 = fast_div(1.0, l[1]);
        gamma = h * rl1;
        if (nst == 0) gammap = gamma;
        gamrat = (nst > 0) ? fast_div(gamma, gammap) : 1.0;
    }

    // cvCompleteStep (cvodes.c:7634): counters and step-size history now; the update of the Nordsieck array is deferred
    // to the next history_pass (pending), the correction is kept in registers and in S.acorbuf (flat layout)
    AB_HD void complete_step() {
        nst++;
        nscon++;
        hu = h;
        qu = q;
        AB_UNROLL
        for (int i = QMAX; i >= 2; --i)
            if (i <= q) tau[i] = tau[i - 1];
        if (q == 1 && nst > 1) tau[2] = tau[1];
        tau[1] = h;
        mirror_tau();
        qwait--;
        const bool save = (qwait == 1 && q != QMAX);
        self().store_correction(save);   // acorbuf = acor, and zn[qmax] = acor if save
        if (save) saved_tq5 = tq[5];
        pending = true;
    }

    // tau[] and l[] live in registers and are indexed with compile-time constants; where the index is the run-time order
    // (tau[q], l[q]), a select chain over the array costs ~16 instructions per use.  Every thread writes the same values
    // to S.ctl and reads back its own writes, so no synchronisation is needed.
    AB_HD void mirror_tau() {
        AB_UNROLL
        for (int i = 1; i <= LMAX; ++i) S.ctl[i] = tau[i];
    }
    AB_HD void mirror_l() {
        AB_UNROLL
        for (int i = 0; i < LMAX; ++i) S.ctl[8 + i] = l[i];
    }
    AB_HD double tau_at(const int i) const { return S.ctl[i]; }
    AB_HD double l_at(const int i) const { return S.ctl[8 + i]; }

    // cvSLdet on the ring buffer, every thread computing all three orders (the solvers may override this)
    AB_HD int sldet_serial() const {
        double s[6][4], ratc[5][4];
        AB_UNROLL
        for (int i = 1; i <= 5; ++i) {
            int idx = ss_head + i - 1;
            idx = idx >= 5 ? idx - 5 : idx;
            AB_UNROLL
            for (int kk = 1; kk <= 3; ++kk) {
                s[i][kk] = S.ssdat[idx][kk];
                if (i <= 4) ratc[i][kk] = S.ssdat[idx][4 + kk];
            }
        }
        const int rc = bdf_sldet(s, ratc, q);
#if defined(AB_SLDET_RECORD) && !defined(__CUDA_ARCH__)
        sldet_record(&S.ssdat[0][0], ss_head, q, rc);   // tests/cpu_emul: harvest real ring buffers as test vectors
#endif
        return rc;
    }

    // cvSetEta (cvodes.c:7798) with eta_min_fx = 0, eta_max_fx = 1.5, hmin = 0
    AB_HD void set_eta() {
        if (eta > 0.0 && eta < 1.5) {
            eta = 1.0;
            hprime = h;
        } else {
            if (eta >= 1.5) {
                eta = ab_min(eta, etamax);
                eta /= ab_max(1.0, fabs(h) * o.hmax_inv * eta);
            } else {
                eta = ab_max(eta, 0.1);
            }
            hprime = h * eta;
            if (qprime < q) nscon = 0;
        }
    }

    // cvPrepareNextStep (cvodes.c:7747) incl. cvComputeEtaqm1/qp1 and cvChooseEta
    AB_HD void prepare_next_step(double dsm) {
        if (etamax == 1.0) {
            qwait = qwait > 2 ? qwait : 2;
            qprime = q;
            hprime = h;
            eta = 1.0;
            return;
        }
        etaq = fast_div(1.0, nth_root(6.0 * dsm, L) + 1e-6);
        if (qwait != 0) {
            eta = etaq;
            qprime = q;
        } else {
            qwait = 2;
            etaqm1 = 0.0;
            if (q > 1) {
                const double ddn = self().wrms_max_zn(q, l_at(q)) * tq[1];
                etaqm1 = fast_div(1.0, nth_root(6.0 * ddn, q) + 1e-6);
            }
            etaqp1 = 0.0;
            if (q != QMAX && saved_tq5 != 0.0) {
                double ratio = fast_div(h, tau[2]), pw = 1.0;
                AB_NOUNROLL
                for (int i = 0; i < L; ++i) pw *= ratio;  // SUNRpowerI(h/tau[2], L)
                const double cquot = fast_div(tq[5], saved_tq5) * pw;
                const double dup = self().wrms_max_up(cquot) * tq[3];   // norm of acor - cquot * zn[qmax]
                etaqp1 = fast_div(1.0, nth_root(10.0 * dup, L + 1) + 1e-6);
            }
            // cvChooseEta (cvodes.c:7935)
            const double etam = ab_max(etaqm1, ab_max(etaq, etaqp1));
            if (etam > 0.0 && etam < 1.5) {
                eta = 1.0;
                qprime = q;
            } else if (etam == etaq) {
                eta = etaq;
                qprime = q;
            } else if (etam == etaqm1) {
                eta = etaqm1;
                qprime = q - 1;
            } else {
                eta = etaqp1;
                qprime = q + 1;
                self().store_correction_high();   // zn[qmax] = acor
            }
        }
        set_eta();
    }

    // cvBDFStab (cvodes.c:8158); the 5-deep history is a ring buffer in shared memory (no shifting)
    AB_HD void bdf_stab() {
        if (q >= 3) {
            const double factorial = (q == 3) ? 2.0 : (q == 4) ? 6.0 : 24.0;   // (q-1)!
            const double sq = factorial * q * (q + 1) * fast_div(acnrm, ab_max(tq[5], 1e-10));
            // WRMS of the state column only (lane 0 owns it)
            double s1 = 0.0, s2 = 0.0;   // valid on the leader
            {
                const double lq = pending ? l_at(q) : 0.0, lqm1 = pending ? l_at(q - 1) : 0.0;   // deferred completion
                self().stab_sums(lq, lqm1, s1, s2);
            }
            const double f2 = factorial * factorial * (1.0 / self().nx());
            ss_head = (ss_head == 0) ? 4 : ss_head - 1;
            self().gsync();
            if (self().leader()) {
                const int older = (ss_head == 4) ? 0 : ss_head + 1;
                const double v1 = f2 * s2, v2 = f2 * q * q * s1, v3 = sq * sq;
                S.ssdat[ss_head][1] = v1;            // sqm2^2
                S.ssdat[ss_head][2] = v2;            // sqm1^2
                S.ssdat[ss_head][3] = v3;
                // ratios to the previous entry: cvSLdet recomputes these 12 quotients on every call, 9 of them unchanged
                S.ssdat[ss_head][5] = fast_div(v1, S.ssdat[older][1]);
                S.ssdat[ss_head][6] = fast_div(v2, S.ssdat[older][2]);
                S.ssdat[ss_head][7] = fast_div(v3, S.ssdat[older][3]);
            }
            self().gsync();
        }
        if (qprime >= q) {
            if (q >= 3 && nscon >= q + 5) {
                const int ldflag = self().sldet();
                if (ldflag > 3) {
                    qprime = q - 1;
                    eta = etaqm1;
                    eta = ab_min(eta, etamax);
                    eta = eta / ab_max(1.0, fabs(h) * o.hmax_inv * eta);
                    hprime = h * eta;
                }
            }
        } else {
            nscon = 0;
        }
    }

    // CVodeSolver::reinit + sens_reinit (reference src/solver_cvodes.cpp:471-503, :586-616): the caller has written the
    // post-event state and sensitivities to zn[0]; restart at order 1 keeping counters, step-size history and linear
    // solver data; the next solve_to call begins with reInit_post_process (:517-565).
    AB_HD void amici_reinit(double t0) {
        tn = t0;
        q = 1;
        L = 2;
        qwait = 2;
        qu = 0;
        hu = 0.0;
        reinit_pending = true;
    }

    // -----------------------------------------------------------------------------------------
    // cvStep (cvodes.c:5800) with cvHandleNFlag (:7309) and cvDoErrorTest (:7461) folded in: every
    // retry path sets flags that are executed at the single history-manipulation site at the loop top.
    // -----------------------------------------------------------------------------------------
    AB_HD void step_begin() {
        st_saved_t = tn;
        st_ncf = st_nef = 0;
        st_ncfS = st_nefS = 0;
        st_nflag = FIRST_CALL;
        st_dsm = 0.0;
        st_restore = st_resc = st_reload = false;
        st_dq = 0;
        if (nst > 0 && hprime != h) {  // cvAdjustParams
            st_dq = qprime - q;
            st_resc = true;
        }
    }
    // One attempt of the step: CONTINUE = repeat with the flags set here, AB_SUCCESS = passed the error test, < 0 = failure
    AB_HD int attempt() {
        // the one history-manipulation site: [restore] -> [complete] -> [order change] -> [rescale] -> predict
        if (st_restore) {
            self().template history_pass<-1>(false, false);
            tn = st_saved_t;
        }
        if (st_dq != 0) {
            flush_pending();   // cvAdjustOrder works on the completed history (rare)
            adjust_order(st_dq);
            q += st_dq;
            L = q + 1;
            qwait = L;
            st_dq = 0;
        }
        if (st_resc) {   // scalar part of cvRescale; the history is scaled in the pass below
            h = hscale * eta;
            hscale = h;
            nscon = 0;
            // a step size that has underflowed to zero (or is no number) after repeated failures ends the solve: with hmin = 0
            // CVODES runs into 0 / 0 in gamrat here and gives up after its next ten corrector failures
            if (!(fabs(h) > 0.0)) return (st_nflag == PREV_ERR_FAIL) ? AB_ERR_FAILURE : AB_CONV_FAILURE;
        }
        if (st_reload) {  // order-1 restart: zn[1] = h * f(tn, zn[0])
            const bool ok = self().rhs_to_zn1(h);
            nfe++;
            if (!ok) return AB_UNREC_RHSFUNC_ERR;
        }
        tn += h;  // cvPredict
        if (tstopset && (tn - tstop) * h > 0.0) tn = tstop;
        self().template history_pass<1>(pending, st_resc);
        pending = false;
        st_restore = st_resc = st_reload = false;
        set_coeffs();
        st_nflag = self().nls(st_nflag);
        if (st_nflag != AB_SUCCESS) {  // cvHandleNFlag
            ncfn++;
            st_ncf++;
            etamax = 1.0;
            if (st_ncf == o.max_conv_fails) {
                tn = st_saved_t;
                return (st_nflag == NLS_RHS_RECVR) ? AB_REPTD_RHSFUNC_ERR : AB_CONV_FAILURE;
            }
            eta = 0.25;  // max(eta_cf, hmin/|h|), hmin = 0
            st_nflag = PREV_CONV_FAIL;
            st_restore = st_resc = true;
            return CONTINUE;
        }
        st_dsm = acnrm * tq[2];  // cvDoErrorTest
        if (st_dsm > 1.0) return error_test_failed(st_nef, st_dsm, true);
        if constexpr (D::STG) {
            if (o.ism == 2 && self().has_sens()) return staggered_stage();
        }
        if constexpr (HasQuad<D>::value) return quadrature_stage();
        return AB_SUCCESS;
    }
    // Quadrature variables (cvStep, cvodes.c:5870-5905): after the states have converged and passed their error test the
    // quadratures are corrected explicitly (cvQuadNls: acorQ = rl1 * (h * fQ(tn, y) - znQ[1])) and take their own error test
    // (errconQ, which AMICI switches on for the backward problem: reference src/solver.cpp:674-691); dsm = max of both.
    AB_HD int quadrature_stage() {
        st_ncf = st_nef = 0;
        if (!self().quad_correct()) {   // fQ failed recoverably: cvHandleNFlag
            st_ncfS++;
            etamax = 1.0;
            if (st_ncfS == o.max_conv_fails) {
                tn = st_saved_t;
                return AB_CONV_FAILURE;
            }
            eta = 0.25;
            st_nflag = PREV_CONV_FAIL;
            st_restore = st_resc = true;
            return CONTINUE;
        }
        const double dsmQ = self().quad_norm() * tq[2];
        if (dsmQ > 1.0) return error_test_failed(st_nefS, dsmQ, false);
        if (dsmQ > st_dsm) st_dsm = dsmQ;
        return AB_SUCCESS;
    }
    // cvDoErrorTest (cvodes.c:7461) after a failed test: count, restore, shrink the step / lower the order / restart at order 1.
    // `nef` is the failure counter of the stage (states, or sensitivities of the staggered corrector).
    AB_HD int error_test_failed(int& nef, const double dsm, const bool count) {
        nef++;
        if (count) netf++;   // ReturnData reports the state stage's counter (CVodeGetNumErrTestFails)
        st_nflag = PREV_ERR_FAIL;
        st_restore = true;
        if (nef == o.max_err_test_fails) {
            tn = st_saved_t;
            return AB_ERR_FAILURE;
        }
        etamax = 1.0;
        if (nef <= 3) {  // MXNEF1
            eta = 1.0 / (nth_root(6.0 * dsm, L) + 1e-6);
            eta = ab_max(0.1, eta);
            if (nef >= 2) eta = ab_min(eta, 0.2);
            st_resc = true;
        } else if (q > 1) {  // force an order reduction
            eta = 0.1;
            st_dq = -1;
            st_resc = true;
        } else {  // already at order 1: restart from scratch with h*eta
            eta = 0.1;
            h *= eta;
            hscale = h;
            qwait = 10;  // LONG_WAIT
            nscon = 0;
            st_reload = true;
        }
        return CONTINUE;
    }
    // The sensitivity stage of the STAGGERED corrector (cvStep, cvodes.c:5936-5990; cvStgrNls :7167-7210; cvodes_nls_stg.c): the
    // states have converged and passed their error test; now the right-hand side is evaluated at the converged state, the
    // sensitivities of all parameters are corrected together by their own Newton iteration (same linear system, no setup
    // unless the iteration fails with stale Jacobian data), and they take their own error test with their own counters.
    AB_HD int staggered_stage() {
        st_ncf = st_nef = 0;   // "reset counters for states"
        nfe++;
        if (!self().stg_rhs_at_converged_state()) {
            // a recoverable failure of f at the converged state: CVODES predicts again without restoring the history (:5950-5954)
            st_nflag = PREV_CONV_FAIL;
            return CONTINUE;
        }
        const int sflag = self().nls_stg();
        if (sflag != AB_SUCCESS) {   // cvHandleNFlag with ncfS / ncfnS
            st_ncfS++;
            etamax = 1.0;
            if (st_ncfS == o.max_conv_fails) {
                tn = st_saved_t;
                return (sflag == NLS_RHS_RECVR) ? AB_REPTD_SRHSFUNC_ERR : AB_CONV_FAILURE;
            }
            eta = 0.25;
            st_nflag = PREV_CONV_FAIL;
            st_restore = st_resc = true;
            return CONTINUE;
        }
        const double dsmS = acnrmS * tq[2];   // errconS = TRUE in AMICI (reference src/solver.cpp:641-656)
        if (dsmS > 1.0) return error_test_failed(st_nefS, dsmS, false);
        if (dsmS > st_dsm) st_dsm = dsmS;     // cvPrepareNextStep sees the larger of the two
        return AB_SUCCESS;
    }
    AB_HD void step_finish() {
        complete_step();
        prepare_next_step(st_dsm);
        if (o.stldet) bdf_stab();
        etamax = 10.0;  // eta_max_es == eta_max_gs == 10
    }
    AB_HD int step() {
        step_begin();
        for (;;) {
            // Step alignment is per ATTEMPT: a warp that repeats a step after a failed error or convergence test joins
            // the others at their next step instead of making them wait a whole attempt.
            self().step_sync_arrive();
            const int r = attempt();
            if (r == CONTINUE) continue;
            if (r != AB_SUCCESS) return r;
            break;
        }
        step_finish();
        return AB_SUCCESS;
    }

    // -----------------------------------------------------------------------------------------
    // solver (re-)initialisation: Solver::setup -> CVodeInit/CVodeSensInit1 (reference src/solver.cpp:193-265);
    // zn[0] must hold the initial state and sensitivities
    // -----------------------------------------------------------------------------------------
    AB_HD void init(double t0) {
        tn = t0;
        q = 1; L = 2; qwait = 2; qprime = 1; qu = 0;
        etamax = 1e4;  // eta_max_fs
        h = hscale = hprime = hu = 0.0;
        eta = 1.0;
        nst = nstlp = nstlj = nscon = nfe = netf = ncfn = nni = nsetups = nje = 0;
        AB_UNROLL
        for (int i = 0; i <= LMAX; ++i) tau[i] = 0.0;
        AB_UNROLL
        for (int i = 0; i < LMAX; ++i) l[i] = 0.0;
        AB_UNROLL
        for (int i = 0; i < 6; ++i) tq[i] = 0.0;
        mirror_tau();
        mirror_l();
        rl1 = gamma = gammap = gamrat = 1.0;
        crate = 1.0; delp = 0.0; acnrm = 0.0; saved_tq5 = 0.0;
        crateS = 1.0; acnrmS = 0.0;
        etaq = etaqm1 = etaqp1 = 0.0;
        tstopset = false; tstop = 0.0; tretlast = t0;
        convfail = CONVFAIL_NONE; jcur = false;         reinit_pending = false;
        pending = false;
        self().init_vectors();   // acor = ewt = 0, root-finding state
        self().gsync();
        ss_head = 0;
        if (self().leader())
            for (int i = 0; i < 5; ++i)
                for (int kk = 0; kk < 8; ++kk) S.ssdat[i][kk] = 0.0;
        self().gsync();
        self().set_ewt();  // cvInitialSetup
    }

    // -----------------------------------------------------------------------------------------
    // CVode(tout, CV_NORMAL) with tstop = tout, as AMICI's Solver::run drives it
    // (reference src/solver.cpp:141-160; cvodes.c:2890-3650): returns AB_TSTOP_RETURN or an error.
    // one_step = true is CVode(tout, CV_ONE_STEP) without a stop time, as Solver::step drives it during
    // steady-state simulation (reference src/solver.cpp:162-178, src/forwardproblem.cpp:197-206): tout only
    // gives the direction and the scale of the initial step; returns AB_SUCCESS after one internal step.
    // Both modes share this function so that the (inlined) step has a single call site.
    // -----------------------------------------------------------------------------------------
    AB_HD int solve_begin(double tout, bool one_step = false) {
        constexpr bool ROOTS = D::ROOTS;
        // Solver::apply_max_num_steps (reference src/solver.cpp:107-119)
        if (o.maxsteps <= nst) return AB_ERROR;
        so_mxstep = o.maxsteps - nst;
        so_tout = tout;
        so_one_step = one_step;
        if (!one_step) {
            tstop = tout;
            tstopset = true;
        } else if (!(D::EVENTS)) {
            tstopset = false;
        } else if (reinit_pending) {
            // Solver::step never sets a stop time, but reInit_post_process does (CVodeSetStopTime(tout), reference
            // src/solver_cvodes.cpp:538) and CVODES keeps it until it is reached: the steady-state simulation of a model
            // with events stops at the tout of the first step after every event
            tstop = tout;
            tstopset = true;
        }
        // reInit_post_process (reference src/solver_cvodes.cpp:517-565): after an event the first step is taken as a
        // "first call" (nst = 0) in one-step mode with tstop = tout, then the step counter is restored
        so_post = D::EVENTS ? reinit_pending : false;
        so_nst_saved = 0;
        if (so_post) {
            reinit_pending = false;
            so_nst_saved = nst;
            nst = 0;
            self().set_ewt();   // cvInitialSetup
        }
        if (nst == 0) {
            // first call: zn[1] = f(t0, y0) (cvodes.c first-call block), initial step, zn[1] *= h
            const bool ok = self().rhs_to_zn1(1.0);
            nfe++;
            if (!ok) return AB_FIRST_RHSFUNC_ERR;
            if (tstopset && (tstop - tn) * (tout - tn) <= 0.0) return AB_ILL_INPUT;
            double h0 = 0.0;
            int nfe_hin = 0;
            const int hflag = self().initial_step(tout, &h0, &nfe_hin);
            nfe += nfe_hin;
            if (hflag != AB_SUCCESS) return hflag;
            h = h0;
            const double rh = fabs(h) * o.hmax_inv;
            if (rh > 1.0) h /= rh;
            if (tstopset && (tn + h - tstop) * h > 0.0) h = (tstop - tn) * (1.0 - 4.0 * DBL_EPSILON);
            hscale = h;
            hprime = h;
            self().scale_zn1(h);
            if constexpr (ROOTS) self().rcheck1();
        } else {
            const double troundoff = 100.0 * DBL_EPSILON * (fabs(tn) + fabs(h));
            if constexpr (ROOTS) {   // cvodes.c "check for a root in the last step taken" on entry
                const int irfndp = self().irfnd;
                const int rc = self().rcheck2();
                if (rc == 3) return AB_ILL_INPUT;
                if (rc == 1) { tretlast = self().tlo; self().collapse_to(self().tlo); return AB_ROOT_RETURN; }
                if (fabs(tn - tretlast) > troundoff) {
                    if (self().rcheck3(tout, one_step)) {
                        self().irfnd = 1;
                        tretlast = self().tlo;
                        self().collapse_to(self().tlo);
                        return AB_ROOT_RETURN;
                    }
                    self().irfnd = 0;
                    if (irfndp == 1 && one_step) { tretlast = tn; return AB_SUCCESS; }
                }
            }
            if (tstopset) {
                if (fabs(tn - tstop) <= troundoff) {
                    tretlast = tstop;
                    tstopset = false;
                    return AB_TSTOP_RETURN;
                } else if ((tn + hprime - tstop) * h > 0.0) {
                    hprime = (tstop - tn) * (1.0 - 4.0 * DBL_EPSILON);
                    eta = fast_div_1ulp(hprime, h);   // tn + h * eta must land within 100 ulp of tstop: a 2^-40 quotient misses it whenever h > 0.02 |tn| (one extra step per output time)
                }
            }
        }
        so_nstloc = 0;
        return CONTINUE;
    }
    // top of the CVode step loop: error weights, step budget, tolsf test
    AB_HD int step_pre() {
#if AB_FULL
        // Solver::time_exceeded (reference src/solver.cpp:1056-1069; checked every 500 right-hand sides there, every 32 steps here)
        if (o.maxtime > 0.0 && (nst & 31) == 0 && self().time_exceeded()) { tretlast = tn; return AB_MAX_TIME_EXCEEDED; }
#endif
        if (nst > 0) self().set_ewt();
        if (so_nstloc >= so_mxstep) { tretlast = tn; return AB_TOO_MUCH_WORK; }
        // cvodes.c "tolsf" test: u * max-WRMS(zn[0]) > 1.  |y_i| * ewt_i < 1/rtol by construction of the weights, so the
        // norm is below 1/rtol and the test can only fire when rtol is of the order of the unit round-off.
        if (!(o.rtol > 8.0 * DBL_EPSILON && o.rtol_fsa > 8.0 * DBL_EPSILON)) {
            const double nrm = self().wrms_max_zn(0, 1.0);
            if (DBL_EPSILON * nrm > 1.0) { tretlast = tn; return AB_TOO_MUCH_ACC; }
        }
        return CONTINUE;
    }
    // after a successful step: root check and stop logic; CONTINUE = take another step
    AB_HD int step_post() {
        constexpr bool ROOTS = D::ROOTS;
        so_nstloc++;
        double troundoff = 100.0 * DBL_EPSILON * (fabs(tn) + fabs(h));
        const bool at_tstop = tstopset && fabs(tn - tstop) <= troundoff;
        if (at_tstop) tn = tstop;
        if constexpr (ROOTS) {
            flush_pending();   // the root finder interpolates on the completed history
            if (self().rcheck3(so_tout, so_one_step || so_post)) {
                self().irfnd = 1;
                tretlast = self().tlo;
                // a root right after re-initialisation is an error in the reference (src/solver_cvodes.cpp:531-540)
                if (so_post) { nst = so_nst_saved + 1; return AB_ERROR; }
                self().collapse_to(self().tlo);
                return AB_ROOT_RETURN;
            }
        }
        if (D::EVENTS && so_post) {
            // end of reInit_post_process: anything but CV_SUCCESS throws there; then the regular call starts
            nst = so_nst_saved + 1;
            so_post = false;
            so_nstloc = 0;
            tretlast = tn;
            if (at_tstop) return AB_ERROR;
            if ((tn + hprime - tstop) * h > 0.0) {
                hprime = (tstop - tn) * (1.0 - 4.0 * DBL_EPSILON);
                eta = fast_div_1ulp(hprime, h);   // tn + h * eta must land within 100 ulp of tstop: a 2^-40 quotient misses it whenever h > 0.02 |tn| (one extra step per output time)
            }
            if constexpr (ROOTS) {
                const int rc2 = self().rcheck2();
                if (rc2 == 3) return AB_ILL_INPUT;
                if (rc2 == 1) { tretlast = self().tlo; self().collapse_to(self().tlo); return AB_ROOT_RETURN; }
            }
            return CONTINUE;
        }
        if (so_one_step && !(D::EVENTS && tstopset)) { tretlast = tn; return AB_SUCCESS; }
        if (at_tstop) {
            tretlast = tstop;
            tstopset = false;
            return AB_TSTOP_RETURN;
        } else if ((tn + hprime - tstop) * h > 0.0) {
            hprime = (tstop - tn) * (1.0 - 4.0 * DBL_EPSILON);
            eta = fast_div_1ulp(hprime, h);   // tn + h * eta must land within 100 ulp of tstop: a 2^-40 quotient misses it whenever h > 0.02 |tn| (one extra step per output time)
        }
        if (D::EVENTS && so_one_step) { tretlast = tn; return AB_SUCCESS; }   // CV_ONE_STEP with a stop time pending (cvodes.c: the tstop test comes first)
        return CONTINUE;
    }
    AB_HD int solve_to(double tout, bool one_step = false) {
        int rc = solve_begin(tout, one_step);
        if (rc != CONTINUE) return rc;
        // Step loop.  Every exit goes through the epilogue below: a pending completion is written to the history (the
        // caller reads zn[0]) and the warp leaves the step-alignment group.
        self().step_sync_join();
        for (;;) {
            rc = step_pre();
            if (rc != CONTINUE) break;
            const int kflag = step();
            if (kflag != AB_SUCCESS) { tretlast = tn; rc = kflag; break; }
            rc = step_post();
            if constexpr (HasHist<D>::value) { if (rc >= 0) self().on_step(); }   // CVodeF: every internal step is a data point of the interpolation (cvodea.c:395-560)
            if (rc != CONTINUE) break;
        }
        self().step_sync_leave();   // first: the other warps of the group must not wait for the flush below
        flush_pending();
        return rc;
    }
};

// ---------------------------------------------------------------------------------------------
// the lane-group solver: one condition per group of G lanes, C columns per lane, work vectors in registers
// ---------------------------------------------------------------------------------------------
// REINIT: compile CVodeSolver::reinit + reInit_post_process also for a model without events (the condition driver
// re-initialises the solver after a presimulation)
template <class M, int G, int C, bool REINIT = false>
struct Bdf : BdfControl<Bdf<M, G, C, REINIT>, typename Layout<M, G, C>::Shared> {
    static constexpr int NX = M::NX;
    static constexpr int NCOLS = Layout<M, G, C>::NCOLS;
    static constexpr bool ROOTS = M::NE_SOLVER > 0;   // the integrator tracks event triggers
    static constexpr bool EVENTS = M::NE > 0 || REINIT;   // AMICI re-initialises the solver after events
    // The STAGGERED sensitivity corrector (InternalSensitivityMethod::staggered, reference include/amici/defines.h:181) is compiled
    // into the full-featured kernel variant (pre-equilibration / presimulation) and into event models' kernels only, so that
    // the headline kernel's step loop does not grow.
    static constexpr bool STG = REINIT || M::NE > 0;
    using Shared = typename Layout<M, G, C>::Shared;
    using Eval = ModelEval<M, C>;
    using Base = BdfControl<Bdf<M, G, C, REINIT>, Shared>;
    // names of the dependent base used by the vector code below
    using Base::S; using Base::o;
    using Base::tn; using Base::h; using Base::hscale; using Base::hprime; using Base::eta; using Base::etamax;
    using Base::tau; using Base::l; using Base::tq; using Base::rl1; using Base::gamma; using Base::gammap;
    using Base::gamrat; using Base::crate; using Base::delp; using Base::acnrm; using Base::q; using Base::L;
    using Base::nst; using Base::nstlp; using Base::nstlj; using Base::nfe; using Base::nni; using Base::nsetups;
    using Base::nje; using Base::convfail; using Base::jcur; using Base::pending; using Base::tstop;
    using Base::tstopset; using Base::tretlast;

    Group<G> g;
    int ncols;  // active columns: 1 + nplist (1 without sensitivities)
    int col_lo = 0;   // first active column: 1 while the staggered corrector iterates on the sensitivities alone

    // per-lane column registers
    double acor[C][NX];
    double ewt[C][NX];

    // root finding (CVodeRootInit, cvodes.c:2674; cvRcheck1/2/3, cvRootfind :8576-9100) for the NES solver-tracked
    // event triggers; all group-uniform, arrays only indexed with compile-time constants
    static constexpr int NES = M::NE_SOLVER > 0 ? M::NE_SOLVER : 1;
    double glo[NES], ghi[NES], grout[NES];
    bool gactive[NES];
    int iroots[NES];
    double tlo, thi, trout, ttol;
    int irfnd, nge;

    // Step alignment of the warps of one block (G == 32 only): see step_sync_* below.
    volatile int* sync_pos = nullptr;
    int sync_n = 0, sync_w = 0, sync_gen = 0;

    bool dense_lu;

    AB_HD Bdf(Shared& s, const SolverOptions& opt, int ncols_) : Base(s, opt), ncols(ncols_) {}

    // -----------------------------------------------------------------------------------------
    // primitives BdfControl calls (see the list at its top)
    // -----------------------------------------------------------------------------------------
    AB_HD int nx() const { return NX; }
    AB_HD bool time_exceeded() const {
#if defined(__CUDA_ARCH__)
        long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        return (double)(now - S.t_start_ns) * 1e-9 > o.maxtime;
#else
        return false;   // the host emulation has no clock of its own
#endif
    }
    AB_HD int sldet() const {
#if defined(__CUDA_ARCH__)
        // one condition per warp: the three orders in three lanes (a shorter dependent chain on the pace-setting path of the
        // aligned warps).  Smaller groups share their warp with other conditions: every collective of the lane version is then
        // a divergence check plus a serialised region per group, and the serial form is both fewer instructions and less code.
        if (G == 32) return bdf_sldet_lanes<G>(g, S.ssdat, Base::ss_head, q);
#if AB_SLDET_SVC && !defined(AB_SLDET_SVC_LOCAL)
        if (G == 16) return sldet_remote();
#endif
        if (G >= 4) return bdf_sldet_xchg<G>(g, S.ssdat, S.sld, Base::ss_head, q);
#endif
        return Base::sldet_serial();
    }
#if AB_SLDET_SVC && defined(__CUDA_ARCH__)
    // cvSLdet evaluated by the block's service warp (sldet_service, launch.cuh): the ring buffer already lives in this group's
    // shared-memory slice (bdf_stab's group barrier has made the newest entry visible), so the request is one word
    // (order, ring head) and the answer one word (return code).  The 5.5 KB routine leaves the step loop of the sixteen
    // conditions of the block: one instruction stream evaluates it for all of them.
    __device__ int sldet_remote() const {
        volatile int* svc = &S.svc;
        if (g.lane == 0) {
            __threadfence_block();
            *svc = SVC_REQ | (q << 8) | Base::ss_head;
        }
        g.sync();   // the request is posted before any lane polls (an older answer is overwritten)
        int v;
        while (((v = *svc) & SVC_ANS) == 0)
            if (AB_SLDET_SVC_SLEEP > 0) __nanosleep(AB_SLDET_SVC_SLEEP);
        g.sync();
        return (v & 0xff) - 16;
    }
#endif
    AB_HD void gsync() const { g.sync(); }
    AB_HD bool leader() const { return g.lane == 0; }
    AB_HD bool rhs_to_zn1(double scale) { return cold_rhs_store<M, G, C>(S, ncols, tn, 0.0, scale, 1); }
    AB_HD int initial_step(double tout, double* h0, int* nfe_hin) { return cold_hin<M, G, C>(S, o, ncols, tn, tout, h0, nfe_hin); }
    AB_HD void scale_zn1(double f) {
        AB_UNROLL
        for (int c = 0; c < C; ++c) {
            if (!active(c)) continue;
            AB_UNROLL
            for (int r = 0; r < NX; ++r) S.zn[1][r][col(c)] *= f;
        }
    }
    AB_HD void init_vectors() {
        dense_lu = false;
        root_init();
        AB_UNROLL
        for (int c = 0; c < C; ++c) {
            AB_UNROLL
            for (int r = 0; r < NX; ++r) { acor[c][r] = 0.0; ewt[c][r] = 0.0; }
        }
    }
    // cvIncreaseBDF's vector part: zn[L] = A1 * zn[qmax], zn[2..q] += ll[j] * zn[L]
    // (rolled over the rows: order changes are rare, and this code sits inside the step loop's address range)
    AB_HD void order_up(const double A1, const double (&ll)[LMAX + 1]) {
        AB_UNROLL
        for (int c = 0; c < C; ++c) {
            if (!active(c)) continue;
            AB_NOUNROLL
            for (int r = 0; r < NX; ++r) {
                const double zL = A1 * S.zn[QMAX][r][col(c)];
                AB_UNROLL
                for (int j = 2; j <= QMAX; ++j) {
                    if (j == L) S.zn[j][r][col(c)] = zL;
                    else if (j <= q) S.zn[j][r][col(c)] += ll[j] * zL;
                }
            }
        }
    }
    // cvDecreaseBDF's vector part: zn[2..q-1] -= ll[j] * zn[q]
    AB_HD void order_down(const double (&ll)[LMAX + 1]) {
        AB_UNROLL
        for (int c = 0; c < C; ++c) {
            if (!active(c)) continue;
            AB_NOUNROLL
            for (int r = 0; r < NX; ++r) {
                double zq = 0.0;
                AB_UNROLL
                for (int j = 2; j <= QMAX; ++j)
                    if (j == q) zq = S.zn[j][r][col(c)];
                AB_UNROLL
                for (int j = 2; j < QMAX; ++j)
                    if (j < q) S.zn[j][r][col(c)] -= ll[j] * zq;
            }
        }
    }
    // cvCompleteStep's vector part while the history update is deferred
    AB_HD void store_correction(const bool save) {
        AB_UNROLL
        for (int c = 0; c < C; ++c) {
            if (!active(c)) continue;
            AB_UNROLL
            for (int r = 0; r < NX; ++r) {
                S.acorbuf[r][col(c)] = acor[c][r];
                if (save) S.zn[QMAX][r][col(c)] = acor[c][r];
            }
        }
    }
    AB_HD void store_correction_high() {
        AB_UNROLL
        for (int c = 0; c < C; ++c) {
            if (!active(c)) continue;
            AB_UNROLL
            for (int r = 0; r < NX; ++r) S.zn[QMAX][r][col(c)] = acor[c][r];
        }
    }
    // max-WRMS of acor - cquot * zn[qmax] (cvComputeEtaqp1, cvodes.c:7900)
    AB_HD double wrms_max_up(const double cquot) const {
        double tmp[C][NX];
        AB_UNROLL
        for (int c = 0; c < C; ++c) {
            AB_UNROLL
            for (int r = 0; r < NX; ++r) tmp[c][r] = active(c) ? acor[c][r] - cquot * S.zn[QMAX][r][col(c)] : 0.0;
        }
        return wrms_max(tmp);
    }
    // cvBDFStab: squared weighted sums of the (completed) state column of zn[q] and zn[q-1]; valid on the leader
    AB_HD void stab_sums(const double lq, const double lqm1, double& s1, double& s2) const {
        if (g.lane == 0) {
            AB_UNROLL
            for (int r = 0; r < NX; ++r) {
                const double a = (S.zn[q][r][0] + lq * acor[0][r]) * ewt[0][r];
                const double b = (S.zn[q - 1][r][0] + lqm1 * acor[0][r]) * ewt[0][r];
                s1 += a * a;
                s2 += b * b;
            }
        }
    }

    // -----------------------------------------------------------------------------------------
    // Step alignment.  The per-step code is larger than the SM's instruction cache; warps that run it at unrelated
    // phases evict each other's lines (profiles/r01_v1_ncu_summary.md: stall_no_instruction is the top stall).  The
    // warps of a block that are inside the step loop therefore start every step together: a generation barrier in
    // shared memory with dynamic membership (a warp leaves while it writes outputs or loads the next condition, and
    // re-joins at the generation of the slowest member).  sync_pos[v] = generation warp v has arrived at, INT_MAX =
    // not a member.  The member with the smallest generation never waits, so the scheme cannot deadlock.
    // -----------------------------------------------------------------------------------------
    AB_HD void step_sync_join() {
#if defined(__CUDA_ARCH__)
        if (G == 32 && sync_pos) {
            int m = 0x7fffffff;
            for (int v = 0; v < sync_n; ++v) {
                const int pv = sync_pos[v];
                if (v != sync_w && pv < m) m = pv;
            }
            if (m != 0x7fffffff) sync_gen = m;
            if (g.lane == 0) sync_pos[sync_w] = sync_gen;
            __syncwarp();
        }
#endif
    }
    AB_HD void step_sync_leave() {
#if defined(__CUDA_ARCH__)
        if (G == 32 && sync_pos) {
            if (g.lane == 0) sync_pos[sync_w] = 0x7fffffff;
            __syncwarp();
        }
        if (G == 16 && sync_pos && g.lane == 0) sync_pos[sync_w] = 0x7fffffff;   // see step_sync_arrive
#endif
    }
    AB_HD void step_sync_arrive() {
#if defined(__CUDA_ARCH__)
        if (G == 32 && sync_pos) {
            ++sync_gen;
            if (g.lane == 0) sync_pos[sync_w] = sync_gen;
            __syncwarp();
            for (;;) {
                int m = 0x7fffffff;
                for (int v = 0; v < sync_n; ++v) {
                    const int pv = sync_pos[v];
                    if (pv < m) m = pv;
                }
                if (m >= sync_gen) break;
                __nanosleep(40);
            }
        }
        // Two conditions per warp (G == 16): the warp takes part in the alignment only while BOTH of its conditions are at the
        // top of a step attempt together (__activemask() is full: the two halves execute this very instruction as one
        // stream, and they stay together through the uniform poll loop).  A half that is here alone — its partner repeats an
        // attempt, writes a timepoint, or waits at a re-convergence point for this half — must not wait for anybody (the
        // partner could be waiting for it), so it takes the warp out of the group and runs free; the warp re-joins, at the
        // generation of the slowest member, the next time both halves arrive together.  sync_pos[w] itself carries the
        // membership and the generation, so the two halves cannot disagree about them.
        if (G == 16 && sync_pos) {
            if (__activemask() != 0xffffffffu) {
                if (g.lane == 0) sync_pos[sync_w] = 0x7fffffff;
                return;
            }
            int cur = sync_pos[sync_w], m = 0x7fffffff;
            for (int v = 0; v < sync_n; ++v) {
                const int pv = sync_pos[v];
                if (v != sync_w && pv < m) m = pv;
            }
            if (cur == 0x7fffffff) cur = (m != 0x7fffffff) ? m : 0;
            const int gen = cur + 1;
            __syncwarp();
            if ((threadIdx.x & 31) == 0) sync_pos[sync_w] = gen;
            __syncwarp();
            for (;;) {
                m = 0x7fffffff;
                for (int v = 0; v < sync_n; ++v) {
                    const int pv = sync_pos[v];
                    if (pv < m) m = pv;
                }
                if (m >= gen) break;
                __nanosleep(40);
            }
        }
#endif
    }

    AB_HD int col(int c) const { return g.lane + G * c; }
    AB_HD bool active(int c) const {
        if constexpr (STG) return col(c) < ncols && col(c) >= col_lo;
        else return col(c) < ncols;
    }
    AB_HD bool has_sens() const { return ncols > 1; }

    // -----------------------------------------------------------------------------------------
    // norms
    // -----------------------------------------------------------------------------------------
    // max over active columns of WRMS(v_col, ewt_col)  (N_VWrmsNorm_SensWrapper + cvSensUpdateNorm,
    // reference sundials_nvector_senswrapper.c:426-440; errconS = TRUE in AMICI, src/solver.cpp:641-656)
    // (three partial sums per column: the dependent DFMA chain of a 9-state column is 3 + 2 long instead of 9)
    AB_HD double wrms_max(const double (&v)[C][NX]) const {
        double m = 0.0;
        AB_UNROLL
        for (int c = 0; c < C; ++c) {
            if (!active(c)) continue;
            double s3[3] = {0.0, 0.0, 0.0};
            AB_UNROLL
            for (int r = 0; r < NX; ++r) {
                const double pr = v[c][r] * ewt[c][r];
                s3[r % 3] += pr * pr;
            }
            const double s = (s3[0] + s3[1]) + s3[2];
            const double n = sqrt(s * (1.0 / NX));
            m = (n > m || n != n) ? n : m;
        }
        return g.max(m);
    }
    // same for a Nordsieck column j held in shared memory (+ lj * acor while the completion of the step is pending)
    AB_HD double wrms_max_zn(int j, const double lj = 0.0) const {
        double m = 0.0;
        AB_UNROLL
        for (int c = 0; c < C; ++c) {
            if (!active(c)) continue;
            double s3[3] = {0.0, 0.0, 0.0};
            AB_UNROLL
            for (int r = 0; r < NX; ++r) {
                const double pr = (pending ? S.zn[j][r][col(c)] + lj * acor[c][r] : S.zn[j][r][col(c)]) * ewt[c][r];
                s3[r % 3] += pr * pr;
            }
            const double s = (s3[0] + s3[1]) + s3[2];
            const double n = sqrt(s * (1.0 / NX));
            m = (n > m || n != n) ? n : m;
        }
        return g.max(m);
    }

    // cvEwtSetSS (cvodes.c:9118) / cvSensEwtSetSS (:9263)
    AB_HD void set_ewt() {
        AB_UNROLL
        for (int c = 0; c < C; ++c) {
            if (!active(c)) continue;
            const bool st = (col(c) == 0);
            const double rt = st ? o.rtol : o.rtol_fsa, at = st ? o.atol : o.atol_fsa;
            AB_UNROLL
            for (int r = 0; r < NX; ++r) {
                const double y = pending ? S.zn[0][r][col(c)] + acor[c][r] : S.zn[0][r][col(c)];   // l[0] = 1
                ewt[c][r] = fast_div(1.0, rt * fabs(y) + at);
            }
        }
    }

    // -----------------------------------------------------------------------------------------
    // linear algebra: emitted sparse LU (M::lu_factor / M::lu_solve) with a dense pivoted fallback
    // -----------------------------------------------------------------------------------------
    // b <- A^{-1} b for every own column
    AB_HD void lu_solve(double (&b)[C][NX]) const {
        if (!dense_lu) {
            AB_UNROLL
            for (int c = 0; c < C; ++c)
                if (active(c)) M::lu_solve(S.lus, b[c]);
        } else {
            AB_UNROLL
            for (int c = 0; c < C; ++c) {
                if (!active(c)) continue;
                double tmp[NX];
                AB_UNROLL
                for (int r = 0; r < NX; ++r) tmp[r] = b[c][r];
                cold_dense_solve<M, G, C>(S, tmp);
                AB_UNROLL
                for (int r = 0; r < NX; ++r) b[c][r] = tmp[r];
            }
        }
    }

    // cvLsSetup + cvLsLinSys + cvNlsLSetupSensSim (cvodes_ls.c:1611-1710, :1386-1470; cvodes_nls_sim.c:300-331)
    AB_HD int lsetup(bool jbad_nls, const Eval& e) {
        if (jbad_nls) convfail = CONVFAIL_BAD_J;
        const double dgamma = fabs(fast_div(gamma, gammap) - 1.0);
        const bool jbad = (nst == 0) || (nst >= nstlj + 51) || (convfail == CONVFAIL_BAD_J && dgamma < 0.2) ||
                          (convfail == CONVFAIL_OTHER);
        double Jv[M::NNZ_J > 0 ? M::NNZ_J : 1];
        g.sync();
        if (jbad) {
            jcur = true;
            M::jac(tn, e.y, S.p, S.k, S.h, e.spl, Jv);
            if (g.lane == 0) {
                AB_UNROLL
                for (int i = 0; i < M::NNZ_J; ++i) S.savedJ[i] = Jv[i];
            }
            nje++;
            nstlj = nst;
        } else {
            jcur = false;
            AB_UNROLL
            for (int i = 0; i < M::NNZ_J; ++i) Jv[i] = S.savedJ[i];
        }
        bool fin = true;
        AB_UNROLL
        for (int i = 0; i < M::NNZ_J; ++i) fin = fin && isfinite(Jv[i]);
        double LUr[M::NLU];
        bool okf = M::lu_factor(Jv, gamma, LUr);   // every lane redundantly: no communication
        dense_lu = false;
        g.sync();
        if (g.lane == 0) {
            AB_UNROLL
            for (int i = 0; i < M::NLU; ++i) S.lus[i] = LUr[i];
        }
        g.sync();
        if (fin && !okf) {   // a static pivot failed the diagonal-preference test: dense partial pivoting
            dense_lu = true;   // S.savedJ holds the Jacobian in either branch above
            okf = cold_dense_factor<M, G, C>(S, gamma);
        }
        nsetups++;
        gamrat = 1.0;
        gammap = gamma;
        crate = 1.0;
        Base::crateS = 1.0;
        nstlp = nst;
        if (!fin || !okf) return NLS_CONV_RECVR;  // recoverable setup failure
        return AB_SUCCESS;
    }

    // -----------------------------------------------------------------------------------------
    // Nordsieck array manipulation (per row in registers)
    // -----------------------------------------------------------------------------------------
    // The three elementwise operations on the Nordsieck history — cvCompleteStep's zn[j] += l[j]*acor (cvodes.c:7634),
    // cvRescale's zn[j] *= eta^j (:6425) and cvPredict / cvRestore's Pascal-triangle update (:6507, :7373) — in ONE pass
    // over shared memory.  CVODES runs them as separate vector sweeps at the end of one step and the beginning of the
    // next; here the completion is deferred (pending) until the next step starts, so that a history entry is loaded and
    // stored once per step instead of three times.  The pass is flat: zn[j] is NX*NCOLS contiguous doubles and an
    // elementwise update does not care which lane owns a column, so all G lanes share the elements (JAK-STAT: 162
    // elements in 6 rounds of a warp instead of 9 rows on 18 lanes).  pred = +1 predict, -1 restore, 0 neither.
    // One instantiation serves every order: rows above the current order are taken as zero and not written back (x + 0 and
    // x * f are exact, so the values are those of an order-specific pass).  Two conditions that share a warp (G = 16) at
    // different orders then run ONE instruction stream instead of two serialised instantiations, and the hot loop holds one
    // copy of the pass instead of one per order in use.
    template <int PRED>
    AB_HD void history_pass(const bool cmpl, const bool resc) {
        constexpr int NEL = NX * NCOLS;
        constexpr int Q = QMAX;
        double* z0 = &S.zn[0][0][0];
        const double* ab = &S.acorbuf[0][0];
        double f[Q + 1], lc[Q + 1];
        f[0] = 1.0;
        AB_UNROLL
        for (int jj = 1; jj <= Q; ++jj) f[jj] = f[jj - 1] * eta;
        AB_UNROLL
        for (int jj = 0; jj <= Q; ++jj) lc[jj] = (cmpl && jj <= q) ? l[jj] : 0.0;   // no pending correction: add 0 * acorbuf
        g.sync();
        AB_NOUNROLL
        for (int e = g.lane; e < NEL; e += G) {
            double z[Q + 1];
            AB_UNROLL
            for (int jj = 0; jj <= Q; ++jj) z[jj] = (jj <= 1 || jj <= q) ? z0[jj * NEL + e] : 0.0;
            const double a = cmpl ? ab[e] : 0.0;
            AB_UNROLL
            for (int jj = 0; jj <= Q; ++jj) z[jj] += lc[jj] * a;
            if (resc) {
                AB_UNROLL
                for (int jj = 1; jj <= Q; ++jj) z[jj] *= f[jj];
            }
            if (PRED != 0) {
                AB_UNROLL
                for (int kk = 1; kk <= Q; ++kk) {
                    AB_UNROLL
                    for (int jj = Q; jj >= kk; --jj) z[jj - 1] = PRED > 0 ? z[jj - 1] + z[jj] : z[jj - 1] - z[jj];
                }
            }
            AB_UNROLL
            for (int jj = 0; jj <= Q; ++jj)
                if (jj <= 1 || jj <= q) z0[jj * NEL + e] = z[jj];
        }
        g.sync();
    }
    // -----------------------------------------------------------------------------------------
    // nonlinear solve: cvNls (cvodes.c:6894) + SUNNonlinSolSolve_Newton (sunnonlinsol_newton.c:176-330)
    // + cvNlsResidualSensSim / cvNlsLSolveSensSim / cvNlsConvTestSensSim (cvodes_nls_sim.c:333-515).
    // One residual site, one solve site.
    // -----------------------------------------------------------------------------------------
    AB_HD int nls(int nflag) {
        if constexpr (STG) {
            if (o.ism == 2 && ncols > 1) {
                // staggered corrector: the states alone (cvNls with ism = CV_STAGGERED, cvodes_nls.c): the sensitivity columns sit
                // this iteration out (they take part in the group's barriers and reductions as inactive columns)
                const int ncols_all = ncols;
                ncols = 1;
                const int rc = nls_impl(nflag);
                ncols = ncols_all;
                return rc;
            }
        }
        return nls_impl(nflag);
    }
    AB_HD int nls_impl(int nflag) {
        convfail = (nflag == FIRST_CALL || nflag == PREV_ERR_FAIL) ? CONVFAIL_NONE : CONVFAIL_OTHER;
        bool callSetup = (nflag == PREV_CONV_FAIL) || (nflag == PREV_ERR_FAIL) || (nst == 0) ||
                         (nst >= nstlp + 20) || (fabs(gamrat - 1.0) > 0.3);
        AB_UNROLL
        for (int c = 0; c < C; ++c) {
            AB_UNROLL
            for (int r = 0; r < NX; ++r) acor[c][r] = 0.0;
        }
        bool jbad = false;
        int ret = AB_SUCCESS;
        Eval e;
        double delta[C][NX];
        for (;;) {          // attempts (at most two: stale Jacobian data, then fresh)
            int m = 0;
            for (;;) {      // Newton iterations
                // residual: rl1*zn1 + ycor - gamma * F(tn, zn0 + ycor)
                {
                    double ycur[C][NX];
                    AB_UNROLL
                    for (int c = 0; c < C; ++c) {
                        AB_UNROLL
                        for (int r = 0; r < NX; ++r) ycur[c][r] = active(c) ? S.zn[0][r][col(c)] + acor[c][r] : 0.0;
                    }
                    const bool ok = model_rhs<M, G, C>(S, g, ncols, tn, ycur, delta, e, m > 0);
                    nfe++;
                    if (!ok) { ret = NLS_RHS_RECVR; break; }
                }
                if (m == 0 && callSetup) {
                    ret = lsetup(jbad, e);
                    if (ret != AB_SUCCESS) break;
                }
                nni++;
                // delta <- -A^{-1} residual
                AB_UNROLL
                for (int c = 0; c < C; ++c) {
                    if (!active(c)) continue;
                    AB_UNROLL
                    for (int r = 0; r < NX; ++r)
                        delta[c][r] = -(rl1 * S.zn[1][r][col(c)] + acor[c][r] - gamma * delta[c][r]);
                }
                lu_solve(delta);
                const double sc = (gamrat != 1.0) ? fast_div(2.0, 1.0 + gamrat) : 1.0;   // cvLsSolve scaling
                AB_UNROLL
                for (int c = 0; c < C; ++c) {
                    if (!active(c)) continue;
                    AB_UNROLL
                    for (int r = 0; r < NX; ++r) {
                        if (gamrat != 1.0) delta[c][r] *= sc;
                        acor[c][r] += delta[c][r];
                    }
                }
                const double Del = wrms_max(delta);
                if (m > 0) crate = ab_max(0.3 * crate, fast_div(Del, delp));
                const double dcon = fast_div(Del * ab_min(1.0, crate), tq[4]);
                if (dcon <= 1.0) {
                    acnrm = (m == 0) ? Del : wrms_max(acor);
                    jcur = false;
                    return AB_SUCCESS;
                }
                if (!(Del == Del) || (m >= 1 && Del > 2.0 * delp)) { ret = NLS_CONV_RECVR; break; }
                delp = Del;
                m++;
                if (m >= o.max_nonlin_iters) { ret = NLS_CONV_RECVR; break; }
            }
            // recoverable failure with stale Jacobian data: retry once with a fresh J
            if (ret > 0 && !jcur) {
                callSetup = true;
                jbad = true;
                AB_UNROLL
                for (int c = 0; c < C; ++c) {
                    AB_UNROLL
                    for (int r = 0; r < NX; ++r) acor[c][r] = 0.0;
                }
                continue;
            }
            break;
        }
        return ret;
    }

    // -----------------------------------------------------------------------------------------
    // staggered corrector, sensitivity stage (BdfControl::staggered_stage)
    // -----------------------------------------------------------------------------------------
    // f(tn, y) at the converged state (cvodes.c:5944-5954); the Jacobian of the sensitivity right-hand sides is evaluated from
    // the same state inside model_rhs, so only the finiteness of f is of interest here
    AB_HD bool stg_rhs_at_converged_state() {
        const int ncols_all = ncols;
        ncols = 1;
        Eval e;
        double ycur[C][NX], f[C][NX];
        AB_UNROLL
        for (int c = 0; c < C; ++c) {
            AB_UNROLL
            for (int r = 0; r < NX; ++r) ycur[c][r] = active(c) ? S.zn[0][r][col(c)] + acor[c][r] : 0.0;
        }
        const bool ok = model_rhs<M, G, C>(S, g, ncols, tn, ycur, f, e);
        ncols = ncols_all;
        return ok;
    }
    // cvStgrNls (cvodes.c:7167-7210) + SUNNonlinSolSolve_Newton with cvNlsResidualSensStg / cvNlsLSolveSensStg /
    // cvNlsConvTestSensStg (cvodes_nls_stg.c:276-405): Newton on the sensitivity columns with the converged state held fixed.
    // No linear-solver setup on entry; one after a recoverable failure with stale Jacobian data (cvNlsLSetupSensStg).  At most
    // NLS_MAXCOR = 3 iterations (CVodeSetMaxNonlinIters only reaches the state solver).
    AB_HD int nls_stg() {
        col_lo = 1;
        AB_UNROLL
        for (int c = 0; c < C; ++c) {
            if (!active(c)) continue;
            AB_UNROLL
            for (int r = 0; r < NX; ++r) acor[c][r] = 0.0;
        }
        bool callSetup = false, jbad = false, jcur_stg = false;
        int ret = AB_SUCCESS;
        Eval e;
        double delta[C][NX];
        for (;;) {
            int m = 0;
            for (;;) {
                {
                    double ycur[C][NX];   // all columns: the state column carries the converged state
                    AB_UNROLL
                    for (int c = 0; c < C; ++c) {
                        AB_UNROLL
                        for (int r = 0; r < NX; ++r) ycur[c][r] = (col(c) < ncols) ? S.zn[0][r][col(c)] + acor[c][r] : 0.0;
                    }
                    const bool ok = model_rhs<M, G, C>(S, g, ncols, tn, ycur, delta, e, m > 0);
                    if (!ok) { ret = NLS_RHS_RECVR; break; }
                }
                if (m == 0 && callSetup) {
                    ret = lsetup(jbad, e);   // cvNlsLSetupSensStg: convfail = CV_FAIL_BAD_J, Jacobian data from the converged state
                    jcur_stg = jcur;
                    if (ret != AB_SUCCESS) break;
                }
                AB_UNROLL
                for (int c = 0; c < C; ++c) {
                    if (!active(c)) continue;
                    AB_UNROLL
                    for (int r = 0; r < NX; ++r)
                        delta[c][r] = -(rl1 * S.zn[1][r][col(c)] + acor[c][r] - gamma * delta[c][r]);
                }
                lu_solve(delta);
                const double sc = (gamrat != 1.0) ? fast_div(2.0, 1.0 + gamrat) : 1.0;
                AB_UNROLL
                for (int c = 0; c < C; ++c) {
                    if (!active(c)) continue;
                    AB_UNROLL
                    for (int r = 0; r < NX; ++r) {
                        if (gamrat != 1.0) delta[c][r] *= sc;
                        acor[c][r] += delta[c][r];
                    }
                }
                const double Del = wrms_max(delta);
                if (m > 0) Base::crateS = ab_max(0.3 * Base::crateS, fast_div(Del, delp));
                const double dcon = fast_div(Del * ab_min(1.0, Base::crateS), tq[4]);
                if (dcon <= 1.0) {
                    Base::acnrmS = (m == 0) ? Del : wrms_max(acor);
                    jcur = false;
                    col_lo = 0;
                    return AB_SUCCESS;
                }
                if (!(Del == Del) || (m >= 1 && Del > 2.0 * delp)) { ret = NLS_CONV_RECVR; break; }
                delp = Del;
                m++;
                if (m >= 3) { ret = NLS_CONV_RECVR; break; }
            }
            if (ret > 0 && !jcur_stg) {
                callSetup = true;
                jbad = true;
                AB_UNROLL
                for (int c = 0; c < C; ++c) {
                    if (!active(c)) continue;
                    AB_UNROLL
                    for (int r = 0; r < NX; ++r) acor[c][r] = 0.0;
                }
                continue;
            }
            break;
        }
        col_lo = 0;
        return ret;
    }

    // -----------------------------------------------------------------------------------------
    // root finding on the state column (every lane evaluates g redundantly from column 0 in shared memory)
    // -----------------------------------------------------------------------------------------
    AB_HD void root_init() {
        AB_UNROLL
        for (int i = 0; i < NES; ++i) { glo[i] = ghi[i] = grout[i] = 0.0; gactive[i] = true; iroots[i] = 0; }
        tlo = thi = trout = ttol = 0.0;
        irfnd = 0;
        nge = 0;
    }
    // CVodeGetDky(t, 0) for the state column: y(t) = sum_j zn[j] * s^j, s = (t - tn) / h
    AB_HD void state_at(double t, double (&y)[NX]) const {
        const double s = (t - tn) / h;
        AB_UNROLL
        for (int r = 0; r < NX; ++r) {
            double v = 0.0;
            AB_UNROLL
            for (int j = QMAX; j >= 0; --j)
                if (j <= q) v = v * s + S.zn[j][r][0];
            y[r] = v;
        }
    }
    // zn[0] + a * zn[1] for the state column
    AB_HD void state_lin(double a, double (&y)[NX]) const {
        AB_UNROLL
        for (int r = 0; r < NX; ++r) y[r] = S.zn[0][r][0] + (a != 0.0 ? a * S.zn[1][r][0] : 0.0);
    }
    AB_HD void eval_g(double t, const double (&y)[NX], double (&gv)[NES]) {
        double spl[M::NSPL > 0 ? M::NSPL : 1], rv[M::NE > 0 ? M::NE : 1];
        AB_UNROLL
        for (int is = 0; is < M::NSPL; ++is) {
            const double v = spline_eval_scaled<M, NCOLS>(S.spl, is, 0, t, true);
            spl[is] = spline_is_log<M>(is) ? exp(v) : v;
        }
        double yp[NX];
        AB_UNROLL
        for (int r = 0; r < NX; ++r) yp[r] = y[r];
        x_pos<M>(S, yp);
        M::root(t, yp, S.p, S.k, S.h, spl, rv);
        AB_UNROLL
        for (int i = 0; i < NES; ++i) gv[i] = rv[i];
        nge++;
    }
    // Solution at the root: overwrite zn[0] of every column by y(trout) and move tn there.  AMICI re-initialises the
    // solver after every root return (reference src/forwardproblem.cpp:389-514), so the history is not needed again.
    AB_HD void collapse_to(double t) {
        const double s = (t - tn) / h;
        g.sync();
        AB_UNROLL
        for (int c = 0; c < C; ++c) {
            if (!active(c)) continue;
            AB_NOUNROLL
            for (int r = 0; r < NX; ++r) {
                double v = 0.0;
                AB_UNROLL
                for (int j = QMAX; j >= 0; --j)
                    if (j <= q) v = v * s + S.zn[j][r][col(c)];
                S.zn[0][r][col(c)] = v;
            }
        }
        g.sync();
        tn = t;
    }
    // cvRcheck1 (cvodes.c:8576)
    AB_HD void rcheck1() {
        double y[NX];
        AB_UNROLL
        for (int i = 0; i < NES; ++i) iroots[i] = 0;
        tlo = tn;
        ttol = (fabs(tn) + fabs(h)) * DBL_EPSILON * 100.0;
        g.sync();
        state_lin(0.0, y);
        eval_g(tlo, y, glo);
        nge = 1;
        bool zroot = false;
        AB_UNROLL
        for (int i = 0; i < NES; ++i)
            if (fabs(glo[i]) == 0.0) { zroot = true; gactive[i] = false; }
        if (!zroot) return;
        const double hratio = ab_max(ttol / fabs(h), 0.1);
        const double tplus = tlo + hratio * h;
        state_lin(hratio, y);
        eval_g(tplus, y, ghi);
        AB_UNROLL
        for (int i = 0; i < NES; ++i)
            if (!gactive[i] && fabs(ghi[i]) != 0.0) { gactive[i] = true; glo[i] = ghi[i]; }
    }
    // cvRcheck2 (cvodes.c:8648): 0, 1 = RTFOUND, 3 = CLOSERT
    AB_HD int rcheck2() {
        if (irfnd == 0) return 0;
        double y[NX];
        g.sync();
        state_at(tlo, y);
        eval_g(tlo, y, glo);
        bool zroot = false;
        AB_UNROLL
        for (int i = 0; i < NES; ++i) iroots[i] = 0;
        AB_UNROLL
        for (int i = 0; i < NES; ++i)
            if (gactive[i] && fabs(glo[i]) == 0.0) { zroot = true; iroots[i] = 1; }
        if (!zroot) return 0;
        ttol = (fabs(tn) + fabs(h)) * DBL_EPSILON * 100.0;
        const double smallh = (h > 0.0) ? ttol : -ttol;
        const double tplus = tlo + smallh;
        if ((tplus - tn) * h >= 0.0) {
            const double hratio = smallh / h;
            AB_UNROLL
            for (int r = 0; r < NX; ++r) y[r] += hratio * S.zn[1][r][0];
        } else {
            state_at(tplus, y);
        }
        eval_g(tplus, y, ghi);
        zroot = false;
        bool closert = false;
        AB_UNROLL
        for (int i = 0; i < NES; ++i) {
            if (!gactive[i]) continue;
            if (fabs(ghi[i]) == 0.0) {
                if (iroots[i] == 1) closert = true;
                zroot = true;
                iroots[i] = 1;
            } else if (iroots[i] == 1) {
                glo[i] = ghi[i];
            }
        }
        if (closert) return 3;
        return zroot ? 1 : 0;
    }
    // sign-change scan shared by cvRootfind's entry and loop (cvodes.c:8871-8900, :8990-9020)
    AB_HD void scan_signs(const double (&gv)[NES], bool& zroot, bool& sgnchg, int& imax) const {
        double maxfrac = 0.0;
        zroot = false;
        sgnchg = false;
        AB_UNROLL
        for (int i = 0; i < NES; ++i) {
            if (!gactive[i]) continue;
            if (fabs(gv[i]) == 0.0) {
                zroot = true;   // rootdir = 0: both directions
            } else if ((glo[i] < 0.0 && gv[i] > 0.0) || (glo[i] > 0.0 && gv[i] < 0.0)) {
                const double gfrac = fabs(gv[i] / (gv[i] - glo[i]));
                if (gfrac > maxfrac) { sgnchg = true; maxfrac = gfrac; imax = i; }
            }
        }
    }
    // cvRootfind (cvodes.c:8859): Illinois iteration on (tlo, thi]; returns true if a root was found
    AB_HD bool rootfind() {
        int imax = 0;
        bool zroot, sgnchg;
        scan_signs(ghi, zroot, sgnchg, imax);
        if (!sgnchg) {
            trout = thi;
            AB_UNROLL
            for (int i = 0; i < NES; ++i) grout[i] = ghi[i];
            if (!zroot) return false;
            AB_UNROLL
            for (int i = 0; i < NES; ++i) {
                iroots[i] = 0;
                if (gactive[i] && fabs(ghi[i]) == 0.0) iroots[i] = glo[i] > 0.0 ? -1 : 1;
            }
            return true;
        }
        double alph = 1.0;
        int side = 0, sideprev = -1;
        for (;;) {
            if (fabs(thi - tlo) <= ttol) break;
            if (sideprev == side) alph = (side == 2) ? alph * 2.0 : alph * 0.5;
            else alph = 1.0;
            const double ghim = pick(ghi, imax), glom = pick(glo, imax);
            double tmid = thi - (thi - tlo) * ghim / (ghim - alph * glom);
            if (fabs(tmid - tlo) < 0.5 * ttol) {
                const double fracint = fabs(thi - tlo) / ttol;
                const double fracsub = (fracint > 5.0) ? 0.1 : 0.5 / fracint;
                tmid = tlo + fracsub * (thi - tlo);
            }
            if (fabs(thi - tmid) < 0.5 * ttol) {
                const double fracint = fabs(thi - tlo) / ttol;
                const double fracsub = (fracint > 5.0) ? 0.1 : 0.5 / fracint;
                tmid = thi - fracsub * (thi - tlo);
            }
            double y[NX];
            state_at(tmid, y);
            eval_g(tmid, y, grout);
            sideprev = side;
            scan_signs(grout, zroot, sgnchg, imax);
            if (sgnchg) {
                thi = tmid;
                AB_UNROLL
                for (int i = 0; i < NES; ++i) ghi[i] = grout[i];
                side = 1;
                if (fabs(thi - tlo) <= ttol) break;
                continue;
            }
            if (zroot) {
                thi = tmid;
                AB_UNROLL
                for (int i = 0; i < NES; ++i) ghi[i] = grout[i];
                break;
            }
            tlo = tmid;
            AB_UNROLL
            for (int i = 0; i < NES; ++i) glo[i] = grout[i];
            side = 2;
            if (fabs(thi - tlo) <= ttol) break;
        }
        trout = thi;
        AB_UNROLL
        for (int i = 0; i < NES; ++i) {
            grout[i] = ghi[i];
            iroots[i] = 0;
            if (!gactive[i]) continue;
            if (fabs(ghi[i]) == 0.0) iroots[i] = glo[i] > 0.0 ? -1 : 1;
            if ((glo[i] < 0.0 && ghi[i] > 0.0) || (glo[i] > 0.0 && ghi[i] < 0.0)) iroots[i] = glo[i] > 0.0 ? -1 : 1;
        }
        return true;
    }
    // cvRcheck3 (cvodes.c:8725): look for a root in (tlo, min(tn, tout)]; true = found (trout, iroots set)
    AB_HD bool rcheck3(double toutc, bool one_step) {
        double y[NX];
        g.sync();
        if (one_step || (toutc - tn) * h >= 0.0) {
            thi = tn;
            state_lin(0.0, y);
        } else {
            thi = toutc;
            state_at(thi, y);
        }
        eval_g(thi, y, ghi);
        ttol = (fabs(tn) + fabs(h)) * DBL_EPSILON * 100.0;
        const bool found = rootfind();
        AB_UNROLL
        for (int i = 0; i < NES; ++i)
            if (!gactive[i] && grout[i] != 0.0) gactive[i] = true;
        tlo = trout;
        AB_UNROLL
        for (int i = 0; i < NES; ++i) glo[i] = grout[i];
        return found;
    }

};

}  // namespace amici_b200
#endif
