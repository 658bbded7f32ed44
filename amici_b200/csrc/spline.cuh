// Device Hermite spline evaluator (values and parameter sensitivities).
//
// Restates the part of the reference's HermiteSpline (reference src/splinefunctions.cpp) that its
// shipped models use: non-equidistant nodes, node slopes by averaged one-sided finite differences
// (:172-195), zero-derivative boundary conditions (:217-218, :263-264), constant extrapolation
// (:865-866, :900-901), optional logarithmic parametrisation (:59-74, :80-87, :295-302, :431-444),
// Horner coefficients (d, c, b, a) per interval (:304-340, sensitivities :736-760), exact node
// values at nodes (:934-935).
//
// Layout: one table per trajectory in the group's shared-memory slice, column fastest:
//   coef[ispl][interval][4][col], nodeval[ispl][node][col]
// col 0 holds the (scaled) value polynomial, col c >= 1 the polynomial of d(scaled value)/d p_plist[c-1],
// so a lane reads its own column's coefficients conflict-free and column 0 by broadcast.
//
// Models whose splines use any OTHER configuration of the reference's HermiteSpline (boundary conditions, extrapolation modes,
// given slopes, equidistant nodes; the emitter sets M::SPL_GENERAL) take the general path below, built on
// hermite_spline.cuh; the path above stays as it is for the default configuration, so the shipped models' code is unchanged.
#ifndef AMICI_B200_SPLINE_CUH
#define AMICI_B200_SPLINE_CUH
#include "hermite_spline.cuh"
#include "model_common.cuh"

namespace amici_b200 {

template <class M, int NCOLS>
struct SplineTable {
    static constexpr int NS = M::NSPL > 0 ? M::NSPL : 1;
    static constexpr int NN = M::SPL_NN > 1 ? M::SPL_NN : 2;
    static constexpr bool GEN = M::SPL_GENERAL;
    double coef[NS][NN - 1][4][NCOLS];
    double nodeval[NS][NN][NCOLS];
    // general configurations only: extrapolation lines (b_first, a_first, b_last, a_last) and the value at t = inf
    double ext[GEN ? NS : 1][GEN ? 4 : 1][GEN ? NCOLS : 1];
    double fin[GEN ? NS : 1][GEN ? NCOLS : 1];
};

template <class M>
AB_HD constexpr bool spline_is_log(const int ispl) {
    if constexpr (M::SPL_GENERAL) return M::spline_cfg(ispl).logarithmic;
    else return M::SPL_LOG;
}

// General configuration: column `col` of spline ispl from hermite_spline.cuh's pieces (same arithmetic as the reference class;
// tests/known_answers.py::check_spline_cases pins those pieces against it bit for bit).
template <class M, int NCOLS>
AB_HD void spline_build_column_general(SplineTable<M, NCOLS>& T, const int ispl, const int col, const int ip, const double* p, const double* k) {
    constexpr int NN = SplineTable<M, NCOLS>::NN;
    const SplineCfg c = M::spline_cfg(ispl);
    const int nn = M::spline_n_nodes(ispl);
    double x[NN], v[NN], m[NN], P[NN], S[NN], ext[4];
    AB_UNROLL
    for (int i = 0; i < NN; ++i) { x[i] = i < nn ? M::spline_node(ispl, i) : 0.0; m[i] = 0.0; v[i] = 1.0; }
    M::spline_values(ispl, p, k, v);
    if (!c.fd) M::spline_slopes(ispl, p, k, m);
    hs_slopes<NN>(c, nn, x, v, m);
    if (ip < 0) {
        hs_scaled<NN>(c, nn, v, m, P, S);
        hs_extrapolation<NN>(c, nn, x, P, S, ext);
    } else {
        double dv[NN], dm[NN];
        AB_UNROLL
        for (int i = 0; i < NN; ++i) dv[i] = dm[i] = 0.0;
        M::spline_dvaluesdp(ispl, ip, p, k, dv);
        if (!c.fd) M::spline_dslopesdp(ispl, ip, p, k, dm);
        hs_slopes_sensi<NN>(c, nn, x, dv, dm);
        hs_scaled_sensi<NN>(c, nn, v, m, dv, dm, P, S);
        hs_extrapolation_sensi<NN>(c, nn, x, P, S, ext);
    }
    AB_UNROLL
    for (int i = 0; i < NN - 1; ++i) {
        if (i >= nn - 1) continue;
        double co[4];
        hs_interval(ip >= 0, x[i + 1] - x[i], P[i], S[i], P[i + 1], S[i + 1], co);
        AB_UNROLL
        for (int j = 0; j < 4; ++j) T.coef[ispl][i][j][col] = co[j];
    }
    AB_UNROLL
    for (int i = 0; i < NN; ++i) T.nodeval[ispl][i][col] = P[i];
    AB_UNROLL
    for (int j = 0; j < 4; ++j) T.ext[ispl][j][col] = ext[j];
    // value at t = inf (compute_final_value / compute_final_sensitivity, reference src/splinefunctions.cpp:762-844)
    const int last = nn - 1;
    double fv;
    if (ip < 0) {
        double q[4] = {0.0, 0.0, 0.0, 0.0};
        AB_UNROLL
        for (int i = 0; i < NN - 1; ++i)
            if (i == last - 1) hs_interval(false, x[i + 1] - x[i], P[i], S[i], P[i + 1], S[i + 1], q);
        if (c.ep_last == SPL_EP_CONSTANT) fv = ext[2];
        else if (c.ep_last == SPL_EP_LINEAR) fv = (c.bc_last == SPL_BC_ZERO_DERIVATIVE || ext[3] == 0) ? ext[2] : (ext[3] < 0 ? -INFINITY : INFINITY);
        else if (c.ep_last == SPL_EP_POLYNOMIAL)
            fv = q[3] != 0 ? (q[3] < 0 ? -INFINITY : INFINITY) : q[2] != 0 ? (q[2] < 0 ? -INFINITY : INFINITY) : q[1] != 0 ? (q[1] < 0 ? -INFINITY : INFINITY) : q[0];
        else {
            fv = P[0];
            AB_UNROLL
            for (int i = 0; i < NN; ++i)
                if (i < nn && (P[i] != P[0] || S[i] != 0)) fv = NAN;
        }
    } else {
        double Sv_last = 0.0;   // scaled slope of the VALUE polynomial at the last node
        AB_UNROLL
        for (int i = 0; i < NN; ++i)
            if (i == last) Sv_last = c.logarithmic ? m[i] / v[i] : m[i];
        if (c.ep_last == SPL_EP_CONSTANT || (c.bc_last == SPL_BC_ZERO_DERIVATIVE && c.ep_last == SPL_EP_LINEAR)) fv = ext[2];
        else if (c.ep_last == SPL_EP_LINEAR) fv = (Sv_last == 0) ? NAN : 0.0;
        else fv = NAN;
    }
    T.fin[ispl][col] = fv;
}

// Scaled value (col 0) / scaled sensitivity (col >= 1) at time t for a general configuration (get_value_scaled /
// get_sensitivity_scaled, reference src/splinefunctions.cpp:846-1049).  An evaluation the reference refuses (no extrapolation)
// gives NaN, which the engine's finiteness check of the right-hand side turns into a failed step.
template <class M, int NCOLS>
AB_HD double spline_eval_general(const SplineTable<M, NCOLS>& T, const int ispl, const int col, const double t) {
    constexpr int NN = SplineTable<M, NCOLS>::NN;
    const SplineCfg c = M::spline_cfg(ispl);
    const int nn = M::spline_n_nodes(ispl), last = nn - 1;
    const double x0 = M::spline_node(ispl, 0), xl = M::spline_node(ispl, last);
    if (isinf(t)) return T.fin[ispl][col];
    double tw = t;
    bool wrapped = false;
    if (t > xl) {
        if (c.ep_last == SPL_EP_NONE) return NAN;
        if (c.ep_last == SPL_EP_CONSTANT) return T.ext[ispl][2][col];
        if (c.ep_last == SPL_EP_LINEAR) return T.ext[ispl][2][col] + t * T.ext[ispl][3][col];
        if (c.ep_last == SPL_EP_PERIODIC) { tw = x0 + fmod(t - x0, xl - x0); wrapped = true; }
    } else if (t < x0) {
        if (c.ep_first == SPL_EP_NONE) return NAN;
        if (c.ep_first == SPL_EP_CONSTANT) return T.ext[ispl][0][col];
        if (c.ep_first == SPL_EP_LINEAR) return T.ext[ispl][0][col] + t * T.ext[ispl][1][col];
        if (c.ep_first == SPL_EP_PERIODIC) { tw = xl + fmod(t - x0, xl - x0); wrapped = true; }
    }
    int i;
    double len, ti;
    bool at_node = false;
    if (tw > xl) {          // polynomial extrapolation: the last interval's polynomial
        i = last - 1;
    } else if (tw < x0) {   // ... the first interval's
        i = 0;
    } else if (c.equidistant) {
        i = (int)trunc((tw - x0) / (M::spline_node(ispl, 1) - x0));
        i = i < nn - 2 ? i : nn - 2;
    } else {
        i = 0;
        AB_UNROLL
        for (int j = 1; j < NN - 1; ++j)
            if (j < nn - 1 && M::spline_node(ispl, j) < tw) i = j;
    }
    ti = M::spline_node(ispl, 0);
    len = M::spline_node(ispl, 1) - M::spline_node(ispl, 0);
    AB_UNROLL
    for (int j = 1; j < NN - 1; ++j)
        if (j == i) { ti = M::spline_node(ispl, j); len = M::spline_node(ispl, j + 1) - M::spline_node(ispl, j); }
    if (!c.equidistant && tw >= x0 && tw <= xl) at_node = (tw == M::spline_node(ispl, i + 1));
    double r;
    if (col == 0 && at_node) {
        r = T.nodeval[ispl][i + 1][col];
    } else {
        const double s = (tw - ti) / len;
        r = T.coef[ispl][i][0][col] + s * (T.coef[ispl][i][1][col] + s * (T.coef[ispl][i][2][col] + s * T.coef[ispl][i][3][col]));
    }
    if (wrapped && c.logarithmic) {
        // reference quirk: periodic extrapolation returns the UNSCALED value / sensitivity of the wrapped time from the scaled function
        if (col == 0) return exp(r);
        double v0;
        const double s = (tw - ti) / len;
        const bool node0 = at_node;
        v0 = node0 ? T.nodeval[ispl][i + 1][0]
                   : T.coef[ispl][i][0][0] + s * (T.coef[ispl][i][1][0] + s * (T.coef[ispl][i][2][0] + s * T.coef[ispl][i][3][0]));
        return r * exp(v0);
    }
    return r;
}

// Build column `col` (ip < 0: value column; else sensitivity w.r.t. parameter ip) of spline ispl.
template <class M, int NCOLS>
AB_HD void spline_build_column(SplineTable<M, NCOLS>& T, const int ispl, const int col, const int ip,
                               const double* p, const double* k) {
    if constexpr (M::SPL_GENERAL) {
        spline_build_column_general<M, NCOLS>(T, ispl, col, ip, p, k);
        return;
    }
    constexpr int NN = SplineTable<M, NCOLS>::NN;
    double v[NN], m[NN], P[NN], S[NN];
    M::spline_values(ispl, p, k, v);
    m[0] = 0.0;
    m[NN - 1] = 0.0;
    AB_UNROLL
    for (int i = 1; i < NN - 1; ++i) {
        const double dl = (v[i] - v[i - 1]) / (M::spline_node(ispl, i) - M::spline_node(ispl, i - 1));
        const double dr = (v[i + 1] - v[i]) / (M::spline_node(ispl, i + 1) - M::spline_node(ispl, i));
        m[i] = (dl + dr) / 2;
    }
    if (ip < 0) {
        AB_UNROLL
        for (int i = 0; i < NN; ++i) {
            P[i] = M::SPL_LOG ? log(v[i]) : v[i];
            S[i] = M::SPL_LOG ? m[i] / v[i] : m[i];
        }
    } else {
        double dv[NN], dm[NN];
        M::spline_dvaluesdp(ispl, ip, p, k, dv);
        dm[0] = 0.0;
        dm[NN - 1] = 0.0;
        AB_UNROLL
        for (int i = 1; i < NN - 1; ++i) {
            const double dl = (dv[i] - dv[i - 1]) / (M::spline_node(ispl, i) - M::spline_node(ispl, i - 1));
            const double dr = (dv[i + 1] - dv[i]) / (M::spline_node(ispl, i + 1) - M::spline_node(ispl, i));
            dm[i] = (dl + dr) / 2;
        }
        AB_UNROLL
        for (int i = 0; i < NN; ++i) {
            P[i] = M::SPL_LOG ? dv[i] / v[i] : dv[i];
            S[i] = M::SPL_LOG ? (dm[i] - dv[i] * m[i] / v[i]) / v[i] : dm[i];
        }
    }
    AB_UNROLL
    for (int i = 0; i < NN - 1; ++i) {
        const double len = M::spline_node(ispl, i + 1) - M::spline_node(ispl, i);
        T.coef[ispl][i][0][col] = P[i];
        T.coef[ispl][i][1][col] = len * S[i];
        if (ip < 0) {
            T.coef[ispl][i][2][col] = -3 * P[i] - 2 * len * S[i] + 3 * P[i + 1] - len * S[i + 1];
            T.coef[ispl][i][3][col] = 2 * P[i] + len * S[i] - 2 * P[i + 1] + len * S[i + 1];
        } else {
            T.coef[ispl][i][2][col] = 3 * (P[i + 1] - P[i]) - len * (2 * S[i] + S[i + 1]);
            T.coef[ispl][i][3][col] = 2 * (P[i] - P[i + 1]) + len * (S[i] + S[i + 1]);
        }
    }
    AB_UNROLL
    for (int i = 0; i < NN; ++i) T.nodeval[ispl][i][col] = P[i];
}

// Position of t relative to the nodes of spline ispl: computed once per evaluation time and shared by the value
// column and the lane's sensitivity column (the reference locates the interval separately in get_value_scaled and
// get_sensitivity_scaled, src/splinefunctions.cpp:846-1049).
struct SplineLoc {
    double t;      // the evaluation time itself (general configurations evaluate from it)
    int region;    // -1: before the first node, +1: after the last node (constant extrapolation), 0: inside
    int i_node;    // interval index
    double s;      // (t - t_i) / (t_{i+1} - t_i)
    bool at_node;  // t equals the right node of the interval exactly
};

template <class M>
AB_HD SplineLoc spline_locate(const int ispl, const double t) {
    constexpr int NN = M::SPL_NN > 1 ? M::SPL_NN : 2;
    SplineLoc L;
    L.t = t;
    if constexpr (M::SPL_GENERAL) {
        L.region = 0; L.i_node = 0; L.s = 0.0; L.at_node = false;
        return L;
    }
    L.region = (t > M::spline_node(ispl, NN - 1)) ? 1 : (t < M::spline_node(ispl, 0)) ? -1 : 0;
    int i_node = 0;
    AB_UNROLL
    for (int i = 1; i < NN - 1; ++i)
        if (M::spline_node(ispl, i) < t) i_node = i;
    double t0 = M::spline_node(ispl, 0), t1 = M::spline_node(ispl, 1);
    double inv = 1.0 / (M::spline_node(ispl, 1) - M::spline_node(ispl, 0));   // node distances are compile-time constants
    AB_UNROLL
    for (int i = 1; i < NN - 1; ++i)
        if (i_node == i) {
            t0 = M::spline_node(ispl, i);
            t1 = M::spline_node(ispl, i + 1);
            inv = 1.0 / (M::spline_node(ispl, i + 1) - M::spline_node(ispl, i));
        }
    L.i_node = i_node;
    L.at_node = (t == t1);
    L.s = (t - t0) * inv;
    return L;
}

// Scaled polynomial value of column `col` at the located time. `exact_nodes`: return the stored node value at
// a node (the reference does this for values only, not for sensitivities).
template <class M, int NCOLS>
AB_HD double spline_eval_at(const SplineTable<M, NCOLS>& T, const int ispl, const int col, const SplineLoc& L,
                            const bool exact_nodes) {
    constexpr int NN = SplineTable<M, NCOLS>::NN;
    if constexpr (M::SPL_GENERAL) return spline_eval_general<M, NCOLS>(T, ispl, col, L.t);
    if (L.region > 0) return T.nodeval[ispl][NN - 1][col];
    if (L.region < 0) return T.nodeval[ispl][0][col];
    if (exact_nodes && L.at_node) return T.nodeval[ispl][L.i_node + 1][col];
    const double d = T.coef[ispl][L.i_node][0][col], c = T.coef[ispl][L.i_node][1][col];
    const double b = T.coef[ispl][L.i_node][2][col], a = T.coef[ispl][L.i_node][3][col];
    return d + L.s * (c + L.s * (b + L.s * a));
}

template <class M, int NCOLS>
AB_HD double spline_eval_scaled(const SplineTable<M, NCOLS>& T, const int ispl, const int col, const double t,
                                const bool exact_nodes) {
    return spline_eval_at<M, NCOLS>(T, ispl, col, spline_locate<M>(ispl, t), exact_nodes);
}

}  // namespace amici_b200
#endif
