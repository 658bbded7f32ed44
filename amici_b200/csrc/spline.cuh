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
#ifndef AMICI_B200_SPLINE_CUH
#define AMICI_B200_SPLINE_CUH
#include "model_common.cuh"

namespace amici_b200 {

template <class M, int NCOLS>
struct SplineTable {
    static constexpr int NS = M::NSPL > 0 ? M::NSPL : 1;
    static constexpr int NN = M::SPL_NN > 1 ? M::SPL_NN : 2;
    double coef[NS][NN - 1][4][NCOLS];
    double nodeval[NS][NN][NCOLS];
};

// Build column `col` (ip < 0: value column; else sensitivity w.r.t. parameter ip) of spline ispl.
template <class M, int NCOLS>
AB_HD void spline_build_column(SplineTable<M, NCOLS>& T, const int ispl, const int col, const int ip,
                               const double* p, const double* k) {
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
    int region;    // -1: before the first node, +1: after the last node (constant extrapolation), 0: inside
    int i_node;    // interval index
    double s;      // (t - t_i) / (t_{i+1} - t_i)
    bool at_node;  // t equals the right node of the interval exactly
};

template <class M>
AB_HD SplineLoc spline_locate(const int ispl, const double t) {
    constexpr int NN = M::SPL_NN > 1 ? M::SPL_NN : 2;
    SplineLoc L;
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
