// General cubic Hermite spline (values and parameter sensitivities), host + device.
//
// Behavioural specification: the reference's HermiteSpline (reference src/splinefunctions.cpp; include/amici/splinefunctions.h),
// complete: node slopes given or by finite differences (:172-195), equidistant or explicit nodes (:27-56), boundary conditions
// given / zeroDerivative / natural / periodic (:197-287), extrapolation none / constant / linear / polynomial / periodic
// (:342-404, :846-944), logarithmic parametrisation (:59-74, :289-302, :431-444), sensitivity polynomials (:406-490, :736-760,
// slope sensitivities by finite differences :492-580, extrapolation :582-734), values at t = inf (:762-844), and the
// evaluation rules incl. "exact at nodes for values only" and the interval search (:846-1049).  Not a translation: no
// classes, no std::vector; plain arrays and free functions that the engine instantiates per lane-group column
// (spline.cuh) and the self-test kernel instantiates with a run-time configuration (tu/selftest.cu).
//
// Two quirks of the reference are reproduced on purpose (parity): periodic extrapolation evaluates the UNSCALED value inside the
// scaled function (:880-882, :909-913 call get_value / get_sensitivity), so a logarithmic periodic spline is exponentiated twice
// outside its node range; the `natural` boundary condition has no slope sensitivities (:507-511 throw).
#ifndef AMICI_B200_HERMITE_SPLINE_CUH
#define AMICI_B200_HERMITE_SPLINE_CUH
#include "model_common.cuh"

namespace amici_b200 {

// numeric values of the reference's enums (include/amici/defines.h:249-264)
enum : int { SPL_BC_GIVEN = -1, SPL_BC_ZERO_DERIVATIVE = 0, SPL_BC_NATURAL = 1, SPL_BC_NATURAL_ZERO_DERIVATIVE = 2, SPL_BC_PERIODIC = 3 };
enum : int { SPL_EP_NONE = -1, SPL_EP_CONSTANT = 0, SPL_EP_LINEAR = 1, SPL_EP_POLYNOMIAL = 2, SPL_EP_PERIODIC = 3 };

struct SplineCfg {
    int bc_first, bc_last, ep_first, ep_last;
    bool fd;            // node slopes by finite differences
    bool equidistant;   // nodes = first .. last in equal steps
    bool logarithmic;   // spline of log(values); value = exp(spline)
};

// the configuration of the reference's shipped models (models/model_jakstat_adjoint_py/create_splines.cpp)
constexpr SplineCfg spline_cfg_default(bool logarithmic) {
    return SplineCfg{SPL_BC_ZERO_DERIVATIVE, SPL_BC_ZERO_DERIVATIVE, SPL_EP_CONSTANT, SPL_EP_CONSTANT, true, false, logarithmic};
}

// a configuration the reference rejects at construction / sensitivity set-up (AmiException there)
AB_HD bool spline_cfg_valid(const SplineCfg& c, bool with_sensitivities) {
    if ((c.bc_first == SPL_BC_PERIODIC) != (c.bc_last == SPL_BC_PERIODIC)) return false;
    if (c.bc_first == SPL_BC_NATURAL_ZERO_DERIVATIVE || c.bc_last == SPL_BC_NATURAL_ZERO_DERIVATIVE) return false;
    if (with_sensitivities && c.fd && (c.bc_first == SPL_BC_NATURAL || c.bc_last == SPL_BC_NATURAL)) return false;
    if (with_sensitivities && c.ep_first == SPL_EP_LINEAR && c.bc_first != SPL_BC_GIVEN && c.bc_first != SPL_BC_ZERO_DERIVATIVE) return false;
    if (with_sensitivities && c.ep_last == SPL_EP_LINEAR && c.bc_last != SPL_BC_GIVEN && c.bc_last != SPL_BC_ZERO_DERIVATIVE) return false;
    return true;
}

// Node slopes m[] of the values v[] on nodes x[] (handle_inner_derivatives + handle_boundary_conditions).  m holds the given
// slopes on entry when !c.fd.
template <int NN>
AB_HD void hs_slopes(const SplineCfg& c, const int nn, const double (&x)[NN], const double (&v)[NN], double (&m)[NN]) {
    const int last = nn - 1;
    if (c.fd) {
        for (int i = 0; i < nn; ++i) m[i] = 0.0;
        if (c.equidistant) {
            const double hx2 = 2 * (x[1] - x[0]);
            for (int i = 1; i < last; ++i) m[i] = (v[i + 1] - v[i - 1]) / hx2;
        } else {
            for (int i = 1; i < last; ++i) {
                const double dl = (v[i] - v[i - 1]) / (x[i] - x[i - 1]);
                const double dr = (v[i + 1] - v[i]) / (x[i + 1] - x[i]);
                m[i] = (dl + dr) / 2;
            }
        }
    }
    if (c.bc_first == SPL_BC_GIVEN) {
        if (c.fd) m[0] = (v[1] - v[0]) / (x[1] - x[0]);
    } else if (c.bc_first == SPL_BC_ZERO_DERIVATIVE) {
        m[0] = 0;
    } else if (c.bc_first == SPL_BC_NATURAL) {
        m[0] = -0.5 * m[1] + 1.5 * (v[1] - v[0]) / (x[1] - x[0]);
    } else if (c.bc_first == SPL_BC_PERIODIC && c.fd) {
        if (c.equidistant) {
            m[0] = (v[1] - v[last - 1]) / (2 * (x[1] - x[0]));
        } else {
            const double dl = (v[last] - v[last - 1]) / (x[last] - x[last - 1]);
            const double dr = (v[1] - v[0]) / (x[1] - x[0]);
            m[0] = (dl + dr) / 2;
        }
    }
    if (c.bc_last == SPL_BC_GIVEN) {
        if (c.fd) m[last] = (v[last] - v[last - 1]) / (x[last] - x[last - 1]);
    } else if (c.bc_last == SPL_BC_ZERO_DERIVATIVE) {
        m[last] = 0;
    } else if (c.bc_last == SPL_BC_NATURAL) {
        m[last] = -0.5 * m[last - 1] + 1.5 * (v[last] - v[last - 1]) / (x[last] - x[last - 1]);
    } else if (c.bc_last == SPL_BC_PERIODIC && c.fd) {
        m[last] = m[0];
    }
}

// Slope sensitivities dm[] for one parameter from the value sensitivities dv[] (compute_slope_sensitivities_by_fd and the
// zero-derivative override of compute_coefficients_sensi).  dm holds the given slope sensitivities on entry when !c.fd.
template <int NN>
AB_HD void hs_slopes_sensi(const SplineCfg& c, const int nn, const double (&x)[NN], const double (&dv)[NN], double (&dm)[NN]) {
    const int last = nn - 1;
    if (c.fd) {
        if (c.bc_first == SPL_BC_GIVEN) {
            dm[0] = (dv[1] - dv[0]) / (x[1] - x[0]);
        } else if (c.bc_first == SPL_BC_ZERO_DERIVATIVE) {
            dm[0] = 0;
        } else if (c.bc_first == SPL_BC_PERIODIC) {
            if (c.equidistant) {
                dm[0] = (dv[1] - dv[last - 1]) / (2 * (x[1] - x[0]));
            } else {
                const double dl = (dv[last] - dv[last - 1]) / (x[last] - x[last - 1]);
                const double dr = (dv[1] - dv[0]) / (x[1] - x[0]);
                dm[0] = (dl + dr) / 2;
            }
        }
        if (c.equidistant) {
            const double hx2 = 2 * (x[1] - x[0]);
            for (int i = 1; i < last; ++i) dm[i] = (dv[i + 1] - dv[i - 1]) / hx2;
        } else {
            for (int i = 1; i < last; ++i) {
                const double dl = (dv[i] - dv[i - 1]) / (x[i] - x[i - 1]);
                const double dr = (dv[i + 1] - dv[i]) / (x[i + 1] - x[i]);
                dm[i] = (dl + dr) / 2;
            }
        }
        if (c.bc_last == SPL_BC_GIVEN) dm[last] = (dv[last] - dv[last - 1]) / (x[last] - x[last - 1]);
        else if (c.bc_last == SPL_BC_ZERO_DERIVATIVE) dm[last] = 0;
        else if (c.bc_last == SPL_BC_PERIODIC) dm[last] = dm[0];
    }
    if (c.bc_first == SPL_BC_ZERO_DERIVATIVE) dm[0] = 0.0;
    if (c.bc_last == SPL_BC_ZERO_DERIVATIVE) dm[last] = 0.0;
}

// (scaled) node values / slopes of the value polynomial, and of the sensitivity polynomial of one parameter
template <int NN>
AB_HD void hs_scaled(const SplineCfg& c, const int nn, const double (&v)[NN], const double (&m)[NN], double (&P)[NN], double (&S)[NN]) {
    for (int i = 0; i < nn; ++i) {
        P[i] = c.logarithmic ? log(v[i]) : v[i];
        S[i] = c.logarithmic ? m[i] / v[i] : m[i];
    }
}
template <int NN>
AB_HD void hs_scaled_sensi(const SplineCfg& c, const int nn, const double (&v)[NN], const double (&m)[NN], const double (&dv)[NN],
                           const double (&dm)[NN], double (&P)[NN], double (&S)[NN]) {
    for (int i = 0; i < nn; ++i) {
        P[i] = c.logarithmic ? dv[i] / v[i] : dv[i];
        S[i] = c.logarithmic ? (dm[i] - dv[i] * m[i] / v[i]) / v[i] : dm[i];
    }
}

// Horner coefficients (d, c, b, a) of interval i from its end values / slopes (compute_coefficients, get_coeffs_sensi_lowlevel:
// the two group the same polynomial differently, and the rounding is kept)
AB_HD void hs_interval(const bool sensi, const double len, const double P0, const double S0, const double P1, const double S1, double (&co)[4]) {
    co[0] = P0;
    co[1] = len * S0;
    if (!sensi) {
        co[2] = -3 * P0 - 2 * len * S0 + 3 * P1 - len * S1;
        co[3] = 2 * P0 + len * S0 - 2 * P1 + len * S1;
    } else {
        co[2] = 3 * (P1 - P0) - len * (2 * S0 + S1);
        co[3] = 2 * (P0 - P1) + len * (S0 + S1);
    }
}

// Extrapolation lines b + t * a before the first / after the last node: ext = [b_first, a_first, b_last, a_last]
// (compute_coefficients_extrapolation; for a sensitivity polynomial compute_coefficients_extrapolation_sensi, where the slope
// sensitivity of a `given` finite-difference boundary is re-derived from the value sensitivities: that equals dm[] here)
template <int NN>
AB_HD void hs_extrapolation(const SplineCfg& c, const int nn, const double (&x)[NN], const double (&P)[NN], const double (&S)[NN], double (&ext)[4]) {
    const int last = nn - 1;
    ext[0] = ext[1] = ext[2] = ext[3] = 0.0;
    if (c.ep_first == SPL_EP_CONSTANT) {
        ext[0] = P[0];
    } else if (c.ep_first == SPL_EP_LINEAR) {
        ext[0] = P[0] - x[0] * S[0];
        ext[1] = S[0];
    }
    if (c.ep_last == SPL_EP_CONSTANT) {
        ext[2] = P[last];
    } else if (c.ep_last == SPL_EP_LINEAR) {
        ext[2] = P[last] - x[last] * S[last];
        ext[3] = S[last];
    }
}
// Same for the sensitivity polynomial of one parameter (compute_coefficients_extrapolation_sensi, :582-734).  It runs AFTER the
// translation to the logarithmic parametrisation (:431-444), so it sees the scaled (Ps, Ss); the slope of a linear
// extrapolation is 0 for a zero-derivative boundary, the given slope sensitivity for given slopes, and — finite-difference
// slopes with a `given` boundary — the one-sided difference of the SCALED value sensitivities (which differs from Ss in the
// logarithmic case); other boundary conditions throw there (spline_cfg_valid).
template <int NN>
AB_HD void hs_extrapolation_sensi(const SplineCfg& c, const int nn, const double (&x)[NN], const double (&Ps)[NN], const double (&Ss)[NN],
                                  double (&ext)[4]) {
    const int last = nn - 1;
    ext[0] = ext[1] = ext[2] = ext[3] = 0.0;
    if (c.ep_first == SPL_EP_CONSTANT || c.ep_first == SPL_EP_LINEAR) {
        double sm0 = 0.0;
        if (c.ep_first == SPL_EP_LINEAR && c.bc_first != SPL_BC_ZERO_DERIVATIVE)
            sm0 = c.fd ? (Ps[1] - Ps[0]) / (x[1] - x[0]) : Ss[0];
        ext[0] = Ps[0] - sm0 * x[0];
        ext[1] = sm0;
    }
    if (c.ep_last == SPL_EP_CONSTANT || c.ep_last == SPL_EP_LINEAR) {
        double sm1 = 0.0;
        if (c.ep_last == SPL_EP_LINEAR && c.bc_last != SPL_BC_ZERO_DERIVATIVE)
            sm1 = c.fd ? (Ps[last] - Ps[last - 1]) / (x[last] - x[last - 1]) : Ss[last];
        ext[2] = Ps[last] - sm1 * x[last];
        ext[3] = sm1;
    }
}

AB_HD double hs_horner(const double s, const double (&co)[4]) { return co[0] + s * (co[1] + s * (co[2] + s * co[3])); }

// Interval of t inside [x[0], x[nn-1]] (:916-934).  at_node: t equals the right node of the interval (non-equidistant search only;
// the reference then returns the stored node value for VALUES).
template <int NN>
AB_HD int hs_interval_of(const SplineCfg& c, const int nn, const double (&x)[NN], const double t, double& len, bool& at_node) {
    int i;
    at_node = false;
    if (c.equidistant) {
        len = x[1] - x[0];
        i = (int)trunc((t - x[0]) / len);
        i = i < nn - 2 ? i : nn - 2;
    } else {
        i = 0;
        while (x[i + 1] < t) ++i;
        at_node = (t == x[i + 1]);
        len = x[i + 1] - x[i];
    }
    return i;
}

// A spline with its value polynomial and the sensitivity polynomials of NPAR parameters in plain arrays (self-test / host use;
// the engine keeps the same numbers column-wise in shared memory, spline.cuh).
template <int NN, int NPAR>
struct HermiteSplineData {
    SplineCfg cfg;
    int nn, npar;
    double x[NN], P[NN];               // nodes, scaled node values
    double co[NN - 1][4], ext[4];      // value polynomial
    double sco[NPAR][NN - 1][4], sext[NPAR][4];
    double final_value_scaled, final_sens_scaled[NPAR];
    bool S_last_zero;                  // scaled slope at the last node == 0 (final sensitivity rule)
};

// Build from nodes (2 entries if equidistant), values, slopes (ignored if fd), dvaluesdp[node][par], dslopesdp[node][par]
template <int NN, int NPAR>
AB_HD void hs_build(HermiteSplineData<NN, NPAR>& H, const SplineCfg& c, const int nn, const double* nodes, const double* values,
                    const double* slopes, const int npar, const double* dvaluesdp, const double* dslopesdp) {
    H.cfg = c;
    H.nn = nn;
    H.npar = npar;
    double v[NN], m[NN], S[NN];
    if (c.equidistant) {
        const double step = (nodes[1] - nodes[0]) / (nn - 1);
        for (int i = 0; i < nn - 1; ++i) H.x[i] = nodes[0] + i * step;
        H.x[nn - 1] = nodes[1];
    } else {
        for (int i = 0; i < nn; ++i) H.x[i] = nodes[i];
    }
    for (int i = 0; i < nn; ++i) {
        v[i] = values[i];
        m[i] = c.fd ? 0.0 : slopes[i];
    }
    hs_slopes<NN>(c, nn, H.x, v, m);
    hs_scaled<NN>(c, nn, v, m, H.P, S);
    for (int i = 0; i < nn - 1; ++i) hs_interval(false, H.x[i + 1] - H.x[i], H.P[i], S[i], H.P[i + 1], S[i + 1], H.co[i]);
    hs_extrapolation<NN>(c, nn, H.x, H.P, S, H.ext);
    H.S_last_zero = (S[nn - 1] == 0.0);
    // compute_final_value (:762-806)
    const int last = nn - 1;
    double fv;
    if (c.ep_last == SPL_EP_CONSTANT) {
        fv = H.ext[2];
    } else if (c.ep_last == SPL_EP_LINEAR) {
        if (c.bc_last == SPL_BC_ZERO_DERIVATIVE) fv = H.ext[2];
        else if (H.ext[3] < 0) fv = -INFINITY;
        else if (H.ext[3] > 0) fv = INFINITY;
        else fv = H.ext[2];
    } else if (c.ep_last == SPL_EP_POLYNOMIAL) {
        const double* q = H.co[last - 1];
        if (q[3] < 0) fv = -INFINITY;
        else if (q[3] > 0) fv = INFINITY;
        else if (q[2] < 0) fv = -INFINITY;
        else if (q[2] > 0) fv = INFINITY;
        else if (q[1] < 0) fv = -INFINITY;
        else if (q[1] > 0) fv = INFINITY;
        else fv = q[0];
    } else {   // periodic (and noExtrapolation, which the reference treats alike here)
        fv = H.P[0];
        for (int i = 0; i < nn; ++i)
            if (H.P[i] != fv || S[i] != 0) { fv = NAN; break; }
    }
    H.final_value_scaled = fv;
    for (int ip = 0; ip < npar; ++ip) {
        double dv[NN], dm[NN], Ps[NN], Ss[NN];
        for (int i = 0; i < nn; ++i) {
            dv[i] = dvaluesdp[i * npar + ip];
            dm[i] = (c.fd || !dslopesdp) ? 0.0 : dslopesdp[i * npar + ip];
        }
        hs_slopes_sensi<NN>(c, nn, H.x, dv, dm);
        hs_scaled_sensi<NN>(c, nn, v, m, dv, dm, Ps, Ss);
        for (int i = 0; i < nn - 1; ++i) hs_interval(true, H.x[i + 1] - H.x[i], Ps[i], Ss[i], Ps[i + 1], Ss[i + 1], H.sco[ip][i]);
        hs_extrapolation_sensi<NN>(c, nn, H.x, Ps, Ss, H.sext[ip]);
        // compute_final_sensitivity (:808-844)
        double fs = 0.0;
        if (c.ep_last == SPL_EP_CONSTANT || (c.bc_last == SPL_BC_ZERO_DERIVATIVE && c.ep_last == SPL_EP_LINEAR)) fs = H.sext[ip][2];
        else if (c.ep_last == SPL_EP_LINEAR) fs = H.S_last_zero ? NAN : 0.0;
        else fs = NAN;
        H.final_sens_scaled[ip] = fs;
    }
}

// value / sensitivity polynomial at a time inside the node range
template <int NN, int NPAR>
AB_HD double hs_inside(const HermiteSplineData<NN, NPAR>& H, const double t, const int ip) {
    double len;
    bool at_node;
    const int i = hs_interval_of<NN>(H.cfg, H.nn, H.x, t, len, at_node);
    if (ip < 0 && at_node) return H.P[i + 1];   // exact on nodes, values only
    return hs_horner((t - H.x[i]) / len, ip < 0 ? H.co[i] : H.sco[ip][i]);
}

// get_value_scaled (:846-944); *err is set when the reference throws (evaluation outside the nodes without extrapolation)
template <int NN, int NPAR>
AB_HD double hs_value_scaled(const HermiteSplineData<NN, NPAR>& H, const double t, bool* err) {
    const SplineCfg& c = H.cfg;
    const int last = H.nn - 1;
    if (isinf(t)) return H.final_value_scaled;
    if (t > H.x[last]) {
        if (c.ep_last == SPL_EP_NONE) { *err = true; return NAN; }
        if (c.ep_last == SPL_EP_CONSTANT) return H.ext[2];
        if (c.ep_last == SPL_EP_LINEAR) return H.ext[2] + t * H.ext[3];
        if (c.ep_last == SPL_EP_POLYNOMIAL) return hs_horner((t - H.x[last - 1]) / (H.x[last] - H.x[last - 1]), H.co[last - 1]);
        const double len = H.x[last] - H.x[0];   // periodic: the UNSCALED value of the wrapped time (reference quirk)
        const double vs = hs_inside(H, H.x[0] + fmod(t - H.x[0], len), -1);
        return c.logarithmic ? exp(vs) : vs;
    }
    if (t < H.x[0]) {
        if (c.ep_first == SPL_EP_NONE) { *err = true; return NAN; }
        if (c.ep_first == SPL_EP_CONSTANT) return H.ext[0];
        if (c.ep_first == SPL_EP_LINEAR) return H.ext[0] + t * H.ext[1];
        if (c.ep_first == SPL_EP_POLYNOMIAL) return hs_horner((t - H.x[0]) / (H.x[1] - H.x[0]), H.co[0]);
        const double len = H.x[last] - H.x[0];
        const double vs = hs_inside(H, H.x[last] + fmod(t - H.x[0], len), -1);
        return c.logarithmic ? exp(vs) : vs;
    }
    return hs_inside(H, t, -1);
}
template <int NN, int NPAR>
AB_HD double hs_value(const HermiteSplineData<NN, NPAR>& H, const double t, bool* err) {
    const double y = hs_value_scaled(H, t, err);
    return H.cfg.logarithmic ? exp(y) : y;
}

// get_sensitivity_scaled (:946-1049) and get_sensitivity (:64-67)
template <int NN, int NPAR>
AB_HD double hs_sensitivity_scaled(const HermiteSplineData<NN, NPAR>& H, const double t, const int ip, bool* err) {
    const SplineCfg& c = H.cfg;
    const int last = H.nn - 1;
    if (isinf(t)) return H.final_sens_scaled[ip];
    if (t > H.x[last]) {
        if (c.ep_last == SPL_EP_NONE) { *err = true; return NAN; }
        if (c.ep_last == SPL_EP_CONSTANT) return H.sext[ip][2];
        if (c.ep_last == SPL_EP_LINEAR) return H.sext[ip][2] + t * H.sext[ip][3];
        if (c.ep_last == SPL_EP_POLYNOMIAL) return hs_horner((t - H.x[last - 1]) / (H.x[last] - H.x[last - 1]), H.sco[ip][last - 1]);
        const double len = H.x[last] - H.x[0];   // periodic: the UNSCALED sensitivity of the wrapped time (reference quirk)
        const double tw = H.x[0] + fmod(t - H.x[0], len);
        const double sw = hs_inside(H, tw, ip);
        return c.logarithmic ? sw * exp(hs_inside(H, tw, -1)) : sw;
    }
    if (t < H.x[0]) {
        if (c.ep_first == SPL_EP_NONE) { *err = true; return NAN; }
        if (c.ep_first == SPL_EP_CONSTANT) return H.sext[ip][0];
        if (c.ep_first == SPL_EP_LINEAR) return H.sext[ip][0] + t * H.sext[ip][1];
        if (c.ep_first == SPL_EP_POLYNOMIAL) return hs_horner((t - H.x[0]) / (H.x[1] - H.x[0]), H.sco[ip][0]);
        const double len = H.x[last] - H.x[0];
        const double tw = H.x[last] + fmod(t - H.x[0], len);
        const double sw = hs_inside(H, tw, ip);
        return c.logarithmic ? sw * exp(hs_inside(H, tw, -1)) : sw;
    }
    return hs_inside(H, t, ip);
}
template <int NN, int NPAR>
AB_HD double hs_sensitivity(const HermiteSplineData<NN, NPAR>& H, const double t, const int ip, bool* err) {
    const double s = hs_sensitivity_scaled(H, t, ip, err);
    if (!H.cfg.logarithmic) return s;
    if (isinf(t)) {   // get_final_sensitivity (:104-125)
        const double v = exp(H.final_value_scaled);
        return isinf(v) ? s : s * v;
    }
    return s * hs_value(H, t, err);
}

}  // namespace amici_b200
#endif
