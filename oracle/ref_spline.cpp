// oracle/ref_spline.cpp — TEST INFRASTRUCTURE (checker), not product code.
//
// C entry point around the UNMODIFIED reference class amici::HermiteSpline (reference include/amici/splinefunctions.h:219-424,
// src/splinefunctions.cpp), linked from oracle/_ref/libamici.a: values and parameter sensitivities of one spline at a list
// of times, with the call sequence of the reference's own unit tests (tests/cpp/unittests/testSplines.cpp): constructor,
// compute_coefficients, compute_coefficients_sensi, compute_final_value, compute_final_sensitivity, get_value / get_sensitivity.
// An evaluation the reference refuses (AmiException) is reported as NaN with bit i of the returned mask... kept simple:
// thrown[i] = 1.  Returns 0, or -1 if the constructor / coefficient set-up throws.
#include <amici/defines.h>
#include <amici/exception.h>
#include <amici/splinefunctions.h>

#include <cmath>
#include <vector>

extern "C" int ref_spline_eval(int n_nodes, int n_nodes_given, const double* nodes, const double* values, const double* slopes,
                               int bc_first, int bc_last, int ep_first, int ep_last, int fd, int equidistant, int logarithmic,
                               int npar, const double* dvaluesdp, const double* dslopesdp, int nt, const double* ts,
                               double* value_out, double* sens_out, int* thrown) {
    using namespace amici;
    try {
        std::vector<double> x(nodes, nodes + n_nodes_given), v(values, values + n_nodes), m;
        if (!fd) m.assign(slopes, slopes + n_nodes);
        HermiteSpline spline(x, v, m, static_cast<SplineBoundaryCondition>(bc_first), static_cast<SplineBoundaryCondition>(bc_last),
                             static_cast<SplineExtrapolation>(ep_first), static_cast<SplineExtrapolation>(ep_last), fd != 0,
                             equidistant != 0, logarithmic != 0);
        spline.compute_coefficients();
        std::vector<double> dv, dm;
        if (npar > 0) {
            dv.assign(dvaluesdp, dvaluesdp + (size_t)n_nodes * npar);
            if (fd) dm.assign((size_t)n_nodes * npar, 0.0);
            else dm.assign(dslopesdp, dslopesdp + (size_t)n_nodes * npar);
            spline.compute_coefficients_sensi(npar, 0, dv, dm);
        }
        spline.compute_final_value();
        if (npar > 0) spline.compute_final_sensitivity(npar, 0, dv, dm);
        for (int i = 0; i < nt; ++i) {
            thrown[i] = 0;
            try {
                value_out[i] = spline.get_value(ts[i]);
            } catch (AmiException const&) {
                value_out[i] = NAN;
                thrown[i] = 1;
            }
            for (int ip = 0; ip < npar; ++ip) {
                try {
                    sens_out[(size_t)i * npar + ip] = std::isinf(ts[i]) ? spline.get_final_sensitivity(ip) : spline.get_sensitivity(ts[i], ip);
                } catch (AmiException const&) {
                    sens_out[(size_t)i * npar + ip] = NAN;
                    thrown[i] = 1;
                }
            }
        }
    } catch (std::exception const&) {
        return -1;
    }
    return 0;
}
