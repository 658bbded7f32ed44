// Common macros for code shared between the CUDA kernels and the host-compiled
// lane-group emulation used by the CPU unit tests (tests/cpu_emul).
#ifndef AMICI_B200_MODEL_COMMON_CUH
#define AMICI_B200_MODEL_COMMON_CUH

#include <cmath>
#include <cfloat>

#if defined(__CUDACC__)
#define AB_HD __host__ __device__ __forceinline__
#define AB_COLD __host__ __device__ __noinline__
#define AB_UNROLL _Pragma("unroll")
#define AB_NOUNROLL _Pragma("unroll 1")
#else
#define AB_HD inline
#define AB_COLD __attribute__((noinline))
#define AB_UNROLL
#define AB_NOUNROLL
#endif

// AB_FULL = 1 in the translation units of the full-featured kernel variants (tu/<model>_preeq.cu, event models) and in the host
// emulation: state non-negativity (Model::set_state_is_non_negative) and the wall-clock guard (Solver::set_max_time) are compiled
// there only, so that the headline kernels' step loop does not grow; launch.cuh routes batches that use them to those variants.
#ifndef AB_FULL
#define AB_FULL 0
#endif

namespace amici_b200 {

// status / return codes: numeric values of the reference (include/amici/defines.h:64-101),
// which are SUNDIALS' CV_* codes (static_asserts in src/solver_cvodes.cpp:40-61).
enum : int {
    AB_SUCCESS = 0,
    AB_TSTOP_RETURN = 1,
    AB_ROOT_RETURN = 2,
    AB_TOO_MUCH_WORK = -1,
    AB_TOO_MUCH_ACC = -2,
    AB_ERR_FAILURE = -3,
    AB_CONV_FAILURE = -4,
    AB_LSETUP_FAIL = -6,
    AB_RHSFUNC_FAIL = -8,
    AB_FIRST_RHSFUNC_ERR = -9,
    AB_REPTD_RHSFUNC_ERR = -10,
    AB_UNREC_RHSFUNC_ERR = -11,
    AB_IC_FAIL = -14,              // IDA_NO_RECOVERY: no consistent initial values (IDACalcIC, DAE models)
    AB_ILL_INPUT = -22,
    AB_REPTD_SRHSFUNC_ERR = -43,   // CV_REPTD_SRHSFUNC_ERR (staggered corrector)
    AB_TOO_CLOSE = -27,
    AB_NO_STEADY_STATE = -81,
    AB_DAMPING_FACTOR_ERROR = -86,
    AB_ERROR = -99,
    AB_SINGULAR_JACOBIAN = -9987,
    AB_T_OVERFLOW = -1002,
    AB_NOT_RUN = -1001,
    AB_MAX_TIME_EXCEEDED = -1000,
};

// parameter scaling (reference include/amici/defines.h ParameterScaling)
enum : int { AB_PSCALE_NONE = 0, AB_PSCALE_LN = 1, AB_PSCALE_LOG10 = 2 };

#if !defined(__CUDACC__)
using std::isfinite;
using std::isinf;
#endif

// a / b for finite, normal, non-zero b: reciprocal seed + one Newton step, then the product — 4 instructions, relative
// error ~2^-40 (1e-12).  Every use is in step-control arithmetic (error weights, BDF coefficients, step-size and order
// heuristics, stability-limit detection), where a relative perturbation of 1e-12 is four orders below the integration
// tolerance; the IEEE sequence is ~25 instructions plus a slow-path call, and with ~45 divisions per step the code size
// matters as much as the count (the step loop competes for the SM's instruction cache, DESIGN.md §2.6: dropping the
// residual-correction step, 2 more instructions per site, was worth +3.7 %).  Host builds divide.
AB_HD double fast_div(double a, double b) {
#if defined(__CUDA_ARCH__) && !defined(AB_EXACT_DIV)   // AB_EXACT_DIV: A/B experiments only (tools/event_spread.py)
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
    r = fma(r, fma(-b, r, 1.0), r);   // seed is good to ~2^-20 (rcp.approx.ftz.f64 reads the upper 32 bits): 2^-40 here
    return a * r;
#else
    return a / b;
#endif
}

// a / b to ~1 ulp: fast_div plus one residual correction (two more FMAs).  For the quotient that must reproduce its numerator when
// multiplied back: eta = hprime / h at the tstop clipping, where tn + h * eta has to land within 100 ulp of tstop.
AB_HD double fast_div_1ulp(double a, double b) {
#if defined(__CUDA_ARCH__) && !defined(AB_EXACT_DIV)
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
    r = fma(r, fma(-b, r, 1.0), r);
    const double q = a * r;
    return fma(fma(-q, b, a), r, q);
#else
    return a / b;
#endif
}

// 1 / b for finite, normal, non-zero b (pivots of the emitted sparse LU): reciprocal seed + two Newton steps, within 1-2 ulp
AB_HD double fast_rcp(double b) {
#if defined(__CUDA_ARCH__) && !defined(AB_EXACT_DIV)
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
    r = fma(r, fma(-b, r, 1.0), r);
    return fma(r, fma(-b, r, 1.0), r);
#else
    return 1.0 / b;
#endif
}

AB_HD double ab_max(double a, double b) { return a > b ? a : b; }
AB_HD double ab_min(double a, double b) { return a < b ? a : b; }

}  // namespace amici_b200
#endif
