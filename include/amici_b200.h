/* amici_b200.h — C ABI of the B200 batched BDF engine.
 *
 * Drop-in boundary for ONE path of AMICI: the batch call
 *     amici::run_simulations(Solver const&, std::vector<ExpData*> const&, Model const&, bool failfast, int num_threads)
 * (reference include/amici/amici.h:36-39, src/amici.cpp:220-264; Python face
 * amici.sim.sundials.run_simulations, reference python/sdist/amici/sim/sundials/_swig_wrappers.py:87-133).
 * The reference has no FFI for this path (it is C++ behind SWIG); the entry points below are what a
 * binding of that call needs: plain pointers and sizes, no C++/torch types, integer return codes,
 * no exceptions across the boundary.  INTEGRATION.md shows the reference-side stub.
 *
 * Conventions
 *  - all arrays row-major doubles/ints, borrowed from the caller, never retained after return;
 *  - one `ab_batch` describes N conditions of one model that share ts / plist / pscale
 *    (the per-condition fields of amici::ExpData, reference include/amici/edata.h:20,
 *    include/amici/simulation_parameters.h, flattened to struct-of-arrays);
 *  - outputs use the reference's ReturnData layouts per condition (reference
 *    python/sdist/amici/sim/sundials/_numpy.py:414-459) with the condition index as the slowest
 *    dimension; a NULL output pointer means "do not report";
 *  - per-condition failures never make the call fail: they are reported in status[] with the
 *    reference's codes (reference include/amici/defines.h:64-101) and NaN-filled outputs from the
 *    failing timepoint on (reference src/rdata.cpp:684-735).
 */
#ifndef AMICI_B200_H
#define AMICI_B200_H

#ifdef __cplusplus
extern "C" {
#endif

/* Solver settings: the subset of amici::Solver that parametrises this path
 * (defaults and meaning: reference include/amici/solver.h:1869-2127). */
typedef struct ab_solver_options {
    double rtol, atol;         /* Solver::set_relative_tolerance / set_absolute_tolerance           */
    double rtol_fsa, atol_fsa; /* resolved values (NaN already replaced by rtol/atol, solver.cpp:874) */
    double t0;                 /* Model::set_t0                                                      */
    double hmax_inv;           /* 1 / Solver::set_max_step_size, 0 = unbounded                       */
    int sensi;                 /* SensitivityOrder: 0 none, 1 first (forward)                        */
    int ism;                   /* InternalSensitivityMethod: 1 simultaneous, 2 staggered (lane-group models) */
    int maxsteps;              /* Solver::set_max_steps (budget for the whole simulation)            */
    int stldet;                /* Solver::set_stability_limit_flag                                   */
    int max_nonlin_iters;      /* Solver::set_max_nonlin_iters (3)                                   */
    int max_conv_fails;        /* Solver::set_max_conv_fails (10)                                    */
    int max_err_test_fails;    /* CVODES default 7                                                   */
    int reporting;             /* RDataReporting: 0 full, 1 residuals, 2 likelihood, 3 observables_likelihood: 2 skips the
                                  per-timepoint outputs; otherwise the non-NULL output pointers select what is reported       */
    /* steady-state (pre-equilibration) settings: amici::Solver / amici::Model, reference
     * include/amici/solver.h:2040-2100, include/amici/defines.h:204-232 */
    double rtol_ss, atol_ss;             /* resolved get_{relative,absolute}_tolerance_steady_state (solver.cpp:981-998)        */
    double rtol_ss_sensi, atol_ss_sensi; /* resolved ..._steady_state_sensi (solver.cpp:1020-1037)                             */
    double newton_damping_lower_bound;   /* Solver::set_newton_damping_factor_lower_bound (1e-8)                               */
    int newton_maxsteps;       /* Solver::set_newton_max_steps (0 = Newton's method off)                                  */
    int newton_damping_mode;   /* NewtonDampingFactorMode: 0 off, 1 on                                                    */
    int sensi_ss_check;        /* Solver::set_sensi_steady_state_check (1)                                                */
    int ss_mode;               /* SteadyStateComputationMode: 0 newtonOnly, 1 integrationOnly, 2 integrateIfNewtonFails   */
    int ss_sensi_mode;         /* SteadyStateSensitivityMode: same values                                                 */
    int nmaxevent;             /* Model::set_n_max_event (10): event-observable slots per event                           */
    int posteq;                /* 1 if ts ends with t = inf output points: post-equilibration after the last finite timepoint
                                  (reference src/forwardproblem.cpp:376-387); ab_run_batch sets it itself from the host ts  */
    /* ExpData::reinitialization_state_idxs_sim / _presim as bit sets (bit i = state i; all ones =
       ExpData::reinitialize_fixed_parameter_initial_states): after a pre-equilibration or presimulation, the selected states whose
       initial value depends on fixed parameters are reset to x0(k) of the next phase and their sensitivities to 0
       (Model::reinitialize, reference src/model.cpp:280-295, src/forwardproblem.cpp:316, :350-352). Models with nx <= 64. */
    unsigned long long reinit_states_sim, reinit_states_presim;
    /* Model::set_state_is_non_negative as a bit set (bit i = state i): every model function sees max(x_i, 0) for the selected
       states, the reported x included (Model::compute_x_pos, reference src/model.cpp:3088-3108).  Lane-group models, nx <= 64. */
    unsigned long long x_nonneg;
    /* Solver::set_max_time in seconds, 0 = unlimited (reference src/solver.cpp:1050-1069, src/solver_cvodes.cpp:1177): a condition
       whose integration has run longer than this on the device ends with status AMICI_MAX_TIME_EXCEEDED (-1000). */
    double maxtime;
    /* SensitivityMethod with sensi = 1: 0 or 1 forward (default), 2 adjoint (reference include/amici/defines.h:160-165,
       src/backwardproblem.cpp).  Adjoint: event-free lane-group models, plain simulations (no pre-/post-equilibration or
       presimulation); reports llh, sllh and the sensitivity-free fields, sx / sy / sres / FIM are not computed (NaN). */
    int sensi_meth;
    int reserved_;
} ab_solver_options;

/* One batch of N conditions (replaces std::vector<ExpData*> in, std::vector<ReturnData> out). */
typedef struct ab_batch {
    /* ---- inputs ---- */
    int N, nt, nplist;
    const double* ts;        /* [nt]            ExpData::ts_ / Model::set_timepoints              */
    const int* plist;        /* [nplist]        SimulationParameters::plist                       */
    const int* pscale;       /* [np]            SimulationParameters::pscale (0 none, 1 ln, 2 log10) */
    const double* p;         /* [N][np]         SimulationParameters::free_parameters (scaled)    */
    const double* k;         /* [N][nk]         SimulationParameters::fixed_parameters            */
    const double* my;        /* [N][nt][ny]     ExpData measurements, NaN = missing; may be NULL  */
    const double* sigmay_in; /* [N][nt][ny]     ExpData noise scales, NaN = model sigma; may be NULL */
    const double* k_preeq;   /* [N][nk]         ExpData::fixed_parameters_pre_equilibration; NULL = no pre-equilibration */
    const double* mz;        /* [N][nmaxevent][nz]  ExpData event measurements, NaN = missing; may be NULL */
    const double* sigmaz_in; /* [N][nmaxevent][nz]  ExpData event noise scales, NaN = model sigma; may be NULL */
    const double* x0_in;     /* [N][nx]         ExpData::x0, initial state instead of the model's x0(p, k); NULL = model's (reference src/edata.cpp:472-480) */
    const double* sx0_in;    /* [N][nplist][nx] ExpData::sx0 on the parameter scale given by pscale; NULL = model's (src/edata.cpp:482-490, src/model.cpp:986-1024) */
    const double* k_presim;  /* [N][nk]         ExpData::fixed_parameters_presimulation; NULL = the simulation's k (reference src/edata.cpp:531-548) */
    const double* t_presim;  /* [N]             ExpData::t_presim: integrate from t0 - t_presim to t0 before the main simulation where > 0; NULL = none (src/forwardproblem.cpp:284-321) */
    /* ---- outputs (ReturnData fields; any may be NULL) ---- */
    int* status;             /* [N]                                  */
    double* llh;             /* [N]                                  */
    double* chi2;            /* [N]                                  */
    double* sllh;            /* [N][nplist]                          */
    double* x;               /* [N][nt][nx]                          */
    double* sx;              /* [N][nt][nplist][nx]                  */
    double* y;               /* [N][nt][ny]                          */
    double* sy;              /* [N][nt][nplist][ny]                  */
    double* sigmay;          /* [N][nt][ny]                          */
    double* ssigmay;         /* [N][nt][nplist][ny]                  */
    double* res;             /* [N][nt*ny]                           */
    double* sres;            /* [N][nt*ny][nplist]                   */
    double* FIM;             /* [N][nplist][nplist]                  */
    double* x0;              /* [N][nx]                              */
    double* sx0;             /* [N][nplist][nx]                      */
    double* xdot;            /* [N][nx]                              */
    double* J;               /* [N][nx][nx]                          */
    int* numsteps;           /* [N][nt]                              */
    int* numrhsevals;        /* [N][nt]                              */
    int* numerrtestfails;    /* [N][nt]                              */
    int* numnonlinsolvconvfails; /* [N][nt]                          */
    int* order;              /* [N][nt]                              */
    /* pre-equilibration (ReturnData::x_ss, sx_ss, preeq_*; reference src/rdata.cpp:261-286) */
    double* x_ss;            /* [N][nx]                              */
    double* sx_ss;           /* [N][nplist][nx]                      */
    int* preeq_status;       /* [N][3]  SteadyStateStatus of (Newton, simulation, Newton) */
    int* preeq_numsteps;     /* [N][3]                               */
    double* preeq_wrms;      /* [N]                                  */
    double* preeq_t;         /* [N]                                  */
    /* post-equilibration (ReturnData::posteq_*; reference src/rdata.cpp:288-312) */
    int* posteq_status;      /* [N][3]                               */
    int* posteq_numsteps;    /* [N][3]                               */
    double* posteq_wrms;     /* [N]                                  */
    double* posteq_t;        /* [N]                                  */
    /* event observables (ReturnData::z, sz, rz, srz, sigmaz; reference src/rdata.cpp:434-512) */
    double* z;               /* [N][nmaxevent][nz]                   */
    double* sz;              /* [N][nmaxevent][nplist][nz]           */
    double* rz;              /* [N][nmaxevent][nz]                   */
    double* srz;             /* [N][nmaxevent][nplist][nz]           */
    double* sigmaz;          /* [N][nmaxevent][nz]                   */
} ab_batch;

/* Model dimensions (amici::ModelDimensions, reference include/amici/model_dimensions.h:14). */
typedef struct ab_dims {
    int nx, np, nk, ny, nz, ne, nspl;
    int group_lanes;   /* lanes of one warp that integrate one condition (FSA kernel) */
    int cols_per_lane; /* state/sensitivity columns per lane                          */
} ab_dims;

/* error codes of the ABI itself (not per-condition status) */
#define AB_OK 0
#define AB_E_UNKNOWN_MODEL -1
#define AB_E_BAD_ARGUMENT -2
#define AB_E_CUDA -3
#define AB_E_UNSUPPORTED -4

/* Registry of compiled models (replaces amici::generic_model::get_model(), reference
 * models/model_*_py/wrapfunctions.cpp). */
int ab_model_count(void);
const char* ab_model_name(int index);
int ab_get_dims(const char* model, ab_dims* out);
/* nominal parameters (Model::get_unscaled_parameters / get_fixed_parameters) */
int ab_get_nominal(const char* model, double* p, double* k);

/* Solve a batch whose arrays are all HOST memory: copies inputs to the device, integrates every
 * condition on the GPU of the calling thread's current device, copies the requested outputs back.
 * This is the replacement for amici::run_simulations (reference src/amici.cpp:220-264).
 * kernel_ms (may be NULL) receives the device time of the solve kernel alone.
 * Outputs are streamed: the batch is one launch (per 16 GB of device buffers), and the outputs of each completed
 * sub-batch (1/64 of the batch, 512 .. 8192 conditions) are copied back while the kernel keeps integrating the rest. Page-locked output
 * arrays make those copies asynchronous. The device buffers are kept for the next call (ab_release_workspace). */
int ab_run_batch(const char* model, const ab_solver_options* opt, const ab_batch* host_batch, float* kernel_ms);
/* Free the device workspace ab_run_batch keeps between calls (the reference's counterpart is the per-thread
 * Solver/Model clone that run_simulations destroys after every condition, src/amici.cpp:241-242). */
void ab_release_workspace(void);

/* ---- list-of-conditions entry point: std::vector<ExpData*> as it is --------------------------------------------------
 * One condition = one amici::ExpData (reference include/amici/edata.h:20, include/amici/simulation_parameters.h:120-215).
 * A NULL pointer / zero count means "not set": the value of ab_model_state applies, exactly as ConditionContext does it
 * (reference src/edata.cpp:409-581). */
typedef struct ab_condition {
    int none;                 /* 1: the ExpData pointer is NULL (simulate the model as it is); every other member is ignored */
    int nt; const double* ts; /* ExpData::ts_                                                                              */
    int nplist; const int* plist;   /* SimulationParameters::plist                                                          */
    const int* pscale;        /* [np]  SimulationParameters::pscale; without `p` the model's parameters are re-scaled so that
                                 their unscaled values stay (Model::set_parameter_scale, reference src/model.cpp:513-527)     */
    const double* p;          /* [np]  free_parameters, on the scale pscale                                                */
    const double* k;          /* [nk]  fixed_parameters                                                                    */
    const double* k_preeq;    /* [nk]  fixed_parameters_pre_equilibration: pre-equilibrate                                 */
    const double* k_presim;   /* [nk]  fixed_parameters_presimulation                                                      */
    double t_presim;          /*       presimulate over [t_start - t_presim, t_start] if > 0                               */
    double t_start;           /*       ExpData::t_start (replaces Model::t0 whenever an ExpData is given)                  */
    const double* x0;         /* [nx]              */
    const double* sx0;        /* [nplist][nx]      */
    const double* my;         /* [nt][ny]   NaN = missing            */
    const double* sigmay;     /* [nt][ny]   NaN = the model's sigma  */
    const double* mz;         /* [nmaxevent][nz]   */
    const double* sigmaz;     /* [nmaxevent][nz]   */
    int reinit_all;           /* ExpData::reinitialize_fixed_parameter_initial_states                                      */
    int n_reinit_sim; const int* reinit_sim;        /* reinitialization_state_idxs_sim    */
    int n_reinit_presim; const int* reinit_presim;  /* reinitialization_state_idxs_presim */
} ab_condition;

/* The amici::Model settings a condition falls back to (Model::get_timepoints / get_parameter_list / get_parameter_scale /
 * get_free_parameters / get_fixed_parameters / t0 / get_reinitialize_fixed_parameter_initial_states /
 * get_reinitialization_state_idxs). */
typedef struct ab_model_state {
    int nt; const double* ts;
    int nplist; const int* plist;
    const int* pscale;        /* [np] */
    const double* p;          /* [np] on the scale pscale */
    const double* k;          /* [nk] */
    double t0;
    int reinit_all; int n_reinit; const int* reinit_idxs;
} ab_model_state;

/* Replacement for amici::run_simulations(solver, std::vector<ExpData*>, model, failfast, num_threads) (reference
 * src/amici.cpp:220-264) on a list of n conditions: applies the ConditionContext rules, groups the conditions that can share
 * a launch, runs ab_run_batch per group and writes every condition's results to ITS OWN buffers: outs[i] is an ab_batch whose
 * OUTPUT pointers (N = 1 layouts: x[nt][nx], sx[nt][nplist][nx], ...) point into the caller's per-condition storage, e.g. the
 * std::vector members of a ReturnData; its input members are ignored.  failfast != 0: conditions after the first failed one
 * get status AMICI_NOT_RUN (-1001) and their buffers are not touched (single-thread semantics of the reference).
 * opt->t0 / reinit_states_* / posteq are set per group by the call.  n_batches (may be NULL): number of launches. */
int ab_run_conditions(const char* model, const ab_solver_options* opt, const ab_model_state* ms, int n, const ab_condition* conds,
                      const ab_batch* outs, int failfast, int* n_batches);

/* Solve a batch whose arrays are all DEVICE memory (ts/plist/pscale included), asynchronously on
 * `stream` (a cudaStream_t passed as void*). `work_counter` is one device int used as the
 * trajectory queue head; it is reset by the call. res/sres/FIM are zero-filled by the call. */
int ab_solve_device(const char* model, const ab_solver_options* opt, const ab_batch* device_batch,
                    int* work_counter, void* stream);

/* Device reduction of the objective over the batch: out[0] = sum llh, out[1+i] = sum sllh[:, i],
 * out[1+nplist] = number of failed conditions (status != 0; their llh/sllh are skipped).
 * All pointers device memory; asynchronous on `stream`.  The result is what a multi-GPU caller
 * all-reduces (the reference sums in Python: petab/v1/_simulations.py:213-232). */
int ab_reduce_objective(int N, int nplist, const int* status, const double* llh, const double* sllh,
                        double* out, void* stream);

/* Measured FP64 FMA throughput of the current device in TFLOP/s (dependent-FMA chains, all SMs busy);
 * bench.py uses it as the FP64 roofline denominator. */
int ab_measure_fp64_peak(double* tflops);

/* Number of kernels launched by this library since load (for bench accounting). */
long ab_kernel_launch_count(void);

/* Device self-test (test hook, not a product path): the engine's forms of CVODES' stability-limit detection (cvSLdet,
 * reference ThirdParty/sundials/src/cvodes/cvodes.c:8256-8570) on n caller-supplied ring buffers ssdat[n][5][8] (host memory;
 * entry [i][1..3] = squared scaled derivative norms, [i][5..7] = their ratios to the next older entry), newest entry at
 * head[n], order q[n] in 3..5.  out[n][4] = results of (all orders per lane, three lanes + shuffles, three lanes + shared
 * memory in a 16-lane group, the same in a 4-lane group); all four must be equal. */
int ab_selftest_sldet(int n, const double* ssdat, const int* head, const int* q, int* out);

/* Self-test of the device Hermite spline (csrc/hermite_spline.cuh; replaces amici::HermiteSpline, reference
 * include/amici/splinefunctions.h:219-424, src/splinefunctions.cpp) on caller-supplied splines, one per case:
 * cfg[c][8] = {n_nodes (2..8), first / last SplineBoundaryCondition, first / last SplineExtrapolation (numeric values of
 * reference include/amici/defines.h:249-264), node_derivative_by_FD, equidistant_spacing, logarithmic_parametrization};
 * nodes / values / slopes [c][8] (equidistant: nodes[c][0..1] = first, last), dvaluesdp / dslopesdp [c][8][npar] (npar <= 4),
 * ts[c][nt] (may contain +inf: the final value).  Out: value_out[c][nt] = get_value, sens_out[c][nt][npar] = get_sensitivity,
 * err_out[c][nt] = 1 where the reference throws (evaluation outside the nodes without extrapolation; outputs NaN). */
int ab_selftest_spline(int ncases, const int* cfg, const double* nodes, const double* values, const double* slopes, int npar,
                       const double* dvaluesdp, const double* dslopesdp, int nt, const double* ts, double* value_out,
                       double* sens_out, int* err_out);

/* Last CUDA error string seen by the library on this thread ("" if none). */
const char* ab_last_error(void);

/* The solve kernel this thread launched last, as "kernel<model,G,C,threads,minblocks,variant>" (G lanes per condition,
 * C columns per lane), "" if none: which instantiation served a batch (bench.py reports it in `roofline.kernel`). */
const char* ab_last_kernel_name(void);

#ifdef __cplusplus
}
#endif
#endif /* AMICI_B200_H */
