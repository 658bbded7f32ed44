/* oracle/ref_driver.h — TEST INFRASTRUCTURE (checker), not product code.
 *
 * Flat C interface around the UNMODIFIED reference `amici::run_simulations`
 * (reference src/amici.cpp:220-264) for one pre-generated model, so that Python
 * tests / bench.py can push the very same SoA batch through the reference's
 * CVODES path and through the CUDA engine.  Built by oracle/build_ref.sh into
 * oracle/_ref/libref_<model>.so.
 *
 * All arrays are row-major, caller-allocated; a NULL output pointer means
 * "do not report this field". Integer enums use the reference's numeric values
 * (reference include/amici/defines.h:124-246).
 */
#ifndef ORACLE_REF_DRIVER_H
#define ORACLE_REF_DRIVER_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    int nx, nx_solver, np, nk, ny, nytrue, nz, nztrue, ne, ne_solver, nw, nspl;
    int is_ode; /* 1 = Model_ODE/CVODES, 0 = Model_DAE/IDAS */
} ref_dims;

typedef struct {
    double rtol, atol;           /* Solver::set_relative/absolute_tolerance          */
    double rtol_fsa, atol_fsa;   /* NaN = leave at default (= rtol/atol)              */
    double ss_tol_factor;        /* NaN = default (100)                               */
    double rtol_ss, atol_ss;     /* NaN = default                                     */
    double t0;                   /* Model::set_t0                                     */
    int sensi_order;             /* SensitivityOrder: 0 none, 1 first                 */
    int sensi_meth;              /* SensitivityMethod: 1 forward                      */
    int ism;                     /* InternalSensitivityMethod: 1 simultaneous, 2 stg  */
    int linsol;                  /* LinearSolver: 1 dense, 9 KLU                      */
    int maxsteps;                /* Solver::set_max_steps                             */
    int newton_maxsteps;         /* Solver::set_newton_max_steps                      */
    int ss_mode;                 /* SteadyStateComputationMode, -1 = default          */
    int ss_sensi_mode;           /* SteadyStateSensitivityMode, -1 = default          */
    int nmaxevent;               /* Model::set_n_max_event, -1 = default              */
    int stldet;                  /* Solver::set_stability_limit_flag, -1 = default    */
    int reporting;               /* RDataReporting: 0 full … 2 likelihood             */
    int num_threads;             /* OpenMP threads for run_simulations                */
    int failfast;
    int reinit_all;              /* ExpData::reinitialize_fixed_parameter_initial_states                          */
    int reinit_states_sim;       /* ExpData::reinitialization_state_idxs_sim as a bit set (bit i = state i)       */
    int reinit_states_presim;    /* ExpData::reinitialization_state_idxs_presim as a bit set                      */
    int x_nonneg;                /* Model::set_state_is_non_negative as a bit set (bit i = state i), 0 = off       */
    double maxtime;              /* Solver::set_max_time in seconds, 0 = unlimited                                 */
} ref_options;

typedef struct {
    /* ---- inputs ---- */
    int N, nt, nplist;
    const double* ts;        /* [nt]                                   */
    const int* plist;        /* [nplist]                               */
    const int* pscale;       /* [np] ParameterScaling                  */
    const double* p;         /* [N][np]   (scaled according to pscale) */
    const double* k;         /* [N][nk]                                */
    const double* k_preeq;   /* [N][nk] or NULL                        */
    const double* my;        /* [N][nt][nytrue] or NULL (NaN missing)  */
    const double* sigmay_in; /* [N][nt][nytrue] or NULL (NaN = model)  */
    const double* mz;        /* [N][nmaxevent][nztrue] or NULL         */
    const double* sigmaz_in; /* [N][nmaxevent][nztrue] or NULL         */
    const double* x0_in;     /* [N][nx] ExpData::x0 or NULL            */
    const double* sx0_in;    /* [N][nplist][nx] ExpData::sx0 or NULL   */
    const double* k_presim;  /* [N][nk] ExpData::fixed_parameters_presimulation or NULL */
    const double* t_presim;  /* [N]     ExpData::t_presim or NULL (no presimulation)    */
    /* ---- outputs (any may be NULL) ---- */
    int* status;             /* [N]                         */
    double* llh;             /* [N]                         */
    double* chi2;            /* [N]                         */
    double* sllh;            /* [N][nplist]                 */
    double* x;               /* [N][nt][nx]                 */
    double* sx;              /* [N][nt][nplist][nx]         */
    double* y;               /* [N][nt][ny]                 */
    double* sy;              /* [N][nt][nplist][ny]         */
    double* sigmay;          /* [N][nt][ny]                 */
    double* ssigmay;         /* [N][nt][nplist][ny]         */
    double* res;             /* [N][nt*nytrue]              */
    double* sres;            /* [N][nt*nytrue][nplist]      */
    double* FIM;             /* [N][nplist][nplist]         */
    double* x0;              /* [N][nx]                     */
    double* sx0;             /* [N][nplist][nx]             */
    double* xdot;            /* [N][nx_solver]              */
    double* J;               /* [N][nx_solver][nx_solver]   */
    double* z;               /* [N][nmaxevent][nz]          */
    double* sz;              /* [N][nmaxevent][nplist][nz]  */
    double* rz;              /* [N][nmaxevent][nz]          */
    double* srz;             /* [N][nmaxevent][nplist][nz]  */
    double* sigmaz;          /* [N][nmaxevent][nz]          */
    double* x_ss;            /* [N][nx]                     */
    double* sx_ss;           /* [N][nplist][nx]             */
    int* preeq_status;       /* [N][3]                      */
    int* preeq_numsteps;     /* [N][3]                      */
    double* preeq_wrms;      /* [N]                         */
    double* preeq_t;         /* [N]                         */
    int* posteq_status;      /* [N][3]                      */
    int* posteq_numsteps;    /* [N][3]                      */
    double* posteq_wrms;     /* [N]                         */
    double* posteq_t;        /* [N]                         */
    int* numsteps;           /* [N][nt]                     */
    int* numrhsevals;        /* [N][nt]                     */
    int* numerrtestfails;    /* [N][nt]                     */
    int* numnonlinsolvconvfails; /* [N][nt]                 */
    int* order;              /* [N][nt]                     */
    double* wall_seconds;    /* [1] steady_clock around run_simulations only */
} ref_batch;

/* dimensions of the linked model */
void ref_get_dims(ref_dims* out);
/* nominal parameters of the linked model: p[np] (unscaled), k[nk] */
void ref_get_nominal(double* p, double* k);
/* returns 0 on success; per-trajectory failures are reported in status[] */
int ref_run(const ref_options* opt, ref_batch* b);
int ref_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
