/* diffsol_b200.h -- C ABI of the B200-native batched implicit ODE/DAE integrator.
 *
 * This is the drop-in boundary for ONE hot path of martinjrobins/diffsol: the implicit step loop
 * (`Bdf::step`, `Sdirk::step`) driven by `solve_dense`, run over a BATCH of independent problem
 * instances, each with its own adaptive step size / order / Newton state.  The reference reaches
 * that path through Rust traits, not an FFI; the entry points below are what
 * a Rust shim crate (`impl OdeSolverMethod for BatchedBdf`, `impl LinearSolver<BatchMat>`) or the
 * reference's own C layer (crates/diffsol-c) would bind.  Conventions follow crates/diffsol-c:
 * opaque handles, every call returns an int (0 = OK, crates/diffsol-c/src/c_api_utils.rs:3-5),
 * message in a thread-local last-error string (crates/diffsol-c/src/error_c.rs:12-46), enums as int.
 *
 * Paths in comments are relative to /root/reference/crates/ (diffsol @ ad3477a).
 * All pointers are plain host or device pointers; there are no torch types in this ABI.
 */
#ifndef DIFFSOL_B200_H
#define DIFFSOL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- return codes (crates/diffsol-c/src/c_api_utils.rs:3-5) ------------------------------------ */
#define DSB_OK 0
#define DSB_ERR (-1)
#define DSB_BAD_ARG (-2)

/* ---- per-instance status: an instance that fails must not abort the batch.  Values mirror the
 * variants of OdeSolverError (diffsol/src/error.rs:40-93) the hot path can raise. ------------------ */
enum dsb_status {
    DSB_STATUS_OK = 0,
    DSB_STATUS_STEP_SIZE_TOO_SMALL = 1,            /* OdeSolverError::StepSizeTooSmall            bdf.rs:559-563 */
    DSB_STATUS_TOO_MANY_ERROR_TEST_FAILURES = 2,   /* OdeSolverError::TooManyErrorTestFailures   bdf.rs:1456-1464 */
    DSB_STATUS_TOO_MANY_NONLINEAR_FAILURES = 3,    /* OdeSolverError::TooManyNonlinearSolverFailures bdf.rs:1366-1375 */
    DSB_STATUS_STOP_TIME_BEFORE_CURRENT = 4,       /* OdeSolverError::StopTimeBeforeCurrentTime  bdf.rs:706-716 */
    DSB_STATUS_STOP_TIME_AT_CURRENT = 5,           /* OdeSolverError::StopTimeAtCurrentTime      bdf.rs:1593-1597 */
    DSB_STATUS_INITIAL_CONDITION_DID_NOT_CONVERGE = 6, /* NonLinearSolverError::InitialConditionDidNotConverge state.rs:152-154 */
    DSB_STATUS_LINESEARCH_FAILED = 7,
    DSB_STATUS_INTERPOLATION_TIME_AFTER_CURRENT = 8,
    DSB_STATUS_LU_SOLVE_FAILED = 9
};

/* ---- OdeSolverType (crates/diffsol-c: OdeSolverType{Bdf,Esdirk34,TrBdf2,Tsit45}) ------------------ */
enum dsb_method {
    DSB_METHOD_BDF = 0,        /* problem.bdf::<LS>()       ode_solver/problem.rs:649-655 */
    DSB_METHOD_TR_BDF2 = 1,    /* problem.tr_bdf2::<LS>()   ode_solver/problem.rs:320-328 */
    DSB_METHOD_ESDIRK34 = 2    /* problem.esdirk34::<LS>() */
};

/* ---- built-in equation sets (device functors; diffsol_b200/csrc/dsb_models.h) -------------------- */
enum dsb_model {
    DSB_EXP_DECAY = 0, DSB_EXP_DECAY_ALGEBRAIC = 1, DSB_ROBERTSON_DAE = 2, DSB_ROBERTSON_ODE = 3,
    DSB_ROBERTSON_ODE_G3 = 4, DSB_DYDT_Y2 = 5, DSB_GAUSSIAN_DECAY = 6, DSB_VAN_DER_POL = 7,
    DSB_VAN_DER_POL_SCALED = 8, DSB_HEAT1D_DAE_256 = 9, DSB_HEAT1D_DAE_32 = 10, DSB_SPM = 11, DSB_SPM99 = 12,
    DSB_EXP_DECAY_ROOT = 13, DSB_SPM_STOP = 14, DSB_SPM99_STOP = 15,
    DSB_HEAT1D_DAE_32_BC = 16, DSB_EXP_DECAY_RESET = 17, DSB_HEAT2D_10 = 18, DSB_BALL_BOUNCE = 19, DSB_EXP_DECAY_TWO_ROOTS = 20, DSB_SPM_CYCLE = 21,
    DSB_EXP_DECAY_ALGEBRAIC_RESET = 22
};

/* ---- statistics: one row of DSB_NSTATS int64 per instance.  Indices 0-9 are the fields of
 * OdeSolverStatistics (ode_solver/mod.rs:27-49), 10-12 the rhs OpStatistics (op/mod.rs:108-145). --- */
enum dsb_stat {
    DSB_STAT_LINEAR_SOLVER_SETUPS = 0,
    DSB_STAT_SETUPS_FROM_CHECKPOINT = 1,
    DSB_STAT_SETUPS_FROM_FIRST_CONVERGENCE_FAIL = 2,
    DSB_STAT_SETUPS_FROM_SECOND_CONVERGENCE_FAIL = 3,
    DSB_STAT_SETUPS_FROM_ERROR_TEST_FAIL = 4,
    DSB_STAT_SETUPS_FROM_STEP_SUCCESS = 5,
    DSB_STAT_STEPS = 6,
    DSB_STAT_ERROR_TEST_FAILURES = 7,
    DSB_STAT_NONLINEAR_SOLVER_ITERATIONS = 8,
    DSB_STAT_NONLINEAR_SOLVER_FAILS = 9,
    DSB_STAT_RHS_CALLS = 10,
    DSB_STAT_RHS_JAC_MULS = 11,
    DSB_STAT_RHS_MATRIX_EVALS = 12,
    DSB_NSTATS = 16
};

/* ---- OdeSolverOptions + InitialConditionSolverOptions (ode_solver/problem.rs:15-152) and the
 * BdfConfig / SdirkConfig growth clamps (ode_solver/config.rs:54-110); same names, same defaults. --- */
typedef struct dsb_options {
    int32_t max_nonlinear_solver_iterations;   /* 10 */
    int32_t max_error_test_failures;           /* 40, per step */
    int32_t max_nonlinear_solver_failures;     /* 50, cumulative */
    int32_t update_jacobian_after_steps;       /* 20 */
    int32_t update_rhs_jacobian_after_steps;   /* 50 */
    int32_t ic_max_linesearch_iterations;      /* 10 */
    int32_t ic_max_newton_iterations;          /* 10 */
    int32_t ic_max_linear_solver_setups;       /* 4 */
    int32_t ic_use_linesearch;                 /* 1 */
    int32_t reserved0;
    double nonlinear_solver_tolerance;         /* 0.2 */
    double min_timestep;                       /* 1e-13 */
    double max_timestep_growth;                /* 2.0 */
    double min_timestep_growth;                /* 2.0 */
    double max_timestep_shrink;                /* 0.9 */
    double min_timestep_shrink;                /* 0.5 */
    double threshold_to_update_jacobian;       /* 0.3 */
    double threshold_to_update_rhs_jacobian;   /* 0.2 */
    double pi_control_proportional;            /* 0.0 */
    double pi_control_integral;                /* 0.5 */
    double ic_step_reduction_factor;           /* 0.5 */
    double ic_armijo_constant;                 /* 1e-4 */
} dsb_options;

void dsb_options_default(dsb_options* opt);

const char* dsb_last_error(void);               /* thread-local, crates/diffsol-c/src/error_c.rs */
const char* dsb_version(void);
int dsb_device_count(int* count);               /* DSB_ERR when the CUDA runtime reports none */

/* ---- problem = what OdeBuilder::new()...build() returns (ode_solver/builder.rs:112-140,1447-1626) --
 * Defaults: t0 = 0, h0 = 1, rtol = 1e-6, atol = [1e-6] broadcast. */
typedef struct dsb_problem dsb_problem;
int dsb_problem_new(int model, dsb_problem** out);
int dsb_problem_free(dsb_problem* p);
int dsb_problem_dims(const dsb_problem* p, int32_t* nstates, int32_t* nparams, int32_t* has_mass);
/* Rows of every solve_dense column: the outputs of the equations' `out` function (OdeEquations::out; dense_write_out,
 * ode_solver/method.rs:822-848) when they have one, else nstates.  Wherever the result layouts below say nstates,
 * read nout. */
int dsb_problem_nout(const dsb_problem* p, int32_t* nout);
int dsb_problem_set_rtol(dsb_problem* p, double rtol);                      /* OdeBuilder::rtol */
int dsb_problem_set_atol(dsb_problem* p, const double* atol, int32_t n);    /* OdeBuilder::atol; n == 1 broadcasts */
int dsb_problem_set_t0(dsb_problem* p, double t0);                          /* OdeBuilder::t0 */
int dsb_problem_set_h0(dsb_problem* p, double h0);                          /* OdeBuilder::h0 */
int dsb_problem_set_use_coloring(dsb_problem* p, int32_t use_coloring);      /* OdeBuilder::use_coloring  builder.rs:1852-1857 */
int dsb_problem_set_options(dsb_problem* p, const dsb_options* opt);        /* OdeBuilder::ode_options / ic_options */
int dsb_problem_get_options(const dsb_problem* p, dsb_options* opt);
/* Forward sensitivities: OdeBuilder::sens_rtol / sens_atol (builder.rs:1466-1477, 1682-1716) and the choice of
 * `problem.bdf_sens::<LS>()` over `problem.bdf::<LS>()` (ode_solver/problem.rs:819-830).  enable != 0: the batch integrates
 * one sensitivity vector d y / d p_q per parameter beside the state (Bdf::sensitivity_solve, ode_solver/bdf.rs:934-989; Sdirk).
 * natol == 0 keeps the sensitivities out of the error test (OdeBuilder::turn_off_sensitivities_error_control); natol == 1
 * broadcasts sens_atol[0], natol == nstates gives it per state (param_scales = 1).  Equation sets qualify when they provide
 * sens_mul / init_sens (OdeEquationsImplicitSens), have no root / output / reset function and nstates <= 16 (ODEs, and
 * singular-mass DAEs, whose sensitivities are made consistent first: set_consistent_augmented, state.rs:167-238); every
 * method (Bdf::sensitivity_solve; Rk::do_stage_sdirk's sensitivity part, runge_kutta.rs:691-745).  Anything else: DSB_ERR
 * from the solve call. */
int dsb_problem_set_sensitivities(dsb_problem* p, int32_t enable, double sens_rtol, const double* sens_atol, int32_t natol);

/* ---- user equation sets: "user RHS closures and DiffSL-JIT modules drop in" -------------------------------------------------
 * The reference takes equations as Rust closures (builder.rs:192-200) or as a compiled DiffSL module consumed through its
 * symbol table (crates/diffsol/src/ode_equations/diffsl.rs:1072-1098, 1221-1232; the table:
 * crates/diffsol-c/tests/external-dynamic-logistic/src/lib.rs:15-600).  Neither can run inside a kernel as it is, so the
 * equations cross this boundary as SOURCE TEXT that nvcc compiles for sm_100a at run time into one instantiation of the
 * library's kernel families:
 *   DSB_MODEL_SOURCE_DIFFSL   the DiffSL symbol table as C: set_u0 / rhs / rhs_grad / set_inputs [/ mass / calc_out /
 *                             calc_stop], every definition prefixed with DSB_SYMBOL, dimensions as #define
 *                             DSB_DIFFSL_STATES / _INPUTS / _DATA / _OUTPUTS / _STOP / _HAS_MASS (csrc/dsb_diffsl_adapter.h)
 *   DSB_MODEL_SOURCE_FUNCTOR  a struct with the closure signatures (x, p, t, y), (x, p, t, v, y), (x, p, t, beta, y), (p, t, y)
 *                             of builder.rs:192-200 as DSB_HD static member functions (csrc/dsb_models.h shows 22 of them)
 * dsb_model_library_build writes a shared object; dsb_model_library_load registers it and returns a model id that
 * dsb_problem_new accepts like a built-in one.  `csrc_dir` = the directory of this library's kernel sources
 * (diffsol_b200/csrc).  Build needs nvcc on PATH (or $NVCC); loading a built library does not. */
enum { DSB_MODEL_SOURCE_FUNCTOR = 0, DSB_MODEL_SOURCE_DIFFSL = 1 };
#define DSB_MODEL_PLUGIN_ID0 1000
int dsb_model_library_build(const char* source_path, int32_t kind, const char* struct_name, const char* csrc_dir,
                            const char* out_path);
int dsb_model_library_load(const char* library_path, int32_t* model_out);

/* ---- batch = Context::nbatch (diffsol-la/src/context/mod.rs:27) + per-instance parameters + all
 * device state of the batched solver, resident on ONE GPU.  Not thread-safe; all work is enqueued on
 * the stream given to the solve call; only the get_* / *_host calls synchronise. */
typedef struct dsb_batch dsb_batch;
int dsb_batch_new(const dsb_problem* p, int64_t nbatch, int32_t device, dsb_batch** out);
int dsb_batch_free(dsb_batch* b);
int64_t dsb_batch_size(const dsb_batch* b);

/* Execution model of the integrator kernels: 0 = automatic (one thread per instance for n <= 16; above, for banded
 * models (ODE or singular-mass DAE), one WARP per instance with the state in shared memory for BDF and one thread per
 * instance with the state in global memory for (E)SDIRK and for equations with a reset function; else one thread
 * block per instance), 1 = thread per instance, 2 = block per instance, 3 = banded thread per instance, 4 = banded
 * warp per instance.  Results are identical. */
int dsb_batch_set_execution(dsb_batch* b, int32_t mode);

/* Parameters, instance-major: params[b*nparams + j], exactly how the reference concatenates batched
 * parameters (diffsol/src/ode_equations/test_models/exponential_decay.rs:297-304). */
int dsb_batch_set_params_host(dsb_batch* b, const double* params, int64_t nbatch, int32_t nparams);
int dsb_batch_set_params_device(dsb_batch* b, const double* params_dev, int64_t nbatch, int32_t nparams, void* stream);

/* `problem.<method>::<LS>()?.solve_dense(t_eval)` for every instance (ode_solver/method.rs:467-505,
 * 721-848; Bdf::step ode_solver/bdf.rs:1277-1589; Sdirk::step ode_solver/sdirk.rs:409-543).
 * t_eval is a HOST array of nt increasing times shared by the batch; t_eval[nt-1] is the stop time.
 * ys_dev: DEVICE buffer of nt*nstates*nbatch doubles, batch-major: ys[(k*nstates + i)*nbatch + b].
 * Asynchronous on `stream` (a cudaStream_t, NULL = default stream). */
int dsb_batch_solve_dense(dsb_batch* b, int32_t method, const double* t_eval, int32_t nt, double* ys_dev, void* stream);

/* The user-level stepping loop of the reference's own tests (ode_solver/mod.rs:132-141), for every instance:
 *   for each k: while |solver.state().t| < |t_points[k]| { solver.step()? }; ys[k] = solver.interpolate(t_points[k])?
 * i.e. OdeSolverMethod::step (method.rs:99) + interpolate (method.rs:106) with NO stop time.  Same
 * layouts as dsb_batch_solve_dense.  This is the call that reproduces the reference's statistics snapshots. */
int dsb_batch_step_and_interpolate(dsb_batch* b, int32_t method, const double* t_points, int32_t npts, double* ys_dev, void* stream);

/* `problem.bdf_sens::<LS>()?.solve_dense_sensitivities(t_eval)` for every instance (ode_solver/sensitivities.rs:114-262,
 * dense_write_out_sensitivities :360-397): the states as dsb_batch_solve_dense writes them, and
 * sens_dev: DEVICE buffer of nt*nparams*nstates*nbatch doubles, sens[((k*nparams + q)*nstates + i)*nbatch + b] =
 * d y_i / d p_q of instance b at t_eval[k].  The problem must have sensitivities enabled (dsb_problem_set_sensitivities).
 * dsb_batch_step_and_interpolate_sensitivities is the free-running loop of the reference's tests with interpolate_sens
 * at every point (test_ode_solver(.., solve_for_sensitivities = true), ode_solver/mod.rs:104-194). */
int dsb_batch_solve_dense_sensitivities(dsb_batch* b, int32_t method, const double* t_eval, int32_t nt, double* ys_dev, double* sens_dev,
                                        void* stream);
int dsb_batch_step_and_interpolate_sensitivities(dsb_batch* b, int32_t method, const double* t_points, int32_t npts, double* ys_dev,
                                                 double* sens_dev, void* stream);
/* `problem.<method>::<LS>()?.solve(final_time)` for every instance (OdeSolverMethod::solve, ode_solver/method.rs:227-258; fn solve
 * :881-961, write_out :965-1000): one column per INTERNAL step -- (state.t, state.y), or out(state.y, state.t) for equations
 * with an output function -- after the initial one; a root ends an instance's solve with the state at the root in its last
 * column.  Every instance takes its own number of steps, so the result is ragged and comes in two passes:
 *   dsb_batch_solve_count   integrates once, writes nothing, and leaves per-instance column counts (dsb_batch_get_root_info's
 *                           ncols) and their exclusive prefix sums in the batch; *total_columns = their sum.  Synchronous.
 *   dsb_batch_solve_offsets copies the nbatch + 1 offsets to the host.
 *   dsb_batch_solve_write   integrates again (the integration is deterministic: same steps) and writes instance b's column k
 *                           at ts_dev[off[b] + k] and ys_dev[(off[b] + k) * nout + i]; asynchronous on `stream`.  Must follow a
 *                           count with the same method and final_time (and unchanged parameters).
 *   dsb_batch_solve_write_host  the same into host arrays of total_columns (x nout) doubles, synchronous.
 * Parameters are set beforehand (dsb_batch_set_params_host / _device).  Built for the thread-per-instance kernels: n <= 16,
 * BDF / TR-BDF2 / ESDIRK34, equations without a reset function. */
int dsb_batch_solve_count(dsb_batch* b, int32_t method, double final_time, int64_t* total_columns);
int dsb_batch_solve_offsets(dsb_batch* b, int64_t* offsets_host);
int dsb_batch_solve_write(dsb_batch* b, int32_t method, double final_time, double* ts_dev, double* ys_dev, void* stream);
int dsb_batch_solve_write_host(dsb_batch* b, int32_t method, double final_time, double* ts_host, double* ys_host);

/* The same two calls with HOST buffers (parameters in, results back, synchronised): ys_host instance-major
 * [nbatch][nt][nstates]; sens_host [nbatch][nt][nparams][nstates] -- instance b's block holds, for every time, the nparams
 * sensitivity vectors one after the other (the Vec of matrices solve_dense_sensitivities returns, interleaved by time). */
int dsb_batch_solve_dense_sensitivities_host(dsb_batch* b, int32_t method, const double* params_host, int32_t nparams, const double* t_eval,
                                             int32_t nt, double* ys_host, double* sens_host, int64_t* stats_host, int32_t* status_host);
int dsb_batch_step_and_interpolate_sensitivities_host(dsb_batch* b, int32_t method, const double* params_host, int32_t nparams,
                                                      const double* t_points, int32_t npts, double* ys_host, double* sens_host,
                                                      int64_t* stats_host, int32_t* status_host);

/* Same call with HOST buffers: copies parameters in, runs, copies results back, synchronises.
 * ys_host layout is instance-major [nbatch][nt][nstates] (each instance's block is the column-major
 * nstates x nt matrix `solve_dense` returns).  stats_host ([nbatch][DSB_NSTATS]) and status_host
 * ([nbatch]) may be NULL. */
int dsb_batch_solve_dense_host(dsb_batch* b, int32_t method, const double* params_host, int32_t nparams,
                               const double* t_eval, int32_t nt,
                               double* ys_host, int64_t* stats_host, int32_t* status_host);

int dsb_batch_step_and_interpolate_host(dsb_batch* b, int32_t method, const double* params_host, int32_t nparams,
                                        const double* t_points, int32_t npts,
                                        double* ys_host, int64_t* stats_host, int32_t* status_host);

/* Per-instance results of the last solve (device -> host copy, synchronises). */
int dsb_batch_get_stats(dsb_batch* b, int64_t* stats_host /* [nbatch][DSB_NSTATS] */);
int dsb_batch_get_status(dsb_batch* b, int32_t* status_host /* [nbatch] */);
/* The same statistics written to a DEVICE buffer [nbatch][DSB_NSTATS] int64, asynchronously on `stream`. */
int dsb_batch_get_stats_device(dsb_batch* b, int64_t* stats_dev, void* stream);
int dsb_batch_get_final_state(dsb_batch* b, double* t_host, double* h_host, int32_t* order_host /* each [nbatch] or NULL */);
/* Events (OdeSolverStopReason::RootFound, ode_solver/method.rs:774-805, 493-503): for equations with root functions
 * an instance stops at its first root.  root_idx[b] = index of that root function, -1 when the instance ran to the
 * last t_eval point; ncols[b] = solve_dense columns written: the points up to the root, then the state AT the root
 * (its time is the final state's t); the columns behind stay NaN.  Each [nbatch] or NULL. */
int dsb_batch_get_root_info(dsb_batch* b, int32_t* root_idx_host, int32_t* ncols_host);
/* Device views (valid until the next solve / free): stats [DSB_NSTATS][nbatch] int32, status [nbatch] int32. */
int dsb_batch_device_views(dsb_batch* b, const int32_t** stats_dev, const int32_t** status_dev);

/* Sum over the batch of one statistic of the last solve (device reduction; synchronises). */
int dsb_batch_sum_stat(dsb_batch* b, int32_t stat, int64_t* total);

/* Device time of the last solve's kernels in milliseconds (CUDA events on the solve's stream). */
int dsb_batch_last_kernel_ms(dsb_batch* b, float* ms);
/* Device time of the integrator kernel alone (without the initialisation kernel). */
int dsb_batch_last_integrator_ms(dsb_batch* b, float* ms);
/* Number of kernels this library launched for the last solve. */
int dsb_batch_last_launch_count(dsb_batch* b, int32_t* launches);
/* Diagnostics: the 31 device words behind the persistent kernels' work counter.  A library built with
 * -DDSB_LANE_PROFILE accumulates warp-scheduler occupancy counters there (tools/lane_profile.py); otherwise zeros. */
int dsb_batch_debug_words(dsb_batch* b, uint64_t* words_host /* [31] */);

/* ---- the LinearSolver<M> pair a Rust `impl LinearSolver<BatchMat>` would call
 * (diffsol-la/src/linear_solver/mod.rs:19-42; replaces NalgebraLU nalgebra/lu.rs:31-51 and the
 * per-instance cuSOLVER loop of linear_solver/cuda/lu.rs:80-95,127-145).
 * a_dev: nbatch column-major n x n matrices, batch-major: a[(j*n + i)*nbatch + b]; overwritten by LU.
 * piv_dev: [n][nbatch] int32 row swaps (row i swapped with piv[i]); info_dev[b] != 0 => zero pivot. */
int dsb_lu_factor_batched(double* a_dev, int32_t n, int64_t nbatch, int32_t* piv_dev, int32_t* info_dev, void* stream);
int dsb_lu_solve_batched(const double* lu_dev, const int32_t* piv_dev, double* b_dev, int32_t n, int64_t nbatch,
                         int32_t* info_dev, void* stream);

/* The same pair for INSTANCE-major storage, the layout of the reference's own CUDA matrices
 * (diffsol-la/src/matrix/cuda.rs: instance b's column-major n x n block at a + b*n*n; vectors at b*n,
 * diffsol-la/src/vector/cuda.rs:119-130): one thread block per instance, panels of 32 columns staged
 * through shared memory, n <= 512.  piv_dev: [nbatch][n]; b_dev: [nbatch][n]. */
int dsb_lu_factor_instance_major(double* a_dev, int32_t n, int64_t nbatch, int32_t* piv_dev, int32_t* info_dev, void* stream);
int dsb_lu_solve_instance_major(const double* lu_dev, const int32_t* piv_dev, double* b_dev, int32_t n, int64_t nbatch,
                                int32_t* info_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DIFFSOL_B200_H */
