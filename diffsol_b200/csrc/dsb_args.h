// dsb_args.h -- plain-old-data blocks handed to the sm_100a kernels by value (constant bank).
//
// `DsbProblemArgs` is the device-side image of what `OdeBuilder::build()` returns in the reference
// (crates/diffsol/src/ode_solver/builder.rs:112-140, problem.rs:98-193): tolerances, t0, h0, the
// solver options, plus the order-dependent BDF constant tables (bdf.rs:253-276, 433-463) which the
// host evaluates once with the same IEEE operations the oracle uses.
#pragma once
#include <stdint.h>

#include "../../include/diffsol_b200.h"

#define DSB_MAX_STATES 64      // register/local-memory lane kernels; larger n uses the block-cooperative path
#define DSB_MAX_ORDER 5        // bdf_state.rs:44
#define DSB_NDIFF (DSB_MAX_ORDER + 3)
#define DSB_DEFAULT_NEWTON_PASSES 3   // Newton iterations a warp may run per trip of its state machine (dsb_bdf_kernel.cuh, NEWTON block)
#define DSB_DEFAULT_QUORUM 16     // lanes of a warp that make a heavy block worth running (see dsb_bdf_kernel.cuh)
#define DSB_LANE_THREADS 128   // block size of the one-thread-per-instance kernels

// Lane-kernel statistics are int32 in device memory, one array per counter (batch-major).
enum dsb_solver_state {        // ode_solver/jacobian_update.rs:3-10
    DSB_STEP_SUCCESS = 0, DSB_FIRST_CONVERGENCE_FAIL = 1, DSB_SECOND_CONVERGENCE_FAIL = 2,
    DSB_ERROR_TEST_FAIL = 3, DSB_CHECKPOINT = 4
};

struct DsbBdfTables {
    double alpha[DSB_MAX_ORDER + 1];
    double gamma[DSB_MAX_ORDER + 1];
    double error_const2[DSB_MAX_ORDER + 1];
    // U = R(order, 1) for order = 1..5, each (order+1)^2 column-major, compact (leading dimension order+1)
    double u[DSB_MAX_ORDER + 1][36];
    double eta_reset;              // 20^1.25   convergence.rs:36-38
    double eta_reset_timestep;     // 100^1.25  convergence.rs:40-42
    double ic_steptol;             // eps^(2/3) line_search.rs:126
    // Quotients of the step loop whose operands only take a handful of values, formed once on the host with the same
    // IEEE division the reference executes at every use (the kernels then read them instead of dividing):
    double inv_int[32];            // 1.0 / k: the exponent 1 / (niter - 1) of the convergence rate (convergence.rs:77) and
                                   // RN(1 / i) for the divisions by the row index in Bdf::_compute_r (bdf.rs:433-463)
    double safety[32];             // 0.9 (2 m + 1) / (2 m + niter), m = max_nonlinear_solver_iterations (bdf.rs:1351-1353)
    double pi_ki[DSB_MAX_ORDER + 2], pi_kp[DSB_MAX_ORDER + 2];   // pi_control_* / order (runge_kutta.rs:1313-1335)
};

// Butcher tableau of an (E)SDIRK method (ode_solver/tableau.rs:41-159), evaluated on the host
struct DsbSdirkTableau {
    int32_t s, order, has_beta, reserved;
    double a[16];        // s x s column-major: a(i, j) = a[j * s + i]
    double b[4], c[4], d[4];
    double beta[8];      // s x 2 column-major (dense-output coefficients), when has_beta
};

struct DsbProblemArgs {
    int64_t nbatch;
    int32_t nt;
    int32_t use_coloring;
    int32_t free_running;      // 1: no stop time; step until t passes each point, then interpolate (the
                               // `while t < t_k { step() }; interpolate(t_k)` loop of ode_solver/mod.rs:132-141)
    int32_t quorum;            // warp scheduler: lanes that make a heavy block worth running (dsb_bdf_kernel.cuh)
    int32_t coop_dense_only, reserved1;   // 1: the block-per-instance path always uses the blocked dense LU (test hook);
                                          // reserved1 != 0: the warp-per-instance banded kernel redoes every solve through its exact path (test hook)
    int32_t ncolors;
    int32_t color_of_col[DSB_MAX_STATES];      // colour index of every column
    uint64_t nz_rows_of_col[DSB_MAX_STATES];   // bit i set <=> (i, col) is in the sparsity pattern
    double rtol;
    double atol[DSB_MAX_STATES];
    double t0, h0;
    dsb_options opt;
    DsbBdfTables tab;
    DsbSdirkTableau rk;
    // forward sensitivities (problem.bdf_sens(), ode_solver/problem.rs:819-830; builder.rs:1682-1716): sens != 0 integrates
    // one sensitivity vector per parameter; sens_error_control puts them into the error test with sens_rtol / sens_atol
    int32_t ragged;                // solve(final_time) form (DsbRagged<M> kernels): 1 = count the columns, 2 = write them
    int32_t newton_passes;         // lane kernels: Newton iterations a warp may run per trip of its state machine (< 1 reads as 1)
    int32_t sens, sens_error_control;
    double sens_rtol;
    double sens_atol[DSB_MAX_STATES];
};

// Device buffers of one batch, all batch-major (instance index fastest) so that a warp's 32 lanes
// touch 32 consecutive doubles.
struct DsbBatchBuffers {
    const double* params;    // [np][B]
    const double* t_eval;    // [nt]
    double* y0;              // [n][B]   state after `new_and_consistent` (state.rs:969-997)
    double* dy0;             // [n][B]
    double* h0;              // [B]
    double* ys;              // [nt][n][B]  solve_dense output
    int32_t* stats;          // [DSB_NSTATS][B]
    int32_t* status;         // [B]
    double* fin_t;           // [B]
    double* fin_h;           // [B]
    int32_t* fin_order;      // [B]
    int32_t* root_idx;       // [B]   index of the root function that stopped the instance, -1: none (OdeSolverStopReason::RootFound)
    int32_t* ncols;          // [B]   solve_dense columns written (nt unless a root or an error stopped the instance)
    double* ss;              // [nt][np][n][B]  solve_dense_sensitivities output (NULL without sensitivities)
    // solve(final_time), writing pass: instance b's column k at rag_ts[rag_off[b] + k], rag_ys[(rag_off[b] + k) * nout + i]
    const int64_t* rag_off;  // [B + 1]
    double* rag_ts;
    double* rag_ys;
    // sensitivity BDF lane kernel built with DSB_SENS_SDIFF_GLOBAL: the difference arrays of the sensitivities, one word of
    // every RESIDENT lane side by side ([np * 8 * n][grid * threads]); set by the launcher
    double* sens_ws;
};
