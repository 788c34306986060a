// dsb_lane.cuh -- per-instance ("lane") building blocks of the batched implicit step loop.
//
// One CUDA thread owns one ODE instance; every vector of the instance lives in registers (all loops
// are unrolled over the compile-time state count N, every array index is static after unrolling).
// Each block below restates one piece of the reference with the SAME floating-point expression
// order as the CPU path, so that a build with --fmad=false reproduces the reference controller's
// decisions bit for bit (paths relative to /root/reference/crates):
//   lane_squared_norm   diffsol-la/src/vector/nalgebra_serial.rs:395-408
//   LaneLU              nalgebra 0.35 `DMatrix::lu()` / `LU::solve_mut` as called from
//                       diffsol-la/src/linear_solver/nalgebra/lu.rs:31-51
//   LaneConvergence     diffsol-nl/src/convergence.rs:7-140
//   LaneJacobianUpdate  diffsol/src/ode_solver/jacobian_update.rs
//   lane_jacobian       diffsol/src/op/closure.rs:140-147, jacobian/mod.rs:236-256,
//                       op/nonlinear_op.rs:211-220
//   lane_mass_matrix    diffsol/src/op/linear_op.rs:42-51
#pragma once
#include "dsb_args.h"
#include "dsb_math.h"

#define DSB_DEV __device__ __forceinline__

// Compile-time loop: f(std::integral_constant<int, I>) for I in [I0, I1).  Used wherever a loop index
// is compared for EQUALITY with a run-time value to guard a register-array access: with an ordinary
// `#pragma unroll` loop the optimiser propagates the equality into the index *before* unrolling and
// turns the access into a run-time-indexed (local-memory) one.
template <int I> struct dsb_int { static constexpr int value = I; };
template <int I0, int I1, class F>
DSB_DEV void dsb_static_for(F&& f) {
    if constexpr (I0 < I1) {
        f(dsb_int<I0>{});
        dsb_static_for<I0 + 1, I1>(f);
    }
}
// same, descending: I1-1 down to I0
template <int I0, int I1, class F>
DSB_DEV void dsb_static_for_down(F&& f) {
    if constexpr (I0 < I1) {
        f(dsb_int<I1 - 1>{});
        dsb_static_for_down<I0, I1 - 1>(f);
    }
}

// Block shape of the integrator kernels.  ONE persistent block per SM (they never synchronise across warps, so the
// block size is free): as many lanes as shared memory (WORDS doubles per lane) and the register file allow.
// Registers are split per scheduler (4 x 16384), so what counts is warps per scheduler: 4 at <= 128 registers,
// 3 at <= 168, 2 at <= 255.  Resident warps are what hides the FP64 dependency latency: 12 -> 15 warps per SM on
// the Robertson sweep was worth 1.12x (profiles/r1_v10_*).
template <int WORDS, int N>
struct LaneBlockShape {
    static constexpr int T_SMEM = (226 * 1024 / (WORDS * 8)) / 32 * 32;
    static constexpr int T_REG = N <= 4 ? 512 : N <= 6 ? 384 : 256;
#ifdef DSB_THREADS
    static constexpr int THREADS = DSB_THREADS;                     // tuning experiments
#else
    static constexpr int THREADS = T_SMEM < 32 ? 32 : (T_SMEM < T_REG ? T_SMEM : T_REG);
#endif
    static constexpr int WARPS_PER_SCHEDULER = (THREADS / 32 + 3) / 4;
    static constexpr int MAXNREG = (512 / WARPS_PER_SCHEDULER) / 8 * 8 > 255 ? 255 : (512 / WARPS_PER_SCHEDULER) / 8 * 8;
};

template <int N>
DSB_DEV double lane_squared_norm(const double (&x)[N], const double (&y)[N], const double* __restrict__ atol, double rtol) {
#define DSB_DIV(a, b) ((a) / (b))
    double acc = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const double term = DSB_DIV(x[i], dsb_abs(y[i]) * rtol + atol[i]);
        acc += term * term;
    }
    return DSB_DIV(acc, (double)N);
#undef DSB_DIV
}

// ---- dense LU with partial pivoting, column-major a[col][row] -------------------------------------
template <int N, class Div = DsbDivInline>
struct LaneLU {
    double a[N][N];
    int piv[N];          // row i was swapped with row piv[i] (piv[i] == i: no swap)
    double rinv[N];      // RN(1 / U_ii) as dsb_rcp() would return it (NaN outside its proven range); filled by factor()

    // a must already hold the matrix to factor.
    DSB_DEV void factor() {
        dsb_static_for<0, N>([&](auto I_) {
            constexpr int i = decltype(I_)::value;
            int p = i;
            double the_max = dsb_abs(a[i][i]);
            dsb_static_for<i + 1, N>([&](auto R_) {
                constexpr int r = decltype(R_)::value;
                const double val = dsb_abs(a[i][r]);
                if (val > the_max) { the_max = val; p = r; }
            });
            double diag = a[i][i];
            dsb_static_for<i + 1, N>([&](auto R_) {
                constexpr int r = decltype(R_)::value;
                diag = (p == r) ? a[i][r] : diag;
            });
            piv[i] = (diag == 0.0) ? i : p;
            rinv[i] = 0.0;
            if (diag != 0.0) {                              // else: no non-zero entries on this column
                // row swap i <-> p written as selects so that every register index stays static
                dsb_static_for<i + 1, N>([&](auto R_) {
                    constexpr int r = decltype(R_)::value;
                    const bool sw = (p == r);
                    dsb_static_for<0, N>([&](auto C_) {
                        constexpr int c = decltype(C_)::value;
                        const double ai = a[c][i], ar = a[c][r];
                        a[c][i] = sw ? ar : ai;
                        a[c][r] = sw ? ai : ar;
                    });
                });
                const double inv_diag = 1.0 / diag;
                rinv[i] = dsb_rcp_from(diag, inv_diag);
#pragma unroll
                for (int r = i + 1; r < N; ++r) a[i][r] *= inv_diag;
#pragma unroll
                for (int k = i + 1; k < N; ++k) {
                    const double mpk = -a[k][i];
#pragma unroll
                    for (int r = i + 1; r < N; ++r) a[k][r] = mpk * a[i][r] + a[k][r];
                }
            }
        });
    }

    // false <=> zero on U's diagonal (LaError::LuSolveFailed).  RCP: divide through the stored reciprocals (rinv must hold
    // what factor() left there): the same quotients in 5 operations each (dsb_math.h: dsb_div_rcp)
    template <bool RCP = false>
    DSB_DEV bool solve(double (&b)[N]) const {
        dsb_static_for<0, N>([&](auto I_) {
            constexpr int i = decltype(I_)::value;
            dsb_static_for<i + 1, N>([&](auto R_) {
                constexpr int r = decltype(R_)::value;
                const bool sw = (piv[i] == r);
                const double bi = b[i], br = b[r];
                b[i] = sw ? br : bi;
                b[r] = sw ? bi : br;
            });
        });
#pragma unroll
        for (int i = 0; i + 1 < N; ++i) {
            const double mc = -(b[i] / 1.0);
#pragma unroll
            for (int r = i + 1; r < N; ++r) b[r] = mc * a[i][r] + b[r];
        }
        bool ok = true;
#pragma unroll
        for (int i = N - 1; i >= 0; --i) {
            const double diag = a[i][i];
            if (diag == 0.0) ok = false;
            if (ok) {
                const double coeff = RCP ? dsb_div_rcp(b[i], diag, rinv[i]) : Div::div(b[i], diag);
                b[i] = coeff;
                const double mc = -coeff;
#pragma unroll
                for (int r = 0; r < i; ++r) b[r] = mc * a[i][r] + b[r];
            }
        }
        return ok;
    }
};

// ---- Convergence ------------------------------------------------------------------------------------
enum { LANE_CONVERGED = 0, LANE_DIVERGED = 1, LANE_CONTINUE = 2 };

struct LaneConvergence {
    double tol, eta, old_norm;
    int max_iter, niter;
    bool has_old_norm;

    DSB_DEV void reset() { niter = 0; has_old_norm = false; }
    // convergence.rs:68-131
    DSB_DEV int check_norm(double norm) {
        niter += 1;
        if (has_old_norm) {
            const double rate = dsb_pow(norm / old_norm, 1.0 / (double)(niter - 1));
            if (rate > 0.9) return LANE_DIVERGED;
            if (dsb_powi(rate, max_iter - niter) / (1.0 - rate) * norm > tol) return LANE_DIVERGED;
            eta = rate / (1.0 - rate);
        } else {
            const double min_eta = 1e4 * 2.220446049250313e-16;
            if (eta < min_eta) eta = min_eta;
            eta = dsb_pow(eta, 0.8);
        }
        if (eta * norm < tol) return LANE_CONVERGED;
        return LANE_CONTINUE;
    }
    // convergence.rs:133-139 -- old_norm is frozen at the FIRST iteration's norm
    DSB_DEV int check_new_iteration(double norm) {
        const int s = check_norm(norm);
        if (niter == 1) { has_old_norm = true; old_norm = norm; }
        return s;
    }
};

// ---- JacobianUpdate ---------------------------------------------------------------------------------
struct LaneJacobianUpdate {
    int steps_since_jacobian_eval, steps_since_rhs_jacobian_eval;
    double h_at_last_jacobian_update;

    DSB_DEV void init(double h_at_last) {
        steps_since_jacobian_eval = 0; steps_since_rhs_jacobian_eval = 0; h_at_last_jacobian_update = h_at_last;
    }
    DSB_DEV void update_jacobian(double h) { steps_since_jacobian_eval = 0; h_at_last_jacobian_update = h; }
    DSB_DEV void update_rhs_jacobian(double h) {
        steps_since_rhs_jacobian_eval = 0; steps_since_jacobian_eval = 0; h_at_last_jacobian_update = h;
    }
    DSB_DEV void step() { ++steps_since_jacobian_eval; ++steps_since_rhs_jacobian_eval; }
    template <class Div = DsbDivInline>
    DSB_DEV bool check_jacobian_update(const dsb_options& o, double h, int s) const {
        if (s == DSB_STEP_SUCCESS)
            return steps_since_jacobian_eval >= o.update_jacobian_after_steps
                   || dsb_abs(Div::div(h, h_at_last_jacobian_update) - 1.0) > o.threshold_to_update_jacobian;
        return true;
    }
    template <class Div = DsbDivInline>
    DSB_DEV bool check_rhs_jacobian_update(const dsb_options& o, double h, int s) const {
        if (s == DSB_STEP_SUCCESS) return steps_since_rhs_jacobian_eval >= o.update_rhs_jacobian_after_steps;
        if (s == DSB_FIRST_CONVERGENCE_FAIL)
            return dsb_abs(Div::div(h, h_at_last_jacobian_update) - 1.0) < o.threshold_to_update_rhs_jacobian;
        if (s == DSB_SECOND_CONVERGENCE_FAIL) return steps_since_rhs_jacobian_eval > 0;
        if (s == DSB_ERROR_TEST_FAIL) return false;
        return true;   // Checkpoint
    }
};

// ---- per-instance statistics (ode_solver/mod.rs:27-69, op/mod.rs:108-145) -----------------------------
template <class V>
DSB_DEV void lane_record_linear_solver_setup(V& v, int s) {
    v[DSB_STAT_LINEAR_SOLVER_SETUPS] += 1;
    if (s == DSB_CHECKPOINT) v[DSB_STAT_SETUPS_FROM_CHECKPOINT] += 1;
    else if (s == DSB_FIRST_CONVERGENCE_FAIL) v[DSB_STAT_SETUPS_FROM_FIRST_CONVERGENCE_FAIL] += 1;
    else if (s == DSB_SECOND_CONVERGENCE_FAIL) v[DSB_STAT_SETUPS_FROM_SECOND_CONVERGENCE_FAIL] += 1;
    else if (s == DSB_ERROR_TEST_FAIL) v[DSB_STAT_SETUPS_FROM_ERROR_TEST_FAIL] += 1;
    else v[DSB_STAT_SETUPS_FROM_STEP_SUCCESS] += 1;
}
struct LaneStats {                 // counters in registers
    int v[DSB_NSTATS];
    DSB_DEV void clear() {
#pragma unroll
        for (int i = 0; i < DSB_NSTATS; ++i) v[i] = 0;
    }
    DSB_DEV void record_linear_solver_setup(int s) { lane_record_linear_solver_setup(v, s); }
};
// The same counters in the lane's shared-memory column, two int32 per 64-bit word (word stride STRIDE2 / 2
// doubles): frees 16 registers per lane, which is what limits the resident warps of the integrator kernels.
template <int STRIDE2>
struct SmemIntColumn {
    int* base;
    DSB_DEV int& operator[](int k) const { return base[(k >> 1) * STRIDE2 + (k & 1)]; }
};
template <int STRIDE2>
struct SmemLaneStats {
    SmemIntColumn<STRIDE2> v;
    DSB_DEV void record_linear_solver_setup(int s) { lane_record_linear_solver_setup(v, s); }
};

// ---- df/dy assembly: J[col][row] ----------------------------------------------------------------------
template <class M>
DSB_DEV void lane_jacobian(const DsbProblemArgs& pa, const double (&x)[M::N], const double* p, double t,
                           double (&J)[M::N][M::N], LaneStats& st) {
    constexpr int N = M::N;
    st.v[DSB_STAT_RHS_MATRIX_EVALS] += 1;
    double v[N], col[N];
#pragma unroll
    for (int i = 0; i < N; ++i) { v[i] = 0.0; col[i] = 0.0; }
    if (pa.use_coloring) {
#pragma unroll
        for (int j = 0; j < N; ++j)
#pragma unroll
            for (int i = 0; i < N; ++i) J[j][i] = 0.0;
        for (int c = 0; c < pa.ncolors; ++c) {
#pragma unroll
            for (int j = 0; j < N; ++j) if (pa.color_of_col[j] == c && pa.nz_rows_of_col[j] != 0) v[j] = 1.0;
            M::jac_mul(x, p, t, v, col);
            st.v[DSB_STAT_RHS_JAC_MULS] += 1;
#pragma unroll
            for (int j = 0; j < N; ++j) {
                if (pa.color_of_col[j] == c) {
                    const uint64_t nz = pa.nz_rows_of_col[j];
#pragma unroll
                    for (int i = 0; i < N; ++i) if ((nz >> i) & 1ull) J[j][i] = col[i];
                }
                v[j] = 0.0;
            }
        }
    } else {
#pragma unroll
        for (int j = 0; j < N; ++j) {
            v[j] = 1.0;
            M::jac_mul(x, p, t, v, col);
            st.v[DSB_STAT_RHS_JAC_MULS] += 1;
#pragma unroll
            for (int i = 0; i < N; ++i) J[j][i] = col[i];
            v[j] = 0.0;
        }
    }
}

// The same assembly for the integrator kernels, written for code size (their loop bodies are instruction-cache
// bound): ONE seed/jac_mul/scatter sequence inside a rolled loop over the colours, results handed to
// `store(col, row, value)` (shared memory, so the run-time column index costs nothing).  Without colouring the host
// fills the colour tables with one colour per column and a full pattern (dsb_capi.cu:fill_problem_args), which makes
// this loop the dense column-by-column assembly of op/nonlinear_op.rs:211-220, same values, same call counts.
template <class M, class Stats, class Store>
DSB_DEV void lane_jacobian_to(const DsbProblemArgs& pa, const double (&x)[M::N], const double* p, double t,
                              Stats& st, Store&& store) {
    constexpr int N = M::N;
    st.v[DSB_STAT_RHS_MATRIX_EVALS] += 1;
    double v[N], col[N];
#pragma unroll
    for (int i = 0; i < N; ++i) col[i] = 0.0;
#pragma unroll 1
    for (int j = 0; j < N; ++j)
#pragma unroll
        for (int i = 0; i < N; ++i) store(j, i, 0.0);
#pragma unroll 1
    for (int c = 0; c < pa.ncolors; ++c) {
#pragma unroll
        for (int j = 0; j < N; ++j) v[j] = (pa.color_of_col[j] == c && pa.nz_rows_of_col[j] != 0) ? 1.0 : 0.0;
        M::jac_mul(x, p, t, v, col);
        st.v[DSB_STAT_RHS_JAC_MULS] += 1;
#pragma unroll 1
        for (int j = 0; j < N; ++j) {
            if (pa.color_of_col[j] == c) {
                const uint64_t nz = pa.nz_rows_of_col[j];
#pragma unroll
                for (int i = 0; i < N; ++i) if ((nz >> i) & 1ull) store(j, i, col[i]);
            }
        }
    }
}

// Mass matrix by unit-vector products with beta = 0; identity when the model has no mass.
template <class M>
DSB_DEV void lane_mass_matrix(const double* p, double t, double (&Mm)[M::N][M::N]) {
    constexpr int N = M::N;
    if (!M::HAS_MASS) {
#pragma unroll
        for (int j = 0; j < N; ++j)
#pragma unroll
            for (int i = 0; i < N; ++i) Mm[j][i] = (i == j) ? 1.0 : 0.0;
        return;
    }
    double v[N], col[N];
#pragma unroll
    for (int i = 0; i < N; ++i) { v[i] = 0.0; col[i] = 0.0; }
#pragma unroll
    for (int j = 0; j < N; ++j) {
        v[j] = 1.0;
        M::mass(v, p, t, 0.0, col);
#pragma unroll
        for (int i = 0; i < N; ++i) Mm[j][i] = col[i];
        v[j] = 0.0;
    }
}
