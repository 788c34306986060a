// dsb_init_kernel.cuh -- `OdeSolverState::new_and_consistent` for every instance of the batch.
//
// Restates (paths relative to /root/reference/crates/diffsol/src):
//   new_without_initialise   ode_solver/state.rs:1086-1124    y = init(p, t0); dy = f(y, t0)
//   set_consistent           ode_solver/state.rs:84-162 + op/init.rs:14-131
//                            Newton with BacktrackingLineSearch (diffsol-nl/src/line_search.rs:115-201)
//                            on F(du, v) = -M_u du + f(u, v); g(u, v) for singular-mass DAEs
//   set_step_size            ode_solver/state.rs:1209-1277    (Hairer/Norsett/Wanner II.4.2)
// One thread per instance; results go to the batch-major state arrays y0 / dy0 / h0 that the
// integrator kernels start from (the reference's `BdfState` / `RkState` at t0).
#pragma once
#include "dsb_lane.cuh"

// The solve is shared by the state (set_consistent, state.rs:84-162: eq_rhs / eq_jac = the equations' own, a fresh
// Convergence, dy zeroed at the algebraic rows) and the sensitivity vectors of a DAE (set_consistent_augmented,
// state.rs:167-238: eq_rhs = SensRhs::call, eq_jac = SensRhs::jacobian_inplace, ONE Convergence for every parameter, ds kept
// at the algebraic rows).  eq_rhs(x, out) and eq_jac(J) do their own counting.
template <class M, class RhsFn, class JacFn>
DSB_DEV int lane_consistent_solve(const DsbProblemArgs& pa, const double* p, double (&y)[M::N], double (&dy)[M::N],
                                  RhsFn&& eq_rhs, JacFn&& eq_jac, LaneConvergence& conv, bool zero_dv, bool use_linesearch) {
    constexpr int N = M::N;
    if (!M::HAS_MASS) return DSB_STATUS_OK;
    const double t0 = pa.t0;
    bool is_alg[N];
    int nalg = 0;
    {
        double Mm[N][N];
        lane_mass_matrix<M>(p, t0, Mm);
#pragma unroll
        for (int i = 0; i < N; ++i) { is_alg[i] = (Mm[i][i] == 0.0); nalg += is_alg[i] ? 1 : 0; }
        if (nalg == 0) return DSB_STATUS_OK;
        // InitOp::new (op/init.rs:22-76): jac = (-M_u | f_v ; 0 | g_v), neg_mass = (-M_u | 0 ; 0 | 0)
        // is built below from Mm; keep Mm alive through the block.
        double rhs_jac[N][N];
        eq_jac(rhs_jac);
        LaneLU<N> lu;
        double neg_mass[N][N];
#pragma unroll
        for (int j = 0; j < N; ++j)
#pragma unroll
            for (int i = 0; i < N; ++i) {
                double jv = 0.0, nm = 0.0;
                if (!is_alg[j]) {
                    if (!is_alg[i]) { const double m_u = Mm[j][i] * -1.0; jv = m_u; nm = m_u; }
                } else {
                    jv = rhs_jac[j][i];
                }
                lu.a[j][i] = jv; neg_mass[j][i] = nm;
            }
        double jac[N][N];
#pragma unroll
        for (int j = 0; j < N; ++j)
#pragma unroll
            for (int i = 0; i < N; ++i) jac[j][i] = lu.a[j][i];

        double y0w[N];     // InitOp.y0
#pragma unroll
        for (int i = 0; i < N; ++i) y0w[i] = y[i];
        auto fun = [&](const double (&x)[N], double (&out)[N]) {
#pragma unroll
            for (int i = 0; i < N; ++i) if (is_alg[i]) y0w[i] = x[i];
            eq_rhs(y0w, out);
#pragma unroll
            for (int j = 0; j < N; ++j)
#pragma unroll
                for (int i = 0; i < N; ++i) out[i] = neg_mass[j][i] * x[j] + out[i];
        };
        double y_tmp[N], yerr[N];
#pragma unroll
        for (int i = 0; i < N; ++i) { y_tmp[i] = is_alg[i] ? y[i] : dy[i]; yerr[i] = y_tmp[i]; }
        const double tau = pa.opt.ic_step_reduction_factor, c_armijo = pa.opt.ic_armijo_constant;
        const double steptol = pa.tab.ic_steptol;
        const int ls_max_iter = pa.opt.ic_max_linesearch_iterations;

        bool ok = false;
        for (int setup = 0; setup < pa.opt.ic_max_linear_solver_setups; ++setup) {
#pragma unroll
            for (int j = 0; j < N; ++j)
#pragma unroll
                for (int i = 0; i < N; ++i) lu.a[j][i] = jac[j][i];
            lu.factor();
            conv.reset();
            double delta[N], x0[N], delta0[N];
#pragma unroll
            for (int i = 0; i < N; ++i) delta[i] = 0.0;
            double ls_norm = 1.0;
            int result = -1;   // 0 ok, 1 max iterations, 2 other error
            for (int it = 0; it < conv.max_iter && result < 0; ++it) {
                int res = LANE_CONTINUE;
                bool have_res = false;
                if (use_linesearch) {
                    if (conv.niter == 0) {
                        fun(y_tmp, delta);
                        if (!lu.solve(delta)) { result = 2; break; }
                        ls_norm = dsb_sqrt(lane_squared_norm<N>(delta, yerr, pa.atol, pa.rtol));
                        if (conv.check_norm(ls_norm) == LANE_CONVERGED) {
#pragma unroll
                            for (int i = 0; i < N; ++i) y_tmp[i] -= delta[i];
                            res = LANE_CONVERGED; have_res = true;
                        }
                    }
                    if (!have_res) {
#pragma unroll
                        for (int i = 0; i < N; ++i) { x0[i] = y_tmp[i]; delta0[i] = delta[i]; }
                        const double norm = ls_norm;
                        const double phi0 = norm * norm * 0.5, two_phi0 = norm * norm;
                        const double min_alpha = steptol / norm;
                        double alpha = 1.0;
                        int ls_status = 1;
                        for (int i = 0; i < ls_max_iter; ++i) {
#pragma unroll
                            for (int q = 0; q < N; ++q) y_tmp[q] = (-alpha) * delta0[q] + y_tmp[q];
                            fun(y_tmp, delta);
                            if (!lu.solve(delta)) { ls_status = 2; break; }
                            const double new_norm = dsb_sqrt(lane_squared_norm<N>(delta, yerr, pa.atol, pa.rtol));
                            const double phi1 = new_norm * new_norm * 0.5;
                            if (phi1 <= phi0 - c_armijo * alpha * two_phi0) {
                                ls_norm = new_norm;
                                res = conv.check_norm(new_norm); have_res = true; ls_status = 0;
                                break;
                            }
                            if (alpha < min_alpha) { ls_status = 2; break; }
                            alpha *= tau;
#pragma unroll
                            for (int q = 0; q < N; ++q) y_tmp[q] = x0[q];
                        }
                        if (ls_status != 0) { result = 2; break; }
                    }
                } else {
                    fun(y_tmp, delta);
                    if (!lu.solve(delta)) { result = 2; break; }
#pragma unroll
                    for (int i = 0; i < N; ++i) y_tmp[i] -= delta[i];
                    res = conv.check_new_iteration(dsb_sqrt(lane_squared_norm<N>(delta, yerr, pa.atol, pa.rtol)));
                }
                if (res == LANE_CONVERGED) result = 0;
                else if (res == LANE_DIVERGED) result = 2;
            }
            if (result < 0) result = 1;
            if (result == 0) { ok = true; break; }
            if (result == 2) return DSB_STATUS_INITIAL_CONDITION_DID_NOT_CONVERGE;
#pragma unroll
            for (int i = 0; i < N; ++i) yerr[i] = y_tmp[i];
        }
        if (!ok) return DSB_STATUS_INITIAL_CONDITION_DID_NOT_CONVERGE;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            if (is_alg[i]) { y[i] = y_tmp[i]; if (zero_dv) dy[i] = 0.0; }
            else dy[i] = y_tmp[i];
        }
    }
    return DSB_STATUS_OK;
}

template <class M>
DSB_DEV int lane_set_consistent(const DsbProblemArgs& pa, const double* p, double (&y)[M::N], double (&dy)[M::N],
                                LaneStats& st) {
    constexpr int N = M::N;
    if (!M::HAS_MASS) return DSB_STATUS_OK;
    LaneConvergence conv;
    conv.tol = pa.opt.nonlinear_solver_tolerance;
    conv.eta = pa.tab.eta_reset;
    conv.max_iter = pa.opt.ic_max_newton_iterations;
    conv.reset();
    const double t0 = pa.t0;
    return lane_consistent_solve<M>(pa, p, y, dy,
                                    [&](const double (&x)[N], double (&out)[N]) { M::rhs(x, p, t0, out); st.v[DSB_STAT_RHS_CALLS] += 1; },
                                    [&](double (&J)[N][N]) { lane_jacobian<M>(pa, y, p, t0, J, st); }, conv, true, pa.opt.ic_use_linesearch != 0);
}

template <class M>
DSB_DEV double lane_initial_step_size(const DsbProblemArgs& pa, const double* p, const double (&y0)[M::N],
                                      const double (&f0)[M::N], int solver_order, LaneStats& st) {
    constexpr int N = M::N;
    const bool is_neg_h = pa.h0 < 0.0;
    const double d0 = dsb_sqrt(lane_squared_norm<N>(y0, y0, pa.atol, pa.rtol));
    const double d1 = dsb_sqrt(lane_squared_norm<N>(f0, y0, pa.atol, pa.rtol));
    const double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * (d0 / d1);
    double y1[N], f1[N];
    if (is_neg_h) {
#pragma unroll
        for (int i = 0; i < N; ++i) y1[i] = f0[i] * (-h0) + y0[i];
        M::rhs(y1, p, pa.t0 - h0, f1);
    } else {
#pragma unroll
        for (int i = 0; i < N; ++i) y1[i] = f0[i] * h0 + y0[i];
        M::rhs(y1, p, pa.t0 + h0, f1);
    }
    st.v[DSB_STAT_RHS_CALLS] += 1;
    double df[N];
#pragma unroll
    for (int i = 0; i < N; ++i) df[i] = f1[i] - f0[i];
    const double d2 = dsb_sqrt(lane_squared_norm<N>(df, y0, pa.atol, pa.rtol)) / dsb_abs(h0);
    double max_d = d2;
    if (max_d < d1) max_d = d1;
    double h1;
    if (max_d < 1e-15) {
        h1 = h0 * 1e-3;
        if (h1 < 1e-6) h1 = 1e-6;
    } else {
        h1 = dsb_pow(0.01 / max_d, 1.0 / (1.0 + (double)solver_order));
    }
    double h = 100.0 * h0;
    if (h > h1) h = h1;
    if (is_neg_h) h = -h;
    return h;
}

// One thread per instance.  solver_order = 1 for Bdf (problem.rs:597-602), the tableau order for Sdirk.
template <class M>
__global__ void __launch_bounds__(128) dsb_init_kernel(const __grid_constant__ DsbProblemArgs pa,
                                                       const __grid_constant__ DsbBatchBuffers bb, int solver_order) {
    constexpr int N = M::N;
    constexpr int NP = M::NP;
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= pa.nbatch) return;
    const int64_t B = pa.nbatch;
    double p[NP > 0 ? NP : 1];
#pragma unroll
    for (int j = 0; j < NP; ++j) p[j] = bb.params[(int64_t)j * B + b];
    LaneStats st;
    st.clear();
    double y[N], dy[N];
    M::init(p, pa.t0, y);
    M::rhs(y, p, pa.t0, dy);
    st.v[DSB_STAT_RHS_CALLS] += 1;
    int status = lane_set_consistent<M>(pa, p, y, dy, st);
    double h = pa.h0;
    if (status == DSB_STATUS_OK) h = lane_initial_step_size<M>(pa, p, y, dy, solver_order, st);
#pragma unroll
    for (int i = 0; i < N; ++i) {
        bb.y0[(int64_t)i * B + b] = y[i];
        bb.dy0[(int64_t)i * B + b] = dy[i];
    }
    bb.h0[b] = h;
    bb.status[b] = status;
#pragma unroll
    for (int s = 0; s < DSB_NSTATS; ++s) bb.stats[(int64_t)s * B + b] = st.v[s];
}
