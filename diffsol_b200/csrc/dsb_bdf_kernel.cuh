// dsb_bdf_kernel.cuh -- `problem.bdf::<LS>()?.solve_dense(t_eval)` for every instance of a batch,
// one CUDA thread per instance, the whole integration inside one kernel.
//
// Restates (paths relative to /root/reference/crates/diffsol/src):
//   Bdf::_new                    ode_solver/bdf.rs:230-368
//   Bdf::step                    ode_solver/bdf.rs:1277-1589   (retry loop, order/step selection)
//   _predict_forward / set_psi   ode_solver/bdf.rs:667-692, op/bdf.rs:182-210
//   BdfCallable::call_inplace    op/bdf.rs:240-256             F(y) = M (y - y0 + psi) - c f(y)
//   BdfCallable::jacobian_inplace op/bdf.rs:273-300            A = M - c J
//   newton_iteration + NoLineSearch  crates/diffsol-nl/src/newton.rs:13-36, line_search.rs:48-69
//   _jacobian_updates            ode_solver/bdf.rs:465-506
//   _update_step_size, _compute_r ode_solver/bdf.rs:433-463, 508-577
//   _update_diff                 ode_solver/bdf.rs:646-664
//   error_control, predict_error_control   ode_solver/bdf.rs:812-932
//   pi_controller_raw            ode_solver/runge_kutta.rs:1313-1335
//   handle_tstop, set_stop_time  ode_solver/bdf.rs:694-731, 1591-1599
//   interpolate                  ode_solver/bdf.rs:767-782, 1080-1106
//   fn solve_dense               ode_solver/method.rs:721-848
// Every instance keeps its own h / order / Newton state / counters: the result for instance b is
// what the reference's CPU path returns for that instance alone (SURVEY.md section 3.6 quirks
// Q1-Q8 included).  LU, the Newton work vectors and all controller scalars are registers; the
// difference array D (n x 8) and df/dy (and M for DAEs), which are only touched between Newton
// solves, live in shared memory as one column per thread (word index * blockDim + tid: conflict-free),
// which also lets them be indexed by the run-time order.
#pragma once
#include "dsb_lane.cuh"

template <class M>
struct BdfLane {
    static constexpr int N = M::N;
    static constexpr int NP = M::NP;

    const DsbProblemArgs& pa;
    double p[NP > 0 ? NP : 1];
    // shared-memory column of this thread: D[j][i] at (j*N + i), then rhs_jac[j][i], then mass_jac[j][i]
    double* sm;
    static constexpr int SM_D = 0, SM_J = DSB_NDIFF * N, SM_M = SM_J + N * N;
    static constexpr int SM_WORDS = SM_M + (M::HAS_MASS ? N * N : 0);
    DSB_DEV double& D(int j, int i) { return sm[(SM_D + j * N + i) * DSB_LANE_THREADS]; }
    DSB_DEV const double& D(int j, int i) const { return sm[(SM_D + j * N + i) * DSB_LANE_THREADS]; }
    DSB_DEV double& Jm(int j, int i) { return sm[(SM_J + j * N + i) * DSB_LANE_THREADS]; }
    DSB_DEV double& Mm(int j, int i) { return sm[(SM_M + j * N + i) * DSB_LANE_THREADS]; }
    // BdfState
    int order;
    double y[N];
    double t, h;
    // Bdf
    LaneConvergence conv;
    LaneLU<N> lu;
    int n_equal_steps;
    double y_delta[N], y_predict[N];
    double t_predict;
    LaneStats st;
    bool has_tstop; double tstop;
    LaneJacobianUpdate ju;
    bool has_prev_error; double prev_error_norm;
    // BdfCallable
    double psi_neg_y0[N];
    double c;
    bool jacobian_is_stale;

    DSB_DEV BdfLane(const DsbProblemArgs& a, double* sm_) : pa(a), sm(sm_) {}

    DSB_DEV void set_c(double hh, double a) { c = hh * a; }

    DSB_DEV void reset_jacobian(const double (&x)[N], double tt) {
        if (jacobian_is_stale) {
            // the fresh Jacobian goes through lu.a (about to be overwritten anyway) on its way to shared memory
            lane_jacobian<M>(pa, x, p, tt, lu.a, st);
#pragma unroll
            for (int j = 0; j < N; ++j)
#pragma unroll
                for (int i = 0; i < N; ++i) Jm(j, i) = lu.a[j][i];
            if (M::HAS_MASS) {
                lane_mass_matrix<M>(p, tt, lu.a);
#pragma unroll
                for (int j = 0; j < N; ++j)
#pragma unroll
                    for (int i = 0; i < N; ++i) Mm(j, i) = lu.a[j][i];
            }
            jacobian_is_stale = false;
        }
        const double mc = -c;
#pragma unroll
        for (int j = 0; j < N; ++j)
#pragma unroll
            for (int i = 0; i < N; ++i) {
                // identity mass when the model has none (op/bdf.rs:141-143)
                const double m_ji = M::HAS_MASS ? Mm(j, i) : ((i == j) ? 1.0 : 0.0);
                lu.a[j][i] = Jm(j, i) * mc + m_ji;
            }
        lu.factor();
    }

    DSB_DEV void jacobian_updates(double cc, int state) {
        bool did_update = false;
        if (ju.check_rhs_jacobian_update(pa.opt, cc, state)) {
            jacobian_is_stale = true;
            reset_jacobian(y, t);
            ju.update_rhs_jacobian(cc);
            ju.update_jacobian(cc);
            conv.eta = pa.tab.eta_reset;
            did_update = true;
        } else if (ju.check_jacobian_update(pa.opt, cc, state)) {
            reset_jacobian(y, t);
            ju.update_jacobian(cc);
            conv.eta = pa.tab.eta_reset;
            did_update = true;
        }
        if (did_update) st.record_linear_solver_setup(state);
    }

    // D[:, 0..=K] <- D[:, 0..=K] * (R(K, factor) * U(K)).  Every product is accumulated in the order of
    // nalgebra's gemm (first term assigned, the rest added one by one: bdf.rs:521, 568-577), but R and
    // RU are produced one ROW at a time and folded into the new columns at once, so that only
    // 2 (K+1) coefficients are live instead of 2 (K+1)^2.
    template <int K>
    DSB_DEV void rescale_diff(double factor) {
        constexpr int NR = K + 1;
        const double* u = pa.tab.u[K];     // U = R(K, 1), column-major, leading dimension NR
        double rrow[NR];                   // R[i, l] for the current row i
        double nd[NR][N];
#pragma unroll
        for (int l = 0; l < NR; ++l) rrow[l] = 1.0;
#pragma unroll
        for (int i = 0; i < NR; ++i) {
            if (i > 0) {
                const double i_t = (double)i;
                rrow[0] = 0.0;
#pragma unroll
                for (int l = 1; l < NR; ++l) rrow[l] = rrow[l] * (i_t - 1.0 - factor * (double)l) / i_t;
            }
#pragma unroll
            for (int j = 0; j < NR; ++j) {
                double ru_ij = rrow[0] * u[j * NR + 0];          // RU[i, j] = sum_l R[i, l] U[l, j]
#pragma unroll
                for (int l = 1; l < NR; ++l) ru_ij = rrow[l] * u[j * NR + l] + ru_ij;
#pragma unroll
                for (int s = 0; s < N; ++s) {
                    if (i == 0) nd[j][s] = D(i, s) * ru_ij;
                    else nd[j][s] = D(i, s) * ru_ij + nd[j][s];
                }
            }
        }
#pragma unroll
        for (int j = 0; j < NR; ++j)
#pragma unroll
            for (int s = 0; s < N; ++s) D(j, s) = nd[j][s];
    }

    DSB_DEV int update_step_size(double factor, double* new_h_out) {
        const double new_h = factor * h;
        n_equal_steps = 0;
        switch (order) {
            case 1: rescale_diff<1>(factor); break;
            case 2: rescale_diff<2>(factor); break;
            case 3: rescale_diff<3>(factor); break;
            case 4: rescale_diff<4>(factor); break;
            default: rescale_diff<5>(factor); break;
        }
        set_c(new_h, pa.tab.alpha[order]);
        h = new_h;
        conv.eta = pa.tab.eta_reset_timestep;
        if (new_h_out) *new_h_out = new_h;
        if (dsb_abs(h) < pa.opt.min_timestep) return DSB_STATUS_STEP_SIZE_TOO_SMALL;
        return DSB_STATUS_OK;
    }

    DSB_DEV void update_diff(int ord, const double (&d)[N]) {
#pragma unroll
        for (int i = 0; i < N; ++i) {
            D(ord + 2, i) = d[i] - D(ord + 1, i);
            D(ord + 1, i) = d[i];
        }
        for (int j = ord; j >= 0; --j) {
#pragma unroll
            for (int i = 0; i < N; ++i) D(j, i) = D(j, i) + 1.0 * D(j + 1, i);
        }
    }

    DSB_DEV void predict_forward() {
#pragma unroll
        for (int i = 0; i < N; ++i) y_predict[i] = 0.0;
        for (int j = 0; j <= order; ++j) {
#pragma unroll
            for (int i = 0; i < N; ++i) y_predict[i] += D(j, i);
        }
#pragma unroll
        for (int i = 0; i < N; ++i) psi_neg_y0[i] = pa.tab.gamma[1] * D(1, i);
        for (int j = 2; j <= order; ++j) {
            const double g = pa.tab.gamma[j];
#pragma unroll
            for (int i = 0; i < N; ++i) psi_neg_y0[i] = g * D(j, i) + psi_neg_y0[i];
        }
        const double a = pa.tab.alpha[order];
#pragma unroll
        for (int i = 0; i < N; ++i) psi_neg_y0[i] *= a;
#pragma unroll
        for (int i = 0; i < N; ++i) psi_neg_y0[i] -= y_predict[i];
        t_predict = t + h;
    }

    DSB_DEV void callable(const double (&x)[N], double tt, double (&out)[N]) {
        M::rhs(x, p, tt, out);
        st.v[DSB_STAT_RHS_CALLS] += 1;
        double tmp[N];
#pragma unroll
        for (int i = 0; i < N; ++i) tmp[i] = x[i] + psi_neg_y0[i];
        const double mc = -c;
        if (M::HAS_MASS) {
            M::mass(tmp, p, tt, mc, out);
        } else {
#pragma unroll
            for (int i = 0; i < N; ++i) out[i] = tmp[i] + mc * out[i];
        }
    }

    DSB_DEV bool newton_solve(double (&xn)[N], double tt, const double (&error_y)[N]) {
        conv.reset();
        for (int it = 0; it < conv.max_iter; ++it) {
            double delta[N];
            callable(xn, tt, delta);
            if (!lu.solve(delta)) return false;
#pragma unroll
            for (int i = 0; i < N; ++i) xn[i] -= delta[i];
            const double norm = dsb_sqrt(lane_squared_norm<N>(delta, error_y, pa.atol, pa.rtol));
            const int s = conv.check_new_iteration(norm);
            if (s == LANE_CONVERGED) return true;
            if (s == LANE_DIVERGED) return false;
        }
        return false;
    }

    // 0 = nothing, 1 = TstopReached, < 0 = -status
    DSB_DEV int handle_tstop(double ts) {
        const double troundoff = 100.0 * 2.220446049250313e-16 * (dsb_abs(t) + dsb_abs(h));
        if (dsb_abs(t - ts) <= troundoff) { has_tstop = false; return 1; }
        if ((h > 0.0 && ts < t - troundoff) || (h < 0.0 && ts > t + troundoff)) {
            has_tstop = false;
            return -DSB_STATUS_STOP_TIME_BEFORE_CURRENT;
        }
        if ((h > 0.0 && t + h > ts + troundoff) || (h < 0.0 && t + h < ts - troundoff)) {
            const double factor = (ts - t) / h;
            (void)update_step_size(factor, nullptr);
        }
        return 0;
    }

    DSB_DEV double error_control() const {
        const double err = lane_squared_norm<N>(y_delta, y, pa.atol, pa.rtol) * pa.tab.error_const2[order - 1];
        return (0.0 < err) ? err : 0.0;
    }
    // squared norm of D[:, ord + 1] scaled by error_const2[ord]
    DSB_DEV double predict_error_control(int ord) const {
        double col[N];
#pragma unroll
        for (int i = 0; i < N; ++i) col[i] = D(ord + 1, i);
        const double err = lane_squared_norm<N>(col, y, pa.atol, pa.rtol) * pa.tab.error_const2[ord];
        return (0.0 < err) ? err : 0.0;
    }
    DSB_DEV double pi_controller_raw(double error_norm, int eff_order) const {
        const double order_f = (double)eff_order;
        const double ki = pa.opt.pi_control_integral / order_f;
        if (pa.opt.pi_control_proportional == 0.0 || !has_prev_error) return dsb_pow(error_norm, -ki);
        const double kp = pa.opt.pi_control_proportional / order_f;
        return dsb_pow(error_norm, -(ki + kp)) * dsb_pow(prev_error_norm, kp);
    }

    // Bdf::_new from the state the init kernel produced
    DSB_DEV void construct(const double (&y_init)[N], const double (&dy_init)[N], double h_init) {
        order = 1;
        t = pa.t0; h = h_init;
        conv.tol = pa.opt.nonlinear_solver_tolerance;
        conv.max_iter = pa.opt.max_nonlinear_solver_iterations;
        conv.eta = pa.tab.eta_reset;
        conv.reset();
        conv.old_norm = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) { y[i] = y_init[i]; psi_neg_y0[i] = 0.0; y_delta[i] = 0.0; y_predict[i] = 0.0; }
        jacobian_is_stale = true;
        set_c(h, pa.tab.alpha[order]);
        reset_jacobian(y, t);
#pragma unroll
        for (int j = 0; j < DSB_NDIFF; ++j)
#pragma unroll
            for (int i = 0; i < N; ++i) D(j, i) = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) { D(0, i) = y[i]; D(1, i) = dy_init[i] * h; }
        st.v[DSB_STAT_LINEAR_SOLVER_SETUPS] += 1;
        st.v[DSB_STAT_SETUPS_FROM_CHECKPOINT] += 1;
        ju.init(1.0);                 // jacobian_update.rs:27 -- h_at_last starts at ONE
        n_equal_steps = 0;
        has_tstop = false; tstop = 0.0;
        has_prev_error = false; prev_error_norm = 0.0;
        t_predict = t;
    }

    // returns 0 = InternalTimestep, 1 = TstopReached, < 0 = -status
    DSB_DEV int step() {
        double safety = 0.0, error_norm = 0.0;
        const int old_num_error_test_failures = st.v[DSB_STAT_ERROR_TEST_FAILURES];
        bool convergence_fail = false;
        double new_h = 0.0;
        predict_forward();
        while (true) {
            const int ord = order;
#pragma unroll
            for (int i = 0; i < N; ++i) y_delta[i] = y_predict[i];
            const bool ok = newton_solve(y_delta, t_predict, y_predict);
            st.v[DSB_STAT_NONLINEAR_SOLVER_ITERATIONS] += conv.niter;
            if (ok) {
#pragma unroll
                for (int i = 0; i < N; ++i) y_delta[i] -= y_predict[i];
            } else {
                st.v[DSB_STAT_NONLINEAR_SOLVER_FAILS] += 1;
                if (st.v[DSB_STAT_NONLINEAR_SOLVER_FAILS] > pa.opt.max_nonlinear_solver_failures)
                    return -DSB_STATUS_TOO_MANY_NONLINEAR_FAILURES;
                if (convergence_fail) {
                    has_prev_error = false;
                    const int e = update_step_size(0.3, &new_h);
                    if (e) return -e;
                    jacobian_updates(new_h * pa.tab.alpha[ord], DSB_SECOND_CONVERGENCE_FAIL);
                    predict_forward();
                } else {
                    has_prev_error = false;
                    jacobian_updates(h * pa.tab.alpha[ord], DSB_FIRST_CONVERGENCE_FAIL);
                    convergence_fail = true;
                }
                continue;
            }
            error_norm = error_control();
            const double maxiter = (double)conv.max_iter;
            const double niter = (double)conv.niter;
            safety = 0.9 * (2.0 * maxiter + 1.0) / (2.0 * maxiter + niter);
            if (error_norm <= 1.0) break;
            double factor = safety * pi_controller_raw(error_norm, ord + 1);
            has_prev_error = false;
            if (factor < pa.opt.min_timestep_shrink) factor = pa.opt.min_timestep_shrink;
            const int e = update_step_size(factor, &new_h);
            if (e) return -e;
            jacobian_updates(new_h * pa.tab.alpha[ord], DSB_ERROR_TEST_FAIL);
            predict_forward();
            st.v[DSB_STAT_ERROR_TEST_FAILURES] += 1;
            if (st.v[DSB_STAT_ERROR_TEST_FAILURES] - old_num_error_test_failures >= pa.opt.max_error_test_failures)
                return -DSB_STATUS_TOO_MANY_ERROR_TEST_FAILURES;
        }
        // accepted
        update_diff(order, y_delta);
#pragma unroll
        for (int i = 0; i < N; ++i) y[i] = y_predict[i];        // Q1: the PREDICTOR
        t = t_predict;
        st.v[DSB_STAT_STEPS] += 1;
        ju.step();
        has_prev_error = true; prev_error_norm = error_norm;
        n_equal_steps += 1;
        if (n_equal_steps > order) {
            const int ord = order;
            const double inf = dsb_from_bits(0x7ff0000000000000ULL);
            const double error_m_norm = ord > 1 ? predict_error_control(ord - 1) : inf;
            const double error_p_norm = ord < DSB_MAX_ORDER ? predict_error_control(ord + 1) : inf;
            const double f0 = pi_controller_raw(error_m_norm, ord);
            const double f1 = pi_controller_raw(error_norm, ord + 1);
            const double f2 = pi_controller_raw(error_p_norm, ord + 2);
            // Iterator::max_by keeps the LAST maximum
            int max_index = 0;
            double fmax = f0;
            if (!(fmax > f1)) { max_index = 1; fmax = f1; }
            if (!(fmax > f2)) { max_index = 2; fmax = f2; }
            const int new_order = ord + (max_index - 1);
            order = new_order;
            double factor = safety * fmax;
            if (factor > pa.opt.max_timestep_growth) factor = pa.opt.max_timestep_growth;
            if (factor < pa.opt.min_timestep_shrink) factor = pa.opt.min_timestep_shrink;
            if (factor >= pa.opt.min_timestep_growth || factor <= pa.opt.max_timestep_shrink
                || max_index == 0 || max_index == 2) {
                const int e = update_step_size(factor, &new_h);
                if (e) return -e;
                jacobian_updates(new_h * pa.tab.alpha[new_order], DSB_STEP_SUCCESS);
            }
        }
        if (has_tstop) {
            const int r = handle_tstop(tstop);
            if (r == 1) return 1;
            if (r < 0) return r;
        }
        return 0;
    }

    DSB_DEV int set_stop_time(double ts) {
        has_tstop = true; tstop = ts;
        const int r = handle_tstop(ts);
        if (r == 1) { has_tstop = false; return DSB_STATUS_STOP_TIME_AT_CURRENT; }
        if (r < 0) return -r;
        return DSB_STATUS_OK;
    }

    DSB_DEV int interpolate(double tq, double (&yo)[N]) const {
        const bool is_forward = h > 0.0;
        if ((is_forward && tq > t) || (!is_forward && tq < t)) return DSB_STATUS_INTERPOLATION_TIME_AFTER_CURRENT;
        double time_factor = 1.0;
#pragma unroll
        for (int i = 0; i < N; ++i) yo[i] = D(0, i);
        for (int j = 0; j < order; ++j) {
            const double j_t = (double)j;
            time_factor *= (tq - (t - h * j_t)) / (h * (1.0 + j_t));
#pragma unroll
            for (int i = 0; i < N; ++i) yo[i] = time_factor * D(j + 1, i) + yo[i];
        }
        return DSB_STATUS_OK;
    }
};

// One thread per instance: Bdf::new + solve_dense (method.rs:721-818).
template <class M>
__global__ void __launch_bounds__(DSB_LANE_THREADS) dsb_bdf_solve_dense_kernel(const __grid_constant__ DsbProblemArgs pa,
                                                                  const __grid_constant__ DsbBatchBuffers bb) {
    constexpr int N = M::N;
    constexpr int NP = M::NP;
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= pa.nbatch) return;
    const int64_t B = pa.nbatch;
    int status = bb.status[b];            // set by the init kernel
    if (status != DSB_STATUS_OK) return;

    extern __shared__ double dsb_lane_smem[];
    BdfLane<M> s(pa, dsb_lane_smem + threadIdx.x);
#pragma unroll
    for (int j = 0; j < NP; ++j) s.p[j] = bb.params[(int64_t)j * B + b];
#pragma unroll
    for (int k = 0; k < DSB_NSTATS; ++k) s.st.v[k] = bb.stats[(int64_t)k * B + b];
    {
        double y_init[N], dy_init[N];
#pragma unroll
        for (int i = 0; i < N; ++i) { y_init[i] = bb.y0[(int64_t)i * B + b]; dy_init[i] = bb.dy0[(int64_t)i * B + b]; }
        s.construct(y_init, dy_init, bb.h0[b]);
    }
    const int nt = pa.nt;
    const bool free_running = pa.free_running != 0;
    if (!free_running) status = s.set_stop_time(bb.t_eval[nt - 1]);
    int col = 0;
    while (status == DSB_STATUS_OK && col < nt) {
        if (free_running) {
            while (col < nt && !(dsb_abs(s.t) < dsb_abs(bb.t_eval[col]))) {
                double yo[N];
                const int e = s.interpolate(bb.t_eval[col], yo);
                if (e) { status = e; break; }
#pragma unroll
                for (int i = 0; i < N; ++i) bb.ys[((int64_t)col * N + i) * B + b] = yo[i];
                ++col;
            }
            if (col >= nt || status != DSB_STATUS_OK) break;
        }
        const int r = s.step();
        if (r < 0) { status = -r; break; }
        if (!free_running) {
            while (col < nt && bb.t_eval[col] <= s.t) {
                double yo[N];
                const int e = s.interpolate(bb.t_eval[col], yo);
                if (e) { status = e; break; }
#pragma unroll
                for (int i = 0; i < N; ++i) bb.ys[((int64_t)col * N + i) * B + b] = yo[i];
                ++col;
            }
            if (r == 1) break;
        }
    }
    bb.status[b] = status;
    bb.fin_t[b] = s.t; bb.fin_h[b] = s.h; bb.fin_order[b] = s.order;
#pragma unroll
    for (int k = 0; k < DSB_NSTATS; ++k) bb.stats[(int64_t)k * B + b] = s.st.v[k];
}
