// oracle/dsb_oracle_sdirk.cpp -- TEST INFRASTRUCTURE (see dsb_oracle.hpp).
// Restatement of `Sdirk` (crates/diffsol/src/ode_solver/sdirk.rs:147-543), the shared Runge-Kutta
// core `Rk` (ode_solver/runge_kutta.rs:100-175, 446-960, 1080-1127), `SdirkCallable`
// (crates/diffsol/src/op/sdirk.rs) and the TR-BDF2 / ESDIRK34 tableaux (ode_solver/tableau.rs:41-159),
// without sensitivities or output integration.  Vector / matrix products follow the
// evaluation order of nalgebra's `gemv` / `axpy` (first term assigned when beta == 0, the others added
// one column at a time; every term is `alpha * a_ij * x_j`).
//
// Reference quirks reproduced on purpose:
//   * SdirkCallable evaluates df/dy at tmp = phi + c * x where x is whatever reset_jacobian was handed
//     (state.y, NOT a stage increment) and phi is left over from the last set_phi (op/sdirk.rs:265-276);
//   * JacobianUpdate is fed h (not a_d * h) and starts at h_at_last = h0 (sdirk.rs:188-190);
//   * the first LU is set up lazily inside the first stage and counted as a Checkpoint setup
//     (runge_kutta.rs:661-665);
//   * accept test is strict `< 1`; the safety factor uses niter of the LAST stage only (sdirk.rs:498-509);
//   * the embedded error estimate is filtered through the LU of the iteration matrix (sdirk.rs:474-495).
#include "dsb_oracle.hpp"

#include <algorithm>

namespace orc {

namespace {

struct Tableau {
    int s = 0, order = 0;
    Vec a;          // s x s col-major
    Vec b, c, d;
    bool has_beta = false;
    Vec beta;       // s x 2 col-major
    double A(int i, int j) const { return a[(size_t)j * s + i]; }
};

// tableau.rs:41-97
Tableau tr_bdf2() {
    Tableau t;
    t.s = 3; t.order = 2;
    const double gamma = 2.0 - std::sqrt(2.0);
    const double d = gamma / 2.0;
    const double w = std::sqrt(2.0) / 4.0;
    t.a = {0.0, d, w, 0.0, d, w, 0.0, 0.0, d};
    t.b = {w, w, d};
    Vec b_hat = {(1.0 - w) / 3.0, (3.0 * w + 1.0) / 3.0, d / 3.0};
    t.d.resize(3);
    for (int i = 0; i < 3; ++i) t.d[i] = t.b[i] - b_hat[i];
    t.has_beta = true;
    t.beta = {2.0 * w, 2.0 * w, gamma - 1.0, -w, -w, 2.0 * w};
    t.c = {0.0, gamma, 1.0};
    return t;
}
// tableau.rs:101-159
Tableau esdirk34() {
    Tableau t;
    t.s = 4; t.order = 3;
    const double g = 0.435866521508459;
    t.a = {0.0, g, 0.1407377747247062, 0.102399400619911,
           0.0, g, -0.1083655513813208, -0.3768784522555561,
           0.0, 0.0, g, 0.8386125301271861,
           0.0, 0.0, 0.0, g};
    t.b = {t.A(3, 0), t.A(3, 1), t.A(3, 2), t.A(3, 3)};
    t.c = {0.0, 0.871733043016918, 0.4682387448518444, 1.0};
    t.d = {-0.05462549724041394, -0.49420889362599496, 0.22193449973506466, 0.32689989113134427};
    return t;
}

struct Sdirk : Method {
    const Problem& pr;
    int n;
    Tableau tab;
    // RkState x 2 (state / old_state, swapped on every accepted step)
    Vec y_, dy_, oy_, ody_;
    double t_ = 0, h_ = 0, ot_ = 0, oh_ = 0;
    Vec diff;                 // n x s col-major
    Vec error;
    Convergence convergence;
    DenseLU lu;
    bool is_jacobian_set = false;
    Stats statistics;
    bool has_tstop = false; double tstop = 0;
    JacobianUpdate jacobian_update;
    bool has_prev_error = false; double prev_error_norm = 0;
    // SdirkCallable
    double c = 0, op_h = 0;
    Vec phi, tmp, rhs_jac, mass_jac, newton_tmp, A;
    bool jacobian_is_stale = true;
    // root finding (runge_kutta.rs:43, 142-147, 935-948)
    RootFinder root_finder;
    double root_t_ = 0.0; int root_idx_ = -1;
    bool is_state_mutated = false;                         // runge_kutta.rs:50: set by state_mut() (:391-394)
    // forward sensitivities (Sdirk<.., SensEquations>, sdirk.rs:222-255; Rk: runge_kutta.rs:46, 518-523, 691-745, 812-822,
    // 917-920, 1237-1310): per parameter a stage-increment array, state.s / ds and their old_state twins; the sensitivity
    // residual SdirkCallable<SensEquations> has its own phi, shares c and h with the main one
    int ns = 0;
    std::vector<Vec> s_, ds_, os_, ods_, sdiff;
    Vec phi_s, sens_S, sens_y, sens_error;

    Sdirk(const Problem& p, const Tableau& t) : pr(p), n(p.n()), tab(t) {}

    double* DC(int j) { return diff.data() + (size_t)j * n; }
    const double* DC(int j) const { return diff.data() + (size_t)j * n; }

    // problem.rs:852-860 -> RkState::new_and_consistent(problem, tableau.order()); Rk::_new; Sdirk::_new
    int construct() {
        InitialState st;
        int err = new_and_consistent(pr, tab.order, &st);
        if (err) return err;
        y_ = st.y; dy_ = st.dy; t_ = st.t; h_ = st.h;
        oy_ = y_; ody_ = dy_; ot_ = t_; oh_ = h_;             // old_state = state.clone()
        if (pr.model.nroots > 0) {                             // Rk::_new, runge_kutta.rs:142-147
            root_finder.resize(pr.model.nroots, n);
            root_finder.init(pr, y_.data(), t_);
        }
        diff.assign((size_t)n * tab.s, 0.0);
        error.assign(n, 0.0);
        jacobian_update.init(pr.opt, 1.0);
        jacobian_update.update_jacobian(h_);
        jacobian_update.update_rhs_jacobian(h_);
        convergence.init(pr.rtol, pr.atol.data(), n, pr.opt.nonlinear_solver_tolerance, &pr.math);
        convergence.max_iter = pr.opt.max_nonlinear_solver_iterations;
        c = tab.A(1, 1);
        op_h = h_;
        phi.assign(n, 0.0); tmp.assign(n, 0.0); newton_tmp.assign(n, 0.0);
        rhs_jac.assign((size_t)n * n, 0.0); mass_jac.assign((size_t)n * n, 0.0); A.assign((size_t)n * n, 0.0);
        if (!pr.model.has_mass) pr.mass_matrix(t_, mass_jac.data());     // identity
        jacobian_is_stale = true;
        is_jacobian_set = false;
        if (pr.sens) {
            // RkState::new_with_sensitivities_and_consistent (state.rs:1032-1080): as for Bdf; Sdirk::new_augmented ends with
            // jacobian_updates(h, Checkpoint) (sdirk.rs:252), so the first LU is NOT the lazy one of the first stage
            if (!pr.model.sens_mul || !pr.model.init_sens) return ST_BAD_ARG;
            ns = pr.model.np;
            s_.assign(ns, Vec(n, 0.0)); ds_ = s_; os_ = s_; ods_ = s_;
            sdiff.assign(ns, Vec((size_t)n * tab.s, 0.0));
            phi_s.assign(n, 0.0); sens_y.assign(n, 0.0); sens_S.assign((size_t)n * ns, 0.0); sens_error.assign(n, 0.0);
            Vec e(ns, 0.0);
            for (int j = 0; j < ns; ++j) {
                e[j] = 1.0;
                pr.model.init_sens(pr.p.data(), pr.t0, e.data(), s_[j].data());
                e[j] = 0.0;
            }
            update_rhs_out_state(y_.data(), t_);
            for (int j = 0; j < ns; ++j) sens_rhs(j, s_[j].data(), t_, ds_[j].data());
            if (pr.model.has_mass) {
                Convergence ic_conv;
                ic_conv.init(pr.rtol, pr.atol.data(), n, pr.opt.nonlinear_solver_tolerance, &pr.math);
                ic_conv.max_iter = pr.opt.ic_max_newton_iterations;
                for (int j = 0; j < ns; ++j) {
                    int e2 = consistent_solve(pr, [this, j](const double* x, double tt, double* out) { sens_rhs(j, x, tt, out); },
                                              [this](const double*, double tt, double* J) { pr.jacobian(sens_y.data(), tt, J); },
                                              s_[j], ds_[j], &ic_conv, false);
                    if (e2) return e2;
                }
            }
            os_ = s_; ods_ = ds_;                              // old_state = state.clone()
            jacobian_updates(h_, CHECKPOINT);
        }
        return ST_OK;
    }

    // SensRhs::update_state / call_inplace (ode_equations/sens_equations.rs:129-134, 168-174), as in the Bdf restatement
    void update_rhs_out_state(const double* y, double t) {
        Vec v(ns, 0.0);
        for (int j = 0; j < ns; ++j) {
            v[j] = 1.0;
            pr.model.sens_mul(y, pr.p.data(), t, v.data(), sens_S.data() + (size_t)j * n);
            v[j] = 0.0;
        }
        for (int k = 0; k < n; ++k) sens_y[k] = y[k];
    }
    void sens_rhs(int index, const double* x, double t, double* out) const {
        pr.jac_mul(sens_y.data(), t, x, out);
        const double* col = sens_S.data() + (size_t)index * n;
        for (int k = 0; k < n; ++k) out[k] += col[k];
    }
    // SdirkCallable<SensEquations>::call_inplace (op/sdirk.rs:231-245): F(x) = M x - h rhs_s(phi_s + c x)
    void callable_sens(int index, const double* x, double t, double* out) {
        for (int i = 0; i < n; ++i) tmp[i] = c * x[i] + phi_s[i];
        sens_rhs(index, tmp.data(), t, out);
        const double beta = -op_h;
        if (pr.model.has_mass) pr.mass_gemv(x, t, beta, out);
        else for (int i = 0; i < n; ++i) out[i] = x[i] + beta * out[i];
    }

    // SdirkCallable::jacobian_inplace (op/sdirk.rs:257-292) + NalgebraLU::set_linearisation
    void reset_jacobian(const double* x, double t) {
        if (jacobian_is_stale) {
            for (int i = 0; i < n; ++i) tmp[i] = c * x[i] + phi[i];      // set_tmp: tmp = phi; tmp.axpy(c, x, 1)
            pr.jacobian(tmp.data(), t, rhs_jac.data());
            if (pr.model.has_mass) pr.mass_matrix(t, mass_jac.data());
            jacobian_is_stale = false;
        }
        const double beta = -(c * op_h);
        for (size_t q = 0; q < (size_t)n * n; ++q) A[q] = rhs_jac[q] * beta + mass_jac[q];
        lu.factor(A.data(), n);
        is_jacobian_set = true;
    }

    // sdirk.rs:256-304
    void jacobian_updates(double h, SolverState state) {
        bool did_update = false;
        if (jacobian_update.check_rhs_jacobian_update(h, state)) {
            jacobian_is_stale = true;
            reset_jacobian(y_.data(), t_);
            jacobian_update.update_rhs_jacobian(h);
            jacobian_update.update_jacobian(h);
            convergence.reset_eta();
            did_update = true;
        } else if (jacobian_update.check_jacobian_update(h, state)) {
            reset_jacobian(y_.data(), t_);
            jacobian_update.update_jacobian(h);
            convergence.reset_eta();
            did_update = true;
        }
        if (did_update) statistics.record_linear_solver_setup(state);
    }

    // SdirkCallable::call_inplace (op/sdirk.rs:231-245): F(x) = M x - h f(phi + c x)
    void callable(const double* x, double t, double* out) {
        for (int i = 0; i < n; ++i) tmp[i] = c * x[i] + phi[i];
        pr.rhs(tmp.data(), t, out);
        const double beta = -op_h;
        if (pr.model.has_mass) pr.mass_gemv(x, t, beta, out);
        else for (int i = 0; i < n; ++i) out[i] = x[i] + beta * out[i];     // y.axpy(1, x, beta): 1*x*1 + beta*y
    }

    bool newton_solve(Vec& xn, double t, const Vec& error_y, int sens_index = -1) {
        convergence.reset();
        for (int it = 0; it < convergence.max_iter; ++it) {
            if (sens_index >= 0) callable_sens(sens_index, xn.data(), t, newton_tmp.data());
            else callable(xn.data(), t, newton_tmp.data());
            if (!lu.solve(newton_tmp.data())) return false;
            for (int i = 0; i < n; ++i) xn[i] -= newton_tmp[i];
            double norm = convergence.norm(newton_tmp.data(), error_y.data());
            ConvStatus s = convergence.check_new_iteration(norm);
            if (s == CONVERGED) return true;
            if (s == DIVERGED) return false;
        }
        return false;
    }

    // Rk::do_stage_sdirk (runge_kutta.rs:631-750), main equation only.  false <=> the Newton solve failed.
    bool do_stage(int i, double h) {
        const double t = t_ + tab.c[i] * h;
        // set_phi(1, diff[:, 0..i], state.y, a_rows[i]): phi = y; phi.gemv(1, cols, a_row, 1)
        for (int k = 0; k < n; ++k) phi[k] = y_[k];
        for (int j = 0; j < i; ++j) {
            const double aij = tab.A(i, j);
            const double* dj = DC(j);
            for (int k = 0; k < n; ++k) phi[k] = dj[k] * aij + phi[k];
        }
        // predict_stage_sdirk (runge_kutta.rs:610-629) into old_state.dy
        if (i == 0) {
            for (int k = 0; k < n; ++k) ody_[k] = h * dy_[k];
        } else if (i == 1) {
            for (int k = 0; k < n; ++k) ody_[k] = DC(0)[k];
        } else {
            const double cc = (tab.c[i] - tab.c[i - 2]) / (tab.c[i - 1] - tab.c[i - 2]);
            const double al = -cc, be = 1.0 + cc;
            for (int k = 0; k < n; ++k) ody_[k] = al * DC(i - 2)[k] + be * DC(i - 1)[k];
        }
        if (!is_jacobian_set) {
            reset_jacobian(y_.data(), t);
            statistics.record_linear_solver_setup(CHECKPOINT);
        }
        const bool ok = newton_solve(ody_, t, y_);
        statistics.v[S_NL_ITERS] += convergence.niter;
        if (!ok) return false;
        // get_f_eval: y_stage = phi + c x
        for (int k = 0; k < n; ++k) oy_[k] = c * ody_[k] + phi[k];
        for (int k = 0; k < n; ++k) DC(i)[k] = ody_[k];
        // the sensitivity equations of the stage (runge_kutta.rs:691-745): f_p at the stage value, then per parameter the
        // same set_phi / predict / Newton solve / get_f_eval on sdiff[j], state.s[j], state.ds[j]; the iterations of a failed
        // solve ARE counted here (the statistics line precedes the `?`)
        if (ns > 0) {
            update_rhs_out_state(oy_.data(), t);
            for (int j = 0; j < ns; ++j) {
                double* sd = sdiff[j].data();
                for (int k = 0; k < n; ++k) phi_s[k] = s_[j][k];
                for (int q = 0; q < i; ++q) {
                    const double aiq = tab.A(i, q);
                    for (int k = 0; k < n; ++k) phi_s[k] = sd[(size_t)q * n + k] * aiq + phi_s[k];
                }
                if (i == 0) {
                    for (int k = 0; k < n; ++k) ods_[j][k] = h * ds_[j][k];
                } else if (i == 1) {
                    for (int k = 0; k < n; ++k) ods_[j][k] = sd[k];
                } else {
                    const double cc = (tab.c[i] - tab.c[i - 2]) / (tab.c[i - 1] - tab.c[i - 2]);
                    const double al = -cc, be = 1.0 + cc;
                    for (int k = 0; k < n; ++k) ods_[j][k] = al * sd[(size_t)(i - 2) * n + k] + be * sd[(size_t)(i - 1) * n + k];
                }
                const bool oks = newton_solve(ods_[j], t, s_[j], j);
                statistics.v[S_NL_ITERS] += convergence.niter;
                if (!oks) return false;
                for (int k = 0; k < n; ++k) os_[j][k] = c * ods_[j][k] + phi_s[k];
                for (int k = 0; k < n; ++k) sd[(size_t)i * n + k] = ods_[j][k];
            }
        }
        return true;
    }

    // runge_kutta.rs:752-781.  0 = nothing, 1 = TstopReached, < 0 = -status
    int handle_tstop(double ts) {
        double troundoff = 100.0 * std::numeric_limits<double>::epsilon() * (std::fabs(t_) + std::fabs(h_));
        if (std::fabs(t_ - ts) <= troundoff) return 1;
        if ((h_ > 0.0 && ts < t_ - troundoff) || (h_ < 0.0 && ts > t_ + troundoff)) return -ST_STOP_TIME_BEFORE_CURRENT;
        if ((h_ > 0.0 && t_ + h_ > ts + troundoff) || (h_ < 0.0 && t_ + h_ < ts - troundoff)) {
            double factor = (ts - t_) / h_;
            h_ *= factor;
        }
        return 0;
    }

    // Rk::factor (runge_kutta.rs:466-495) + pi_controller_raw (:1313-1335)
    double factor_of(double error_norm, double safety_factor) const {
        const double safety = 0.9 * safety_factor;
        const double order_f = (double)(tab.order + 1);
        const double ki = pr.opt.pi_control_integral / order_f;
        double raw;
        if (pr.opt.pi_control_proportional == 0.0 || !has_prev_error) raw = pr.math.pow(error_norm, -ki);
        else {
            const double kp = pr.opt.pi_control_proportional / order_f;
            raw = pr.math.pow(error_norm, -(ki + kp)) * pr.math.pow(prev_error_norm, kp);
        }
        double factor = safety * raw;
        if (factor > pr.opt.max_timestep_shrink && factor < pr.opt.min_timestep_growth) factor = 1.0;
        if (factor < pr.opt.min_timestep_shrink) factor = pr.opt.min_timestep_shrink;
        if (factor > pr.opt.max_timestep_growth) factor = pr.opt.max_timestep_growth;
        return factor;
    }

    // sdirk.rs:409-543
    StopReason step(int* err) override {
        if (is_state_mutated) {                              // rk.start_step() (runge_kutta.rs:446-464)
            if (pr.model.nroots > 0) root_finder.init(pr, y_.data(), t_);
            if (has_tstop) {
                int e = set_stop_time(tstop);
                if (e) { *err = e; return STEP_ERROR; }
            }
            is_state_mutated = false;
        }
        double h = h_;
        if (std::fabs(h) < pr.opt.min_timestep) { *err = ST_STEP_SIZE_TOO_SMALL; return STEP_ERROR; }
        op_h = h;
        int nattempts = 0;
        bool updated_jacobian = false;
        const int start = (tab.A(0, 0) == 0.0) ? 1 : 0;
        double factor = 1.0, error_norm = 0.0;
        while (true) {
            if (start == 1) {                                                      // start_step_attempt
                for (int k = 0; k < n; ++k) DC(0)[k] = h * dy_[k];
                for (int j = 0; j < ns; ++j)
                    for (int k = 0; k < n; ++k) sdiff[j][k] = h * ds_[j][k];
            }
            bool failed = false;
            for (int i = start; i < tab.s; ++i) {
                if (!do_stage(i, h)) { failed = true; break; }
            }
            if (failed) {
                if (!updated_jacobian) {
                    updated_jacobian = true;
                    jacobian_updates(h, FIRST_CONVERGENCE_FAIL);
                } else {
                    h *= 0.3;
                    convergence.reset_eta_timestep_change();
                    op_h = h;
                    jacobian_updates(h, SECOND_CONVERGENCE_FAIL);
                }
                has_prev_error = false;
                // rk.solve_fail
                statistics.v[S_NL_FAILS] += 1;
                if (statistics.v[S_NL_FAILS] > pr.opt.max_nonlinear_solver_failures) {
                    *err = ST_TOO_MANY_NONLINEAR_FAILURES; return STEP_ERROR;
                }
                if (std::fabs(h) < pr.opt.min_timestep) { *err = ST_STEP_SIZE_TOO_SMALL; return STEP_ERROR; }
                continue;
            }
            // rk.error_norm: error = diff . d ; [error = M error] ; error = LU^-1 error
            for (int k = 0; k < n; ++k) error[k] = DC(0)[k] * tab.d[0];
            for (int j = 1; j < tab.s; ++j)
                for (int k = 0; k < n; ++k) error[k] = DC(j)[k] * tab.d[j] + error[k];
            if (pr.model.has_mass) {
                Vec nx = error;
                for (int k = 0; k < n; ++k) error[k] = mass_jac[k] * nx[0];
                for (int j = 1; j < n; ++j)
                    for (int k = 0; k < n; ++k) error[k] = mass_jac[(size_t)j * n + k] * nx[j] + error[k];
            }
            if (!lu.solve(error.data())) { *err = ST_LU_SOLVE_FAILED; return STEP_ERROR; }
            {
                double e = squared_norm(error.data(), y_.data(), pr.atol.data(), pr.rtol, n);
                error_norm = (0.0 < e) ? e : 0.0;             // 0.max(err)
            }
            if (pr.sens_error_control) {                      // sdiff[j] . d, NOT filtered through the LU (runge_kutta.rs:812-822)
                for (int j = 0; j < ns; ++j) {
                    const double* sd = sdiff[j].data();
                    for (int k = 0; k < n; ++k) sens_error[k] = sd[k] * tab.d[0];
                    for (int q = 1; q < tab.s; ++q)
                        for (int k = 0; k < n; ++k) sens_error[k] = sd[(size_t)q * n + k] * tab.d[q] + sens_error[k];
                    const double es = squared_norm(sens_error.data(), s_[j].data(), pr.sens_atol.data(), pr.sens_rtol, n);
                    error_norm = (error_norm < es) ? es : error_norm;
                }
            }
            const double maxiter = (double)convergence.max_iter;
            const double niter = (double)convergence.niter;
            const double safety_factor = (2.0 * maxiter + 1.0) / (2.0 * maxiter + niter);
            factor = factor_of(error_norm, safety_factor);
            if (error_norm < 1.0) break;
            h *= factor;
            convergence.reset_eta_timestep_change();
            op_h = h;
            jacobian_updates(h, ERROR_TEST_FAIL);
            nattempts += 1;
            has_prev_error = false;
            // rk.error_test_fail
            statistics.v[S_ERROR_TEST_FAILS] += 1;
            if (nattempts >= pr.opt.max_error_test_failures) { *err = ST_TOO_MANY_ERROR_TEST_FAILURES; return STEP_ERROR; }
            if (std::fabs(h) < pr.opt.min_timestep) { *err = ST_STEP_SIZE_TOO_SMALL; return STEP_ERROR; }
        }
        // accept
        const double new_h = h * factor;
        if (factor != 1.0) convergence.reset_eta_timestep_change();
        op_h = new_h;
        jacobian_updates(new_h, STEP_SUCCESS);
        jacobian_update.step();
        has_prev_error = true; prev_error_norm = error_norm;
        // rk.step_accepted(h, new_h, rescale_dy = true)
        ot_ = t_ + h;
        oh_ = new_h;
        {
            const double inv_h = 1.0 / h;
            for (int k = 0; k < n; ++k) ody_[k] *= inv_h;
            for (int j = 0; j < ns; ++j)
                for (int k = 0; k < n; ++k) ods_[j][k] *= inv_h;
        }
        std::swap(y_, oy_); std::swap(dy_, ody_); std::swap(t_, ot_); std::swap(h_, oh_);
        std::swap(s_, os_); std::swap(ds_, ods_);
        statistics.v[S_STEPS] += 1;
        // check for a root within the accepted step (runge_kutta.rs:935-948)
        if (pr.model.nroots > 0) {
            auto interp = [this](double tq, double* yq) { return interpolate(tq, yq); };
            if (root_finder.check_root(pr, interp, y_.data(), t_, &root_t_, &root_idx_)) return ROOT_FOUND;
        }
        if (has_tstop) {
            int r = handle_tstop(tstop);
            if (r == 1) { has_tstop = false; return TSTOP_REACHED; }
            if (r < 0) { *err = -r; return STEP_ERROR; }
        }
        return INTERNAL_TIMESTEP;
    }

    // runge_kutta.rs:431-441
    int set_stop_time(double ts) override {
        has_tstop = true; tstop = ts;
        int r = handle_tstop(ts);
        if (r == 1) { has_tstop = false; return ST_STOP_TIME_AT_CURRENT; }
        if (r < 0) return -r;
        return ST_OK;
    }

    // runge_kutta.rs:1080-1127 (+ :962-981 beta dense output, :1004-1024 Hermite)
    int interpolate(double t, double* y) const override {
        if (is_state_mutated) {                              // runge_kutta.rs:1089-1096
            if (t != t_) return ST_INTERPOLATION_TIME_AFTER_CURRENT;
            for (int k = 0; k < n; ++k) y[k] = y_[k];
            return ST_OK;
        }
        const bool is_forward = h_ > 0.0;
        if ((is_forward && (t > t_ || t < ot_)) || (!is_forward && (t < t_ || t > ot_)))
            return ST_INTERPOLATION_TIME_AFTER_CURRENT;
        const double dt = t_ - ot_;
        const double theta = (dt == 0.0) ? 1.0 : (t - ot_) / dt;
        if (tab.has_beta) {
            const int s = tab.s;
            const double th1 = theta, th2 = theta * theta;
            Vec beta_f(s);
            for (int k = 0; k < s; ++k) beta_f[k] = tab.beta[k] * th1;
            for (int k = 0; k < s; ++k) beta_f[k] = tab.beta[(size_t)s + k] * th2 + beta_f[k];
            for (int k = 0; k < n; ++k) y[k] = oy_[k];
            for (int j = 0; j < s; ++j)
                for (int k = 0; k < n; ++k) y[k] = DC(j)[k] * beta_f[j] + y[k];
        } else {
            const double* f0 = DC(0);
            const double* f1 = DC(tab.s - 1);
            for (int k = 0; k < n; ++k) y[k] = y_[k];
            for (int k = 0; k < n; ++k) y[k] -= oy_[k];
            {
                const double al = theta - 1.0, be = 1.0 - 2.0 * theta;
                for (int k = 0; k < n; ++k) y[k] = al * f0[k] + be * y[k];
            }
            for (int k = 0; k < n; ++k) y[k] = theta * f1[k] + y[k];
            {
                const double al = 1.0 - theta, be = theta * (theta - 1.0);
                for (int k = 0; k < n; ++k) y[k] = al * oy_[k] + be * y[k];
            }
            for (int k = 0; k < n; ++k) y[k] = theta * y_[k] + y[k];
        }
        return ST_OK;
    }

    // runge_kutta.rs:1237-1310: the same dense output on (old_state.s[j], state.s[j], sdiff[j])
    int interpolate_sens(double t, double* out) const override {
        if (ns == 0 || is_state_mutated) return ST_BAD_ARG;
        const bool is_forward = h_ > 0.0;
        if ((is_forward && (t > t_ || t < ot_)) || (!is_forward && (t < t_ || t > ot_)))
            return ST_INTERPOLATION_TIME_AFTER_CURRENT;
        const double dt = t_ - ot_;
        const double theta = (dt == 0.0) ? 1.0 : (t - ot_) / dt;
        for (int j = 0; j < ns; ++j) {
            double* y = out + (size_t)j * n;
            const double* sd = sdiff[j].data();
            if (tab.has_beta) {
                const int s = tab.s;
                const double th1 = theta, th2 = theta * theta;
                Vec beta_f(s);
                for (int k = 0; k < s; ++k) beta_f[k] = tab.beta[k] * th1;
                for (int k = 0; k < s; ++k) beta_f[k] = tab.beta[(size_t)s + k] * th2 + beta_f[k];
                for (int k = 0; k < n; ++k) y[k] = os_[j][k];
                for (int q = 0; q < s; ++q)
                    for (int k = 0; k < n; ++k) y[k] = sd[(size_t)q * n + k] * beta_f[q] + y[k];
            } else {
                const double* f0 = sd;
                const double* f1 = sd + (size_t)(tab.s - 1) * n;
                for (int k = 0; k < n; ++k) y[k] = s_[j][k];
                for (int k = 0; k < n; ++k) y[k] -= os_[j][k];
                {
                    const double al = theta - 1.0, be = 1.0 - 2.0 * theta;
                    for (int k = 0; k < n; ++k) y[k] = al * f0[k] + be * y[k];
                }
                for (int k = 0; k < n; ++k) y[k] = theta * f1[k] + y[k];
                {
                    const double al = 1.0 - theta, be = theta * (theta - 1.0);
                    for (int k = 0; k < n; ++k) y[k] = al * os_[j][k] + be * y[k];
                }
                for (int k = 0; k < n; ++k) y[k] = theta * s_[j][k] + y[k];
            }
        }
        return ST_OK;
    }

    double root_t() const override { return root_t_; }
    int root_index() const override { return root_idx_; }
    // runge_kutta.rs:396-434 (no integrate_out, no sensitivities): y, dy interpolated at t, state.t = t
    int state_mut_back(double t) override {
        Vec ynew(n);
        int e = interpolate(t, ynew.data());
        if (e) return e;
        y_ = ynew;                          // dy is interpolated as well in the reference; apply_reset overwrites it
        t_ = t;
        is_state_mutated = true;            // through state_mut()
        return ST_OK;
    }
    // sdirk.rs:368-374 -> state.rs:279-306 through Rk::state_mut() (no mass matrix: y <- reset(y, t), dy <- f(y, t));
    // the step size, the Jacobian and its LU stay as they are: the next Sdirk::step only re-initialises the root finder
    // and the stop time (Rk::start_step)
    int apply_reset() override {
        if (!pr.model.reset || pr.model.has_mass) return ST_BAD_ARG;
        is_state_mutated = true;
        Vec ynew(n);
        pr.model.reset(y_.data(), pr.p.data(), t_, ynew.data());
        y_ = ynew;
        pr.rhs(y_.data(), t_, dy_.data());
        return ST_OK;
    }

    int residual_known_answer(double cc, double hh, const double* vec, const double* x, double t, double* F, double* Aout) override {
        c = cc; op_h = hh;                                          // SdirkCallable::new(eqn, c); set_h(h)
        for (int i = 0; i < n; ++i) phi[i] = vec[i];                // set_phi_direct
        callable(x, t, F);
        jacobian_is_stale = true;
        reset_jacobian(x, t);
        for (size_t q = 0; q < (size_t)n * n; ++q) Aout[q] = A[q];
        return ST_OK;
    }

    double t() const override { return t_; }
    double h() const override { return h_; }
    int cur_order() const override { return tab.order; }
    const double* y() const override { return y_.data(); }
    const Stats& stats() const override { return statistics; }
};

}  // namespace

Method* new_sdirk(const Problem& pr, int tableau, int* err) {
    Sdirk* s = new Sdirk(pr, tableau == 0 ? tr_bdf2() : esdirk34());
    *err = s->construct();
    if (*err) { delete s; return nullptr; }
    return s;
}

}  // namespace orc
