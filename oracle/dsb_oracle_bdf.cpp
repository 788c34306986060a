// oracle/dsb_oracle_bdf.cpp -- TEST INFRASTRUCTURE (see dsb_oracle.hpp).
// Restatement of `Bdf` (crates/diffsol/src/ode_solver/bdf.rs), `BdfState` (bdf_state.rs),
// `BdfCallable` (crates/diffsol/src/op/bdf.rs) and `newton_iteration` + `NoLineSearch`
// (crates/diffsol-nl/src/newton.rs:13-36, line_search.rs:48-69), without sensitivities, output
// integration or root finding.  Quirks Q1-Q8 of SURVEY.md section 3.6 are reproduced on purpose.
#include "dsb_oracle.hpp"

#include <algorithm>

namespace orc {

namespace {

const int MAX_ORDER = 5;              // bdf_state.rs:44
const int NCOLS = MAX_ORDER + 3;      // diff is n x (MAX_ORDER + 3)

struct Bdf : Method {
    const Problem& pr;
    int n;
    // BdfState
    int order = 1;
    Vec diff, diff_tmp;               // n x NCOLS col-major; ping-pong buffers (bdf.rs:568-577)
    Vec y_, dy_;
    double t_ = 0, h_ = 0;
    // Bdf
    Convergence convergence;
    DenseLU lu;
    int n_equal_steps = 0;
    Vec y_delta, y_predict;
    double t_predict = 0;
    Vec u;                            // (order+1)^2 col-major
    double alpha[MAX_ORDER + 1], gamma[MAX_ORDER + 1], error_const2[MAX_ORDER + 1];
    Stats statistics;
    bool has_tstop = false; double tstop = 0;
    JacobianUpdate jacobian_update;
    bool has_prev_error = false; double prev_error_norm = 0;
    // BdfCallable
    Vec psi_neg_y0, tmp, rhs_jac, mass_jac, newton_tmp, A;
    double c = 0;
    bool jacobian_is_stale = true;
    // root finding (bdf.rs:143, 301-306, 1566-1579)
    RootFinder root_finder;
    double root_t_ = 0.0; int root_idx_ = -1;
    bool is_state_modified = false;                        // bdf.rs:1223-1226: set by state_mut() / state_mut_back()
    // forward sensitivities (Bdf<.., SensEquations>, bdf.rs:128-141): one difference array per parameter, all rescaled
    // through the ONE diff_tmp ping-pong buffer (bdf.rs:539-541); the sensitivity residual keeps its OWN c, which stays 0
    // until the first step-size update (op/bdf.rs:61 new_no_jacobian, bdf.rs:551-553)
    int ns = 0;
    std::vector<Vec> s_, ds_, sdiff, s_deltas;
    Vec s_predict, sens_S, sens_y, psi_s;                  // sens_S = f_p at the predictor, n x np col-major (SensRhs::update_state)
    double c_sens = 0.0;

    explicit Bdf(const Problem& p) : pr(p), n(p.n()) {}

    double* D(int j) { return diff.data() + (size_t)j * n; }
    const double* D(int j) const { return diff.data() + (size_t)j * n; }

    // bdf.rs:244-368
    int construct() {
        InitialState st;
        int err = new_and_consistent(pr, 1, &st);      // problem.rs:597-602: bdf_state -> solver_order = 1
        if (err) return err;
        y_ = st.y; dy_ = st.dy; t_ = st.t; h_ = st.h; order = 1;
        if (pr.model.nroots > 0) {                         // bdf.rs:301-306
            root_finder.resize(pr.model.nroots, n);
            root_finder.init(pr, y_.data(), t_);
        }
        const double kappa[6] = {0.0, -0.1850, -1.0 / 9.0, -0.0823, -0.0415, 0.0};
        alpha[0] = 0.0; gamma[0] = 0.0; error_const2[0] = 1.0;
        for (int i = 1; i <= MAX_ORDER; ++i) {
            double i_t = (double)i;
            double one_over_i = 1.0 / i_t;
            double one_over_i_plus_one = 1.0 / (i_t + 1.0);
            gamma[i] = gamma[i - 1] + one_over_i;
            alpha[i] = 1.0 / ((1.0 - kappa[i]) * gamma[i]);
            double e = kappa[i] * gamma[i] + one_over_i_plus_one;
            error_const2[i] = e * e;
        }
        convergence.init(pr.rtol, pr.atol.data(), n, pr.opt.nonlinear_solver_tolerance, &pr.math);
        convergence.max_iter = pr.opt.max_nonlinear_solver_iterations;
        psi_neg_y0.assign(n, 0.0); tmp.assign(n, 0.0); newton_tmp.assign(n, 0.0);
        rhs_jac.assign((size_t)n * n, 0.0); mass_jac.assign((size_t)n * n, 0.0); A.assign((size_t)n * n, 0.0);
        if (!pr.model.has_mass) pr.mass_matrix(t_, mass_jac.data());   // identity, op/bdf.rs:141-143
        set_c(h_, alpha[order]);
        reset_jacobian(y_.data(), t_);
        // state.set_problem -> initialise_diff_to_first_order (bdf_state.rs:72-78)
        diff.assign((size_t)n * NCOLS, 0.0); diff_tmp.assign((size_t)n * NCOLS, 0.0);
        for (int i = 0; i < n; ++i) { D(0)[i] = y_[i]; D(1)[i] = dy_[i] * h_; }
        y_delta.assign(n, 0.0); y_predict.assign(n, 0.0);
        u = compute_r(order, 1.0);
        statistics.v[S_SETUPS] = 1; statistics.v[S_SETUPS_CHECKPOINT] = 1;
        jacobian_update.init(pr.opt, 1.0);             // jacobian_update.rs:27 -- h_at_last starts at ONE
        if (pr.sens) {
            // state.rs:1158-1180 initialise_augmented_state (s_i = (d y0 / d p) e_i), :178-189 set_consistent_augmented without
            // algebraic rows (ds_i = J(y0) s_i + f_p e_i), bdf_state.rs:88-98 initialise_sdiff_to_first_order
            if (!pr.model.sens_mul || !pr.model.init_sens) return ST_BAD_ARG;
            ns = pr.model.np;
            s_.assign(ns, Vec(n, 0.0)); ds_ = s_; s_deltas = s_;
            sdiff.assign(ns, Vec((size_t)n * NCOLS, 0.0));
            s_predict.assign(n, 0.0); psi_s.assign(n, 0.0); sens_y.assign(n, 0.0); sens_S.assign((size_t)n * ns, 0.0);
            Vec e(ns, 0.0);
            for (int i = 0; i < ns; ++i) {
                e[i] = 1.0;
                pr.model.init_sens(pr.p.data(), pr.t0, e.data(), s_[i].data());
                e[i] = 0.0;
            }
            update_rhs_out_state(y_.data(), t_);
            for (int i = 0; i < ns; ++i) sens_rhs(i, s_[i].data(), t_, ds_[i].data());
            if (pr.model.has_mass) {
                // set_consistent_augmented's algebraic part (state.rs:191-237): one Convergence for every parameter
                Convergence ic_conv;
                ic_conv.init(pr.rtol, pr.atol.data(), n, pr.opt.nonlinear_solver_tolerance, &pr.math);
                ic_conv.max_iter = pr.opt.ic_max_newton_iterations;
                for (int i = 0; i < ns; ++i) {
                    int e = consistent_solve(pr, [this, i](const double* x, double tt, double* out) { sens_rhs(i, x, tt, out); },
                                             [this](const double*, double tt, double* J) { pr.jacobian(sens_y.data(), tt, J); },
                                             s_[i], ds_[i], &ic_conv, false);
                    if (e) return e;
                }
            }
            for (int i = 0; i < ns; ++i)
                for (int k = 0; k < n; ++k) { sdiff[i][k] = s_[i][k]; sdiff[i][(size_t)n + k] = ds_[i][k] * h_; }
        }
        return ST_OK;
    }

    // SensRhs::update_state (sens_equations.rs:129-134) through _default_sens_inplace (op/nonlinear_op.rs:72-81)
    void update_rhs_out_state(const double* y, double t) {
        Vec v(ns, 0.0);
        for (int j = 0; j < ns; ++j) {
            v[j] = 1.0;
            pr.model.sens_mul(y, pr.p.data(), t, v.data(), sens_S.data() + (size_t)j * n);
            v[j] = 0.0;
        }
        for (int k = 0; k < n; ++k) sens_y[k] = y[k];
    }
    // SensRhs::call_inplace (sens_equations.rs:168-174): J(y stored) x + S[:, index]
    void sens_rhs(int index, const double* x, double t, double* out) const {
        pr.jac_mul(sens_y.data(), t, x, out);
        const double* col = sens_S.data() + (size_t)index * n;
        for (int k = 0; k < n; ++k) out[k] += col[k];
    }
    // BdfCallable<SensEquations>::call_inplace (op/bdf.rs:240-256) with its own psi and c
    void callable_sens(int index, const double* x, double t, double* out) {
        sens_rhs(index, x, t, out);
        for (int k = 0; k < n; ++k) tmp[k] = x[k] + psi_s[k];
        const double mc = -c_sens;
        if (pr.model.has_mass) pr.mass_gemv(tmp.data(), t, mc, out);
        else for (int k = 0; k < n; ++k) out[k] = tmp[k] + mc * out[k];
    }
    // bdf.rs:934-989.  false <=> a sensitivity solve failed (the iterations of the failed solve are NOT counted: the `?`
    // returns before the statistics line)
    bool sensitivity_solve(double t_new) {
        update_rhs_out_state(y_predict.data(), t_new);
        for (int i = 0; i < ns; ++i) {
            const double* d = sdiff[i].data();
            for (int k = 0; k < n; ++k) s_predict[k] = 0.0;
            for (int j = 0; j <= order; ++j)
                for (int k = 0; k < n; ++k) s_predict[k] += d[(size_t)j * n + k];
            for (int k = 0; k < n; ++k) psi_s[k] = gamma[1] * d[(size_t)n + k];
            for (int j = 2; j <= order; ++j)
                for (int k = 0; k < n; ++k) psi_s[k] = gamma[j] * d[(size_t)j * n + k] + psi_s[k];
            for (int k = 0; k < n; ++k) psi_s[k] *= alpha[order];
            for (int k = 0; k < n; ++k) psi_s[k] -= s_predict[k];
            s_[i] = s_predict;
            const bool ok = newton_solve_with([this, i](const double* x, double t, double* out) { callable_sens(i, x, t, out); },
                                              s_[i], t_new, s_predict);
            if (!ok) return false;
            statistics.v[S_NL_ITERS] += convergence.niter;
            for (int k = 0; k < n; ++k) s_deltas[i][k] = s_[i][k];
            for (int k = 0; k < n; ++k) s_deltas[i][k] -= s_predict[k];
        }
        return true;
    }

    void set_c(double h, double a) { c = h * a; }       // op/bdf.rs:179-181

    // BdfCallable::jacobian_inplace (op/bdf.rs:273-300) + NalgebraLU::set_linearisation (lu.rs:42-51)
    void reset_jacobian(const double* x, double t) {
        if (jacobian_is_stale) {
            pr.jacobian(x, t, rhs_jac.data());
            if (pr.model.has_mass) pr.mass_matrix(t, mass_jac.data());
            jacobian_is_stale = false;
        }
        // scale_add_and_assign(mass, -c, rhs_jac): A = rhs_jac; A *= -c; A += mass
        const double mc = -c;
        for (size_t q = 0; q < (size_t)n * n; ++q) A[q] = rhs_jac[q] * mc + mass_jac[q];
        lu.factor(A.data(), n);
    }

    // bdf.rs:433-463
    static Vec compute_r(int order, double factor) {
        int nr = order + 1;
        Vec r((size_t)nr * nr, 0.0);
        for (int j = 0; j < nr; ++j) r[(size_t)j * nr] = 1.0;
        for (int j = 1; j < nr; ++j) {
            double j_t = (double)j;
            for (int i = 1; i < nr; ++i) {
                double i_t = (double)i;
                size_t idx = (size_t)j * nr + i;
                r[idx] = r[idx - 1] * (i_t - 1.0 - factor * j_t) / i_t;
            }
        }
        return r;
    }

    // bdf.rs:465-506
    void jacobian_updates(double cc, SolverState state) {
        bool did_update = false;
        if (jacobian_update.check_rhs_jacobian_update(cc, state)) {
            jacobian_is_stale = true;
            reset_jacobian(y_.data(), t_);
            jacobian_update.update_rhs_jacobian(cc);
            jacobian_update.update_jacobian(cc);
            convergence.reset_eta();
            did_update = true;
        } else if (jacobian_update.check_jacobian_update(cc, state)) {
            reset_jacobian(y_.data(), t_);
            jacobian_update.update_jacobian(cc);
            convergence.reset_eta();
            did_update = true;
        }
        if (did_update) statistics.record_linear_solver_setup(state);
    }

    // bdf.rs:508-577.  RU = R(order, factor) * U, D[:, 0..=order] = D[:, 0..=order] * RU, both
    // evaluated the way nalgebra's small-matrix gemm does (column-by-column gemv, sequential axpys).
    int update_step_size(double factor, double* new_h_out) {
        double new_h = factor * h_;
        n_equal_steps = 0;
        int nr = order + 1;
        Vec r = compute_r(order, factor);
        Vec ru((size_t)nr * nr, 0.0);
        for (int j = 0; j < nr; ++j)
            for (int l = 0; l < nr; ++l) {
                double ulj = u[(size_t)j * nr + l];
                for (int i = 0; i < nr; ++i) {
                    size_t ij = (size_t)j * nr + i;
                    if (l == 0) ru[ij] = r[(size_t)l * nr + i] * ulj;
                    else ru[ij] = r[(size_t)l * nr + i] * ulj + ru[ij];
                }
            }
        auto rescale = [&](Vec& dd) {                       // _update_diff_for_step_size (bdf.rs:567-577)
            for (int j = 0; j < nr; ++j)
                for (int l = 0; l < nr; ++l) {
                    double rulj = ru[(size_t)j * nr + l];
                    const double* dl = dd.data() + (size_t)l * n;
                    double* out = diff_tmp.data() + (size_t)j * n;
                    for (int i = 0; i < n; ++i) {
                        if (l == 0) out[i] = dl[i] * rulj;
                        else out[i] = dl[i] * rulj + out[i];
                    }
                }
            std::swap(dd, diff_tmp);
        };
        rescale(diff);
        for (int i = 0; i < ns; ++i) rescale(sdiff[i]);     // bdf.rs:539-541: the same diff_tmp, so the buffers rotate
        set_c(new_h, alpha[order]);
        if (ns > 0) c_sens = new_h * alpha[order];          // bdf.rs:551-553
        h_ = new_h;
        convergence.reset_eta_timestep_change();
        if (new_h_out) *new_h_out = new_h;
        if (std::fabs(h_) < pr.opt.min_timestep) return ST_STEP_SIZE_TOO_SMALL;
        return ST_OK;
    }

    // bdf.rs:646-664
    void update_diff(int ord, const Vec& d) { update_diff_of(diff, ord, d); }
    void update_diff_of(Vec& dd, int ord, const Vec& d) {
        auto C = [&](int j) { return dd.data() + (size_t)j * n; };
        for (int i = 0; i < n; ++i) C(ord + 2)[i] = d[i] - C(ord + 1)[i];
        for (int i = 0; i < n; ++i) C(ord + 1)[i] = d[i];
        for (int j = ord; j >= 0; --j)
            for (int i = 0; i < n; ++i) C(j)[i] = C(j)[i] + 1.0 * C(j + 1)[i];
    }

    // bdf.rs:667-692 + op/bdf.rs:182-210
    void predict_forward() {
        for (int i = 0; i < n; ++i) y_predict[i] = 0.0;
        for (int j = 0; j <= order; ++j)
            for (int i = 0; i < n; ++i) y_predict[i] += D(j)[i];
        // set_psi: psi = gamma[1]*D1 ; psi += gamma[i]*Di ; psi *= alpha[order]
        for (int i = 0; i < n; ++i) psi_neg_y0[i] = gamma[1] * D(1)[i];
        for (int j = 2; j <= order; ++j)
            for (int i = 0; i < n; ++i) psi_neg_y0[i] = gamma[j] * D(j)[i] + psi_neg_y0[i];
        for (int i = 0; i < n; ++i) psi_neg_y0[i] *= alpha[order];
        for (int i = 0; i < n; ++i) psi_neg_y0[i] -= y_predict[i];
        t_predict = t_ + h_;
    }

    // BdfCallable::call_inplace (op/bdf.rs:240-256): F(y) = M (y - y0 + psi) - c f(y)
    void callable(const double* x, double t, double* out) {
        pr.rhs(x, t, out);
        for (int i = 0; i < n; ++i) tmp[i] = x[i] + psi_neg_y0[i];
        const double mc = -c;
        if (pr.model.has_mass) pr.mass_gemv(tmp.data(), t, mc, out);
        else for (int i = 0; i < n; ++i) out[i] = tmp[i] + mc * out[i];
    }

    // newton_iteration + NoLineSearch::take_optimal_step.  Returns true on convergence.
    bool newton_solve(Vec& xn, double t, const Vec& error_y) {
        return newton_solve_with([this](const double* x, double tt, double* out) { callable(x, tt, out); }, xn, t, error_y);
    }
    template <class F>
    bool newton_solve_with(F&& residual, Vec& xn, double t, const Vec& error_y) {
        convergence.reset();
        for (int it = 0; it < convergence.max_iter; ++it) {
            residual(xn.data(), t, newton_tmp.data());
            if (!lu.solve(newton_tmp.data())) return false;     // LuSolveFailed
            for (int i = 0; i < n; ++i) xn[i] -= newton_tmp[i];
            double norm = convergence.norm(newton_tmp.data(), error_y.data());
            ConvStatus s = convergence.check_new_iteration(norm);
            if (s == CONVERGED) return true;
            if (s == DIVERGED) return false;
        }
        return false;                                           // NewtonMaxIterations
    }

    // bdf.rs:694-731.  ret: 0 = nothing, 1 = TstopReached, <0 = error code negated
    int handle_tstop(double ts) {
        double troundoff = 100.0 * std::numeric_limits<double>::epsilon() * (std::fabs(t_) + std::fabs(h_));
        if (std::fabs(t_ - ts) <= troundoff) { has_tstop = false; return 1; }
        if ((h_ > 0.0 && ts < t_ - troundoff) || (h_ < 0.0 && ts > t_ + troundoff)) {
            has_tstop = false;
            return -ST_STOP_TIME_BEFORE_CURRENT;
        }
        if ((h_ > 0.0 && t_ + h_ > ts + troundoff) || (h_ < 0.0 && t_ + h_ < ts - troundoff)) {
            double factor = (ts - t_) / h_;
            (void)update_step_size(factor, nullptr);            // "step size too small" ignored
        }
        return 0;
    }

    // bdf.rs:812-869 (state part only): NB error_const2[order - 1]
    double error_control() const {
        double err = squared_norm(y_delta.data(), y_.data(), pr.atol.data(), pr.rtol, n) * error_const2[order - 1];
        err = std::max(0.0, err);
        if (pr.sens_error_control)                          // NB error_const2[order] for the sensitivities (bdf.rs:844-858)
            for (int i = 0; i < ns; ++i)
                err = std::max(err, squared_norm(s_deltas[i].data(), s_[i].data(), pr.sens_atol.data(), pr.sens_rtol, n) * error_const2[order]);
        return err;
    }
    // bdf.rs:871-932
    double predict_error_control(int ord) const {
        double err = squared_norm(D(ord + 1), y_.data(), pr.atol.data(), pr.rtol, n) * error_const2[ord];
        err = std::max(0.0, err);
        if (pr.sens_error_control)                          // bdf.rs:908-919
            for (int i = 0; i < ns; ++i)
                err = std::max(err, squared_norm(sdiff[i].data() + (size_t)(ord + 1) * n, s_[i].data(), pr.sens_atol.data(), pr.sens_rtol, n)
                                        * error_const2[ord]);
        return err;
    }
    // runge_kutta.rs:1313-1335
    double pi_controller_raw(double error_norm, int eff_order) const {
        double order_f = (double)eff_order;
        double ki = pr.opt.pi_control_integral / order_f;
        if (pr.opt.pi_control_proportional == 0.0 || !has_prev_error) return pr.math.pow(error_norm, -ki);
        double kp = pr.opt.pi_control_proportional / order_f;
        return pr.math.pow(error_norm, -(ki + kp)) * pr.math.pow(prev_error_norm, kp);
    }

    // bdf.rs:1277-1589
    StopReason step(int* err) override {
        double safety = 0.0, error_norm = 0.0;
        const int64_t old_num_error_test_failures = statistics.v[S_ERROR_TEST_FAILS];
        bool convergence_fail = false;
        double new_h = 0.0;
        if (is_state_modified) {                           // bdf.rs:1291-1318
            if (pr.model.nroots > 0) root_finder.init(pr, y_.data(), t_);
            // initialise_to_first_order (bdf.rs:733-763, bdf_state.rs:72-78): order 1, D[:, 0] = y, D[:, 1] = h dy (the
            // higher columns keep what they held)
            n_equal_steps = 0;
            order = 1;
            for (int i = 0; i < n; ++i) { D(0)[i] = y_[i]; D(1)[i] = dy_[i] * h_; }
            u = compute_r(1, 1.0);
            is_state_modified = false;
            if (ns > 0) { *err = ST_BAD_ARG; return STEP_ERROR; }   // resets with sensitivities (bdf.rs:1022-1078) are not restated
            const double c_new = h_ * alpha[order];
            set_c(h_, alpha[order]);
            jacobian_updates(c_new, STEP_SUCCESS);
            has_prev_error = false;
            if (has_tstop) {
                int e = set_stop_time(tstop);
                if (e) { *err = e; return STEP_ERROR; }
            }
        }
        predict_forward();
        while (true) {
            const int ord = order;
            y_delta = y_predict;
            bool ok = newton_solve(y_delta, t_predict, y_predict);
            statistics.v[S_NL_ITERS] += convergence.niter;
            if (ok) {
                for (int i = 0; i < n; ++i) y_delta[i] -= y_predict[i];
                if (ns > 0 && !sensitivity_solve(t_predict)) ok = false;     // SensitivitySolveFailed (bdf.rs:1355-1361)
            }
            if (!ok) {
                statistics.v[S_NL_FAILS] += 1;
                if (statistics.v[S_NL_FAILS] > pr.opt.max_nonlinear_solver_failures) {
                    *err = ST_TOO_MANY_NONLINEAR_FAILURES; return STEP_ERROR;
                }
                if (convergence_fail) {
                    has_prev_error = false;
                    int e = update_step_size(0.3, &new_h);
                    if (e) { *err = e; return STEP_ERROR; }
                    jacobian_updates(new_h * alpha[ord], SECOND_CONVERGENCE_FAIL);
                    predict_forward();
                } else {
                    has_prev_error = false;
                    jacobian_updates(h_ * alpha[ord], FIRST_CONVERGENCE_FAIL);
                    convergence_fail = true;
                }
                continue;
            }
            error_norm = error_control();
            double maxiter = (double)convergence.max_iter;
            double niter = (double)convergence.niter;
            safety = 0.9 * (2.0 * maxiter + 1.0) / (2.0 * maxiter + niter);
            if (error_norm <= 1.0) break;
            // rejected: has_prev_error is still the previous accepted step's here (P-only by default)
            double factor = safety * pi_controller_raw(error_norm, ord + 1);
            has_prev_error = false;
            if (factor < pr.opt.min_timestep_shrink) factor = pr.opt.min_timestep_shrink;
            int e = update_step_size(factor, &new_h);
            if (e) { *err = e; return STEP_ERROR; }
            jacobian_updates(new_h * alpha[ord], ERROR_TEST_FAIL);
            predict_forward();
            statistics.v[S_ERROR_TEST_FAILS] += 1;
            if (statistics.v[S_ERROR_TEST_FAILS] - old_num_error_test_failures >= pr.opt.max_error_test_failures) {
                *err = ST_TOO_MANY_ERROR_TEST_FAILURES; return STEP_ERROR;
            }
        }
        // accepted
        update_diff(order, y_delta);
        for (int i = 0; i < ns; ++i) update_diff_of(sdiff[i], order, s_deltas[i]);   // bdf.rs:629-633
        for (int i = 0; i < n; ++i) y_[i] = y_predict[i];       // Q1: the PREDICTOR
        t_ = t_predict;
        {
            double inv_h = 1.0 / h_;
            for (int i = 0; i < n; ++i) dy_[i] = D(1)[i] * inv_h;
        }
        statistics.v[S_STEPS] += 1;
        jacobian_update.step();
        has_prev_error = true; prev_error_norm = error_norm;
        n_equal_steps += 1;
        if (n_equal_steps > order) {
            const int ord = order;
            const double inf = std::numeric_limits<double>::infinity();
            double error_m_norm = ord > 1 ? predict_error_control(ord - 1) : inf;
            double error_p_norm = ord < MAX_ORDER ? predict_error_control(ord + 1) : inf;
            double factors[3] = {
                pi_controller_raw(error_m_norm, ord),
                pi_controller_raw(error_norm, ord + 1),
                pi_controller_raw(error_p_norm, ord + 2),
            };
            // Iterator::max_by keeps the LAST maximum
            int max_index = 0;
            for (int i = 1; i < 3; ++i) if (!(factors[max_index] > factors[i])) max_index = i;
            int new_order = ord + (max_index - 1);
            order = new_order;
            if (max_index != 1) u = compute_r(new_order, 1.0);
            double factor = safety * factors[max_index];
            if (factor > pr.opt.max_timestep_growth) factor = pr.opt.max_timestep_growth;
            if (factor < pr.opt.min_timestep_shrink) factor = pr.opt.min_timestep_shrink;
            if (factor >= pr.opt.min_timestep_growth || factor <= pr.opt.max_timestep_shrink
                || max_index == 0 || max_index == 2) {
                int e = update_step_size(factor, &new_h);
                if (e) { *err = e; return STEP_ERROR; }
                jacobian_updates(new_h * alpha[new_order], STEP_SUCCESS);
            }
        }
        // check for a root within the accepted step (bdf.rs:1566-1579)
        if (pr.model.nroots > 0) {
            auto interp = [this](double tq, double* yq) { return interpolate(tq, yq); };
            if (root_finder.check_root(pr, interp, y_.data(), t_, &root_t_, &root_idx_)) return ROOT_FOUND;
        }
        if (has_tstop) {
            int r = handle_tstop(tstop);
            if (r == 1) return TSTOP_REACHED;
            // handle_tstop(...).unwrap(): an Err here would panic in the reference
            if (r < 0) { *err = -r; return STEP_ERROR; }
        }
        return INTERNAL_TIMESTEP;
    }

    // bdf.rs:1591-1599
    int set_stop_time(double ts) override {
        has_tstop = true; tstop = ts;
        int r = handle_tstop(ts);
        if (r == 1) { has_tstop = false; return ST_STOP_TIME_AT_CURRENT; }
        if (r < 0) return -r;
        return ST_OK;
    }

    // bdf.rs:1080-1106 + interpolate_from_diff :767-782
    int interpolate(double t, double* y) const override {
        bool is_forward = h_ > 0.0;
        if ((is_forward && t > t_) || (!is_forward && t < t_)) return ST_INTERPOLATION_TIME_AFTER_CURRENT;
        double time_factor = 1.0;
        for (int i = 0; i < n; ++i) y[i] = D(0)[i];
        for (int j = 0; j < order; ++j) {
            double j_t = (double)j;
            time_factor *= (t - (t_ - h_ * j_t)) / (h_ * (1.0 + j_t));
            for (int i = 0; i < n; ++i) y[i] = time_factor * D(j + 1)[i] + y[i];
        }
        return ST_OK;
    }

    // bdf.rs:1162-1215: interpolate_from_diff on every sdiff
    int interpolate_sens(double t, double* out) const override {
        if (ns == 0) return ST_BAD_ARG;
        bool is_forward = h_ > 0.0;
        if ((is_forward && t > t_) || (!is_forward && t < t_)) return ST_INTERPOLATION_TIME_AFTER_CURRENT;
        for (int q = 0; q < ns; ++q) {
            const double* d = sdiff[q].data();
            double* y = out + (size_t)q * n;
            double time_factor = 1.0;
            for (int i = 0; i < n; ++i) y[i] = d[i];
            for (int j = 0; j < order; ++j) {
                double j_t = (double)j;
                time_factor *= (t - (t_ - h_ * j_t)) / (h_ * (1.0 + j_t));
                for (int i = 0; i < n; ++i) y[i] = time_factor * d[(size_t)(j + 1) * n + i] + y[i];
            }
        }
        return ST_OK;
    }

    double root_t() const override { return root_t_; }
    int root_index() const override { return root_idx_; }
    // bdf.rs:1228-1262 (is_state_modified is false after a step; no integrate_out, no sensitivities)
    int state_mut_back(double t) override {
        if (is_state_modified) return t == t_ ? ST_OK : ST_INTERPOLATION_TIME_AFTER_CURRENT;
        const bool is_forward = h_ > 0.0;
        if ((is_forward && t > t_) || (!is_forward && t < t_)) return ST_INTERPOLATION_TIME_AFTER_CURRENT;
        Vec ynew(n);
        int e = interpolate(t, ynew.data());
        if (e) return e;
        y_ = ynew;
        // interpolate_derivative_from_diff (bdf.rs:788-810): the guess apply_reset_with_mass's set_consistent starts from
        {
            double pi = 1.0, d_pi = 0.0;
            for (int i = 0; i < n; ++i) dy_[i] = 0.0;
            for (int j = 0; j < order; ++j) {
                const double j_t = (double)j;
                const double denom = h_ * (1.0 + j_t);
                const double w = (t - (t_ - h_ * j_t)) / denom;
                const double dw = 1.0 / denom;
                const double new_d_pi = d_pi * w + pi * dw;
                pi *= w;
                d_pi = new_d_pi;
                for (int i = 0; i < n; ++i) dy_[i] = d_pi * D(j + 1)[i] + dy_[i];
            }
        }
        t_ = t;
        is_state_modified = true;
        return ST_OK;
    }
    // method.rs:175-181 -> state.rs:246-270 through state_mut() (no mass matrix: dy = f(y, t))
    int apply_reset() override {
        if (!pr.model.reset) return ST_BAD_ARG;
        is_state_modified = true;
        Vec ynew(n);
        pr.model.reset(y_.data(), pr.p.data(), t_, ynew.data());
        y_ = ynew;
        if (pr.model.has_mass) {
            // state.apply_reset_with_mass (state.rs:279-306): set_consistent with a Newton solver WITHOUT line search, from the
            // reset y and the dy interpolated at the root; InitOp and the mass matrix are evaluated at problem.t0 (state.rs:114-119)
            return consistent_solve(pr, [this](const double* x, double t, double* out) { pr.rhs(x, t, out); },
                                    [this](const double* x, double t, double* J) { pr.jacobian(x, t, J); }, y_, dy_, nullptr, true, 0);
        }
        pr.rhs(y_.data(), t_, dy_.data());
        return ST_OK;
    }

    int residual_known_answer(double cc, double, const double* vec, const double* x, double t, double* F, double* Aout) override {
        c = cc;                                                     // set_c_direct
        for (int i = 0; i < n; ++i) psi_neg_y0[i] = vec[i];         // set_psi_neg_y0_direct
        callable(x, t, F);
        jacobian_is_stale = true;
        reset_jacobian(x, t);
        for (size_t q = 0; q < (size_t)n * n; ++q) Aout[q] = A[q];
        return ST_OK;
    }

    double t() const override { return t_; }
    double h() const override { return h_; }
    int cur_order() const override { return order; }
    const double* y() const override { return y_.data(); }
    const Stats& stats() const override { return statistics; }
};

}  // namespace

Method* new_bdf(const Problem& pr, int* err) {
    Bdf* b = new Bdf(pr);
    *err = b->construct();
    if (*err) { delete b; return nullptr; }
    return b;
}

}  // namespace orc
