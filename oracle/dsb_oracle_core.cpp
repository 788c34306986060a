// oracle/dsb_oracle_core.cpp -- TEST INFRASTRUCTURE (see dsb_oracle.hpp).
// Linear algebra primitives, Newton/convergence logic, Jacobian assembly, consistent
// initialisation and initial step size, restated from the reference (paths relative to
// /root/reference/crates).
#include "dsb_oracle.hpp"

#include <algorithm>

namespace orc {

namespace {
struct ModelMaker {
    Model* out;
    template <class M> void operator()() { *out = make_model<M>(); }
};
}  // namespace

// user equation sets loaded at run time (orc_load_model_plugin, dsb_oracle_capi.cpp): ids ORC_PLUGIN_ID0, +1, ...
static std::vector<Model>& plugin_models() { static std::vector<Model> v; return v; }
int register_plugin_model(const Model& m) { plugin_models().push_back(m); return ORC_PLUGIN_ID0 + (int)plugin_models().size() - 1; }

bool model_by_id(int id, Model* out) {
    if (id >= ORC_PLUGIN_ID0) {
        const size_t k = (size_t)(id - ORC_PLUGIN_ID0);
        if (k >= plugin_models().size()) return false;
        *out = plugin_models()[k];
        return true;
    }
    ModelMaker f{out};
    return dsb_dispatch_model(id, f);
}

// diffsol-la/src/vector/nalgebra_serial.rs:395-408 -- sequential sum, MEAN of squares, no sqrt
double squared_norm(const double* x, const double* y, const double* atol, double rtol, int n) {
    double acc = 0.0;
    for (int i = 0; i < n; ++i) {
        double term = x[i] / (std::fabs(y[i]) * rtol + atol[i]);
        acc += term * term;
    }
    return acc / (double)n;
}

// ---- Problem -------------------------------------------------------------------------------------

// jacobian/coloring.rs:27-47 (nonzeros2graph: columns that share a row are adjacent) + jacobian/greedy_coloring.rs:14-34
// (color_graph_greedy): colour (1-based) of every column
std::vector<int> greedy_coloring(const std::vector<std::pair<int, int>>& non_zeros, int nn) {
    std::vector<std::vector<int>> cols_by_rows(nn);
    for (auto& ij : non_zeros) cols_by_rows[ij.first].push_back(ij.second);
    std::vector<std::vector<int>> adj(nn);
    for (auto& ij : non_zeros)
        for (int next_col : cols_by_rows[ij.first])
            if (next_col < ij.second) { adj[ij.second].push_back(next_col); adj[next_col].push_back(ij.second); }
    std::vector<int> result(nn, 0);
    if (nn > 0) result[0] = 1;
    std::vector<char> available(nn, 0);
    for (int ii = 1; ii < nn; ++ii) {
        for (int j : adj[ii]) if (result[j] != 0) available[result[j] - 1] = 1;
        for (int i = 0; i < nn; ++i) if (!available[i]) { result[ii] = i + 1; break; }
        std::fill(available.begin(), available.end(), 0);
    }
    return result;
}

// jacobian/mod.rs:16-48 (NaN probe, one jac_mul per column, counted in OpStatistics),
// jacobian/coloring.rs:27-47 (graph), jacobian/greedy_coloring.rs:14-34, jacobian/mod.rs:178-214
void Problem::build_coloring() {
    const int nn = n();
    Vec y0(nn), v(nn, 0.0), col(nn, 0.0);
    model.init(p.data(), t0, y0.data());
    non_zeros.clear();
    for (int j = 0; j < nn; ++j) {
        v[j] = std::numeric_limits<double>::quiet_NaN();
        jac_mul(y0.data(), t0, v.data(), col.data());
        for (int i = 0; i < nn; ++i)
            if (std::isnan(col[i])) non_zeros.push_back({i, j});
        std::fill(col.begin(), col.end(), 0.0);
        v[j] = 0.0;
    }
    const std::vector<int> result = greedy_coloring(non_zeros, nn);
    int max_color = 0;
    for (int c : result) max_color = std::max(max_color, c);
    color_inputs.clear(); color_entries.clear();
    for (int c = 1; c <= max_color; ++c) {
        std::vector<int> inputs; std::vector<std::pair<int, int>> entries;
        for (auto& ij : non_zeros) if (result[ij.second] == c) {
            entries.push_back(ij);
            if (std::find(inputs.begin(), inputs.end(), ij.second) == inputs.end()) inputs.push_back(ij.second);
        }
        color_inputs.push_back(inputs); color_entries.push_back(entries);
    }
}

// op/closure.rs:140-147 -> jacobian/mod.rs:236-256 or op/nonlinear_op.rs:211-220
void Problem::jacobian(const double* x, double t, double* J) const {
    const int nn = n();
    ++n_matrix_evals;
    Vec v(nn, 0.0), col(nn, 0.0);
    if (use_coloring) {
        // the matrix keeps whatever it held outside the non-zero pattern; it is allocated zeroed
        // and only pattern entries are ever written, so zero-fill is equivalent.
        std::fill(J, J + (size_t)nn * nn, 0.0);
        for (size_t c = 0; c < color_inputs.size(); ++c) {
            for (int j : color_inputs[c]) v[j] = 1.0;
            jac_mul(x, t, v.data(), col.data());
            for (auto& ij : color_entries[c]) J[(size_t)ij.second * nn + ij.first] = col[ij.first];
            for (int j : color_inputs[c]) v[j] = 0.0;
        }
    } else {
        for (int j = 0; j < nn; ++j) {
            v[j] = 1.0;
            jac_mul(x, t, v.data(), col.data());
            for (int i = 0; i < nn; ++i) J[(size_t)j * nn + i] = col[i];
            v[j] = 0.0;
        }
    }
}

void Problem::mass_matrix(double t, double* M) const {
    const int nn = n();
    if (!model.has_mass) {   // op/bdf.rs:141-143: identity from_diagonal
        std::fill(M, M + (size_t)nn * nn, 0.0);
        for (int i = 0; i < nn; ++i) M[(size_t)i * nn + i] = 1.0;
        return;
    }
    Vec v(nn, 0.0), col(nn, 0.0);
    for (int j = 0; j < nn; ++j) {
        v[j] = 1.0;
        mass_gemv(v.data(), t, 0.0, col.data());     // LinearOp::call_inplace: beta = 0
        for (int i = 0; i < nn; ++i) M[(size_t)j * nn + i] = col[i];
        v[j] = 0.0;
    }
}

// ---- nalgebra LU ---------------------------------------------------------------------------------
// nalgebra-0.35 src/linalg/lu.rs `LU::new`, `gauss_step`, `gauss_step_swap`; pivot = icamax (first
// strict maximum of |.|); sub-column scaled by the RECIPROCAL of the pivot; rank-1 update column by
// column as axpy(-pivot_row[k], coeffs, 1).
// EXPERIMENT SWITCH (tools/dmma_rounding_check.py): the trailing updates of the factorisation with a FUSED multiply-add,
// i.e. the arithmetic an FP64 tensor-core (DMMA) trailing update would perform.  The reference path is unfused.
int g_fused_lu_updates = 0;

void DenseLU::factor(const double* A, int n_) {
    n = n_;
    lu.assign(A, A + (size_t)n * n);
    perm.clear();
    auto at = [&](int i, int j) -> double& { return lu[(size_t)j * n + i]; };
    for (int i = 0; i < n; ++i) {
        int piv = i; double the_max = std::fabs(at(i, i));
        for (int r = i + 1; r < n; ++r) { double val = std::fabs(at(r, i)); if (val > the_max) { the_max = val; piv = r; } }
        double diag = at(piv, i);
        if (diag == 0.0) continue;                   // no non-zero entries on this column
        if (piv != i) {
            perm.push_back({i, piv});
            for (int c = 0; c < i; ++c) std::swap(at(i, c), at(piv, c));     // columns_range_mut(..i).swap_rows
            // gauss_step_swap
            double inv_diag = 1.0 / diag;
            std::swap(at(i, i), at(piv, i));
            for (int r = i + 1; r < n; ++r) at(r, i) *= inv_diag;
            for (int k = i + 1; k < n; ++k) {
                std::swap(at(i, k), at(piv, k));
                double mpk = -at(i, k);
                if (g_fused_lu_updates) { for (int r = i + 1; r < n; ++r) at(r, k) = __builtin_fma(mpk, at(r, i), at(r, k)); }
                else for (int r = i + 1; r < n; ++r) at(r, k) = mpk * at(r, i) + at(r, k);
            }
        } else {
            double inv_diag = 1.0 / diag;
            for (int r = i + 1; r < n; ++r) at(r, i) *= inv_diag;
            for (int k = i + 1; k < n; ++k) {
                double mpk = -at(i, k);
                if (g_fused_lu_updates) { for (int r = i + 1; r < n; ++r) at(r, k) = __builtin_fma(mpk, at(r, i), at(r, k)); }
                else for (int r = i + 1; r < n; ++r) at(r, k) = mpk * at(r, i) + at(r, k);
            }
        }
    }
}

// nalgebra `LU::solve_mut`: permute rows, unit-lower forward substitution
// (solve_lower_triangular_with_diag_mut(b, 1): coeff = b[i] / 1), upper back substitution
// (coeff = b[i] / U[i,i]; zero diagonal => false), both in column-axpy form.
bool DenseLU::solve(double* b) const {
    for (auto& pq : perm) std::swap(b[pq.first], b[pq.second]);
    auto at = [&](int i, int j) -> double { return lu[(size_t)j * n + i]; };
    for (int i = 0; i + 1 < n; ++i) {
        double coeff = b[i] / 1.0;
        double mc = -coeff;
        for (int r = i + 1; r < n; ++r) b[r] = mc * at(r, i) + b[r];
    }
    for (int i = n - 1; i >= 0; --i) {
        double diag = at(i, i);
        if (diag == 0.0) return false;
        double coeff = b[i] / diag;
        b[i] = coeff;
        double mc = -coeff;
        for (int r = 0; r < i; ++r) b[r] = mc * at(r, i) + b[r];
    }
    return true;
}

// ---- Convergence ---------------------------------------------------------------------------------
void Convergence::init(double rtol_, const double* atol_, int n_, double tol_, const Math* m) {
    rtol = rtol_; atol = atol_; n = n_; tol = tol_; max_iter = 10; niter = 0; has_old_norm = false;
    math = m; eta = math->pow(20.0, 1.25);
}

// diffsol-nl/src/convergence.rs:68-131
ConvStatus Convergence::check_norm(double norm) {
    niter += 1;
    if (has_old_norm) {
        double rate = math->pow(norm / old_norm, 1.0 / (double)(niter - 1));
        if (rate > 0.9) return DIVERGED;
        if (math->powi(rate, max_iter - niter) / (1.0 - rate) * norm > tol) return DIVERGED;
        eta = rate / (1.0 - rate);
    } else {
        double min_eta = 1e4 * std::numeric_limits<double>::epsilon();
        if (eta < min_eta) eta = min_eta;
        eta = math->pow(eta, 0.8);
    }
    if (eta * norm < tol) return CONVERGED;
    return CONTINUE;
}

// diffsol-nl/src/convergence.rs:133-139 -- old_norm is frozen at the FIRST iteration's norm
ConvStatus Convergence::check_new_iteration(double norm) {
    ConvStatus s = check_norm(norm);
    if (niter == 1) { has_old_norm = true; old_norm = norm; }
    return s;
}

// ---- JacobianUpdate --------------------------------------------------------------------------------
void JacobianUpdate::init(const Options& o, double h_at_last) {
    steps_since_jacobian_eval = 0; steps_since_rhs_jacobian_eval = 0;
    h_at_last_jacobian_update = h_at_last;
    threshold_to_update_jacobian = o.threshold_to_update_jacobian;
    threshold_to_update_rhs_jacobian = o.threshold_to_update_rhs_jacobian;
    update_jacobian_after_steps = o.update_jacobian_after_steps;
    update_rhs_jacobian_after_steps = o.update_rhs_jacobian_after_steps;
}
bool JacobianUpdate::check_jacobian_update(double h, SolverState s) const {
    if (s == STEP_SUCCESS)
        return steps_since_jacobian_eval >= update_jacobian_after_steps
               || std::fabs(h / h_at_last_jacobian_update - 1.0) > threshold_to_update_jacobian;
    return true;
}
bool JacobianUpdate::check_rhs_jacobian_update(double h, SolverState s) const {
    switch (s) {
        case STEP_SUCCESS: return steps_since_rhs_jacobian_eval >= update_rhs_jacobian_after_steps;
        case FIRST_CONVERGENCE_FAIL:
            return std::fabs(h / h_at_last_jacobian_update - 1.0) < threshold_to_update_rhs_jacobian;
        case SECOND_CONVERGENCE_FAIL: return steps_since_rhs_jacobian_eval > 0;
        case ERROR_TEST_FAIL: return false;
        case CHECKPOINT: return true;
    }
    return false;
}

void Stats::record_linear_solver_setup(SolverState s) {
    v[S_SETUPS] += 1;
    switch (s) {
        case CHECKPOINT: v[S_SETUPS_CHECKPOINT] += 1; break;
        case FIRST_CONVERGENCE_FAIL: v[S_SETUPS_FIRST_FAIL] += 1; break;
        case SECOND_CONVERGENCE_FAIL: v[S_SETUPS_SECOND_FAIL] += 1; break;
        case ERROR_TEST_FAIL: v[S_SETUPS_ERROR_FAIL] += 1; break;
        case STEP_SUCCESS: v[S_SETUPS_SUCCESS] += 1; break;
    }
}

// ---- consistent initialisation (ode_solver/state.rs:84-162, op/init.rs:14-131) -------------------
// Newton with BacktrackingLineSearch (diffsol-nl/src/line_search.rs:115-201) on
//   F(du, v) = -M_u du + f(u, v) ; g(u, v)      unknown x = (du at differential idx, v at algebraic idx)
// The same solve serves the state (set_consistent, state.rs:84-162: eq_rhs / eq_jacobian = the equations' own, a fresh
// Convergence, dy zeroed at the algebraic rows afterwards) and every sensitivity vector (set_consistent_augmented,
// state.rs:167-238: eq_rhs = SensRhs::call, eq_jacobian = SensRhs::jacobian_inplace, ONE Convergence for all parameters,
// ds kept at the algebraic rows).
int consistent_solve(const Problem& pr, const std::function<void(const double*, double, double*)>& eq_rhs,
                     const std::function<void(const double*, double, double*)>& eq_jacobian, Vec& y_state, Vec& dy_state,
                     Convergence* shared_conv, bool zero_dv, int use_linesearch) {
    const bool linesearch = use_linesearch < 0 ? pr.opt.ic_use_linesearch : use_linesearch != 0;
    const int n = pr.n();
    if (!pr.model.has_mass) return ST_OK;
    Vec M((size_t)n * n);
    pr.mass_matrix(pr.t0, M.data());
    std::vector<char> is_alg(n, 0);
    int nalg = 0;
    for (int i = 0; i < n; ++i) if (M[(size_t)i * n + i] == 0.0) { is_alg[i] = 1; ++nalg; }
    if (nalg == 0) return ST_OK;
    struct { Vec& y; Vec& dy; } state{y_state, dy_state};
    auto* st = &state;

    // InitOp::new: rhs_jac at (y0, t0); jac = (-M_u | df/dv ; 0 | dg/dv); neg_mass = (-M_u | 0 ; 0 | 0)
    Vec rhs_jac((size_t)n * n), jac((size_t)n * n, 0.0), neg_mass((size_t)n * n, 0.0);
    eq_jacobian(st->y.data(), pr.t0, rhs_jac.data());
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < n; ++i) {
            size_t ij = (size_t)j * n + i;
            if (!is_alg[j]) {
                if (!is_alg[i]) { double m_u = M[ij] * -1.0; jac[ij] = m_u; neg_mass[ij] = m_u; }
            } else {
                jac[ij] = rhs_jac[ij];
            }
        }
    Vec y0 = st->y;   // InitOp.y0
    auto fun = [&](const Vec& x, Vec& out) {
        for (int i = 0; i < n; ++i) if (is_alg[i]) y0[i] = x[i];
        eq_rhs(y0.data(), pr.t0, out.data());
        // neg_mass.gemv(1, x, 1, out): column sweep, out = (1 * col_j) * x_j + out
        for (int j = 0; j < n; ++j)
            for (int i = 0; i < n; ++i) out[i] = neg_mass[(size_t)j * n + i] * x[j] + out[i];
    };
    DenseLU lu;
    Vec y_tmp = st->dy;
    for (int i = 0; i < n; ++i) if (is_alg[i]) y_tmp[i] = st->y[i];
    Vec yerr = y_tmp;
    Convergence own_conv;
    if (!shared_conv) {
        own_conv.init(pr.rtol, pr.atol.data(), n, pr.opt.nonlinear_solver_tolerance, &pr.math);
        own_conv.max_iter = pr.opt.ic_max_newton_iterations;
    }
    Convergence& conv = shared_conv ? *shared_conv : own_conv;

    const double tau = pr.opt.ic_step_reduction_factor, c_armijo = pr.opt.ic_armijo_constant;
    const double steptol = pr.math.pow(std::numeric_limits<double>::epsilon(), 2.0 / 3.0);
    const int ls_max_iter = pr.opt.ic_max_linesearch_iterations;

    bool ok = false;
    for (int setup = 0; setup < pr.opt.ic_max_linear_solver_setups; ++setup) {
        lu.factor(jac.data(), n);                    // reset_jacobian: InitOp's Jacobian is constant
        // newton_iteration (diffsol-nl/src/newton.rs:13-36)
        conv.reset();
        Vec delta(n, 0.0), x0(n), delta0(n);
        double ls_norm = 1.0;
        int result = -1;   // 0 ok, 1 max iterations, 2 other error
        for (int it = 0; it < conv.max_iter && result < 0; ++it) {
            ConvStatus res = CONTINUE;
            bool have_res = false;
            if (linesearch) {
                if (conv.niter == 0) {
                    fun(y_tmp, delta);
                    if (!lu.solve(delta.data())) { result = 2; break; }
                    ls_norm = conv.norm(delta.data(), yerr.data());
                    if (conv.check_norm(ls_norm) == CONVERGED) {
                        for (int i = 0; i < n; ++i) y_tmp[i] -= delta[i];
                        res = CONVERGED; have_res = true;
                    }
                }
                if (!have_res) {
                    x0 = y_tmp; delta0 = delta;
                    const double norm = ls_norm;
                    const double phi0 = norm * norm * 0.5, two_phi0 = norm * norm;
                    const double min_alpha = steptol / norm;
                    double alpha = 1.0;
                    int ls_status = 1;   // 1 = max iterations
                    for (int i = 0; i < ls_max_iter; ++i) {
                        for (int q = 0; q < n; ++q) y_tmp[q] = (-alpha) * delta0[q] + y_tmp[q];
                        fun(y_tmp, delta);
                        if (!lu.solve(delta.data())) { ls_status = 2; break; }
                        double new_norm = conv.norm(delta.data(), yerr.data());
                        double phi1 = new_norm * new_norm * 0.5;
                        if (phi1 <= phi0 - c_armijo * alpha * two_phi0) {
                            ls_norm = new_norm;
                            res = conv.check_norm(new_norm); have_res = true; ls_status = 0;
                            break;
                        }
                        if (alpha < min_alpha) { ls_status = 2; break; }
                        alpha *= tau;
                        y_tmp = x0;
                    }
                    if (ls_status != 0) { result = 2; break; }
                }
            } else {
                // NoLineSearch (line_search.rs:48-69)
                fun(y_tmp, delta);
                if (!lu.solve(delta.data())) { result = 2; break; }
                for (int i = 0; i < n; ++i) y_tmp[i] -= delta[i];
                res = conv.check_new_iteration(conv.norm(delta.data(), yerr.data()));
            }
            if (res == CONVERGED) result = 0;
            else if (res == DIVERGED) result = 2;
        }
        if (result < 0) result = 1;                  // NewtonMaxIterations
        if (result == 0) { ok = true; break; }
        if (result == 2) return ST_INITIAL_CONDITION_DID_NOT_CONVERGE;
        yerr = y_tmp;
    }
    if (!ok) return ST_INITIAL_CONDITION_DID_NOT_CONVERGE;
    // InitOp::scatter_soln (op/init.rs:77-82) [+ zero dv for the state (state.rs:155-160)]
    for (int i = 0; i < n; ++i) {
        if (is_alg[i]) { st->y[i] = y_tmp[i]; if (zero_dv) st->dy[i] = 0.0; }
        else st->dy[i] = y_tmp[i];
    }
    return ST_OK;
}
static int set_consistent(const Problem& pr, InitialState* st) {
    return consistent_solve(pr, [&pr](const double* x, double t, double* out) { pr.rhs(x, t, out); },
                            [&pr](const double* x, double t, double* J) { pr.jacobian(x, t, J); }, st->y, st->dy, nullptr, true);
}

// ode_solver/state.rs:1209-1277 (Hairer/Norsett/Wanner II.4.2)
static void set_step_size(const Problem& pr, int solver_order, InitialState* st) {
    const int n = pr.n();
    const double h0_user = pr.h0;
    const bool is_neg_h = h0_user < 0.0;
    const double* atol = pr.atol.data();
    const double rtol = pr.rtol;
    const Vec& y0 = st->y; const Vec& f0 = st->dy;
    double d0 = std::sqrt(squared_norm(y0.data(), y0.data(), atol, rtol, n));
    double d1 = std::sqrt(squared_norm(f0.data(), y0.data(), atol, rtol, n));
    double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * (d0 / d1);
    Vec y1(n), f1(n);
    if (is_neg_h) {
        for (int i = 0; i < n; ++i) y1[i] = f0[i] * (-h0) + y0[i];
        pr.rhs(y1.data(), st->t - h0, f1.data());
    } else {
        for (int i = 0; i < n; ++i) y1[i] = f0[i] * h0 + y0[i];
        pr.rhs(y1.data(), st->t + h0, f1.data());
    }
    Vec df(n);
    for (int i = 0; i < n; ++i) df[i] = f1[i] - f0[i];
    double d2 = std::sqrt(squared_norm(df.data(), y0.data(), atol, rtol, n)) / std::fabs(h0);
    double max_d = d2;
    if (max_d < d1) max_d = d1;
    double h1;
    if (max_d < 1e-15) {
        h1 = h0 * 1e-3;
        if (h1 < 1e-6) h1 = 1e-6;
    } else {
        h1 = pr.math.pow(0.01 / max_d, 1.0 / (1.0 + (double)solver_order));
    }
    double h = 100.0 * h0;
    if (h > h1) h = h1;
    if (is_neg_h) h = -h;
    st->h = h;
}

// ode_solver/state.rs:969-997 (new_and_consistent) and :1086-1124 (new_without_initialise)
int new_and_consistent(const Problem& pr, int solver_order, InitialState* st) {
    const int n = pr.n();
    st->t = pr.t0; st->h = pr.h0;
    st->y.assign(n, 0.0); st->dy.assign(n, 0.0);
    pr.model.init(pr.p.data(), pr.t0, st->y.data());
    pr.rhs(st->y.data(), pr.t0, st->dy.data());
    int err = set_consistent(pr, st);
    if (err) return err;
    set_step_size(pr, solver_order, st);
    return ST_OK;
}

// Vector::root_finding (diffsol-la/src/vector/nalgebra_serial.rs:484-504)
void RootFinder::root_finding(const Vec& g0v, const Vec& g1v, bool* found_root, int* imax) {
    double max_frac = 0.0;
    int max_frac_index = -1;
    bool found = false;
    for (size_t i = 0; i < g0v.size(); ++i) {
        const double g0 = g0v[i], g1 = g1v[i];
        if (g1 == 0.0) found = true;
        if (g0 * g1 < 0.0) {
            const double frac = std::fabs(g1 / (g1 - g0));
            if (frac > max_frac) { max_frac = frac; max_frac_index = (int)i; }
        }
    }
    *found_root = found; *imax = max_frac_index;
}

void RootFinder::init(const Problem& pr, const double* y, double t) {
    pr.model.root(y, pr.p.data(), t, g0.data());
    t0 = t;
}

// root.rs:60-160 (the modified secant / Illinois iteration of SUNDIALS)
bool RootFinder::check_root(const Problem& pr, const std::function<int(double, double*)>& interpolate, const double* y,
                            double t, double* t_root, int* idx) {
    const double* p = pr.p.data();
    pr.model.root(y, p, t, g1.data());
    bool rootfnd; int imax;
    root_finding(g0, g1, &rootfnd, &imax);
    if (imax < 0) {
        std::swap(g0, g1);
        t0 = t;
        if (rootfnd) {
            // find_zero_index (root.rs:44-58): the entry of smallest magnitude, first one on ties
            int min_idx = 0; double min_val = std::fabs(g0[0]);
            for (size_t i = 1; i < g0.size(); ++i) { const double v = std::fabs(g0[i]); if (v < min_val) { min_val = v; min_idx = (int)i; } }
            *t_root = t; *idx = min_idx;
            return true;
        }
        return false;
    }
    double alpha = 1.0;
    bool sign_change[2] = {false, true};
    int i = 0;
    double t1 = t, tl = t0;
    const double tol = 100.0 * 2.220446049250313e-16 * (std::fabs(t1) + std::fabs(t1 - tl));
    while (std::fabs(t1 - tl) > tol) {
        const double g1_val = g1[imax], g0_val = g0[imax];
        double t_mid = t1 - (t1 - tl) * g1_val / (g1_val - alpha * g0_val);
        if (std::fabs(t_mid - tl) < 0.5 * tol) {
            const double fracint = std::fabs(t1 - tl) / tol;
            const double fracsub = fracint > 5.0 ? 0.1 : 0.5 / fracint;
            t_mid = tl + fracsub * (t1 - tl);
        }
        if (std::fabs(t1 - t_mid) < 0.5 * tol) {
            const double fracint = std::fabs(t1 - tl) / tol;
            const double fracsub = fracint > 5.0 ? 0.1 : 0.5 / fracint;
            t_mid = t1 - fracsub * (t1 - tl);
        }
        interpolate(t_mid, ymid.data());                   // .unwrap() in the reference
        pr.model.root(ymid.data(), p, t_mid, gmid.data());
        bool rf; int im;
        root_finding(g0, gmid, &rf, &im);
        const bool lower = im >= 0;
        if (lower) {
            t1 = t_mid; imax = im; std::swap(g1, gmid);
        } else if (rf) {
            pr.model.root(y, p, t, g0.data());
            *t_root = t_mid; *idx = imax;
            return true;
        } else {
            tl = t_mid; std::swap(g0, gmid);
        }
        sign_change[i % 2] = lower;
        if (i >= 2) alpha = (sign_change[0] != sign_change[1]) ? 1.0 : (sign_change[0] ? 0.5 * alpha : 2.0 * alpha);
        ++i;
    }
    pr.model.root(y, p, t, g0.data());
    *t_root = t1; *idx = imax;
    return true;
}

// fn solve_dense (ode_solver/method.rs:721-818) + dense_write_out (:822-848) + the root column of
// OdeSolverMethod::solve_dense (:493-503); no reset, no checkpointing
int solve_dense(Method& s, const double* t_eval, int nt, int n, double* out, int* ncols, double* root_t, int* root_idx,
                const Problem* pr) {
    if (ncols) *ncols = nt;
    if (root_idx) *root_idx = -1;
    if (nt <= 0) return ST_BAD_ARG;
    const bool has_out = pr && pr->model.nout > 0;
    const int nrow = has_out ? pr->model.nout : n;
    Vec tmp_nstates(n);
    // dense_write_out (method.rs:822-848): interpolate, then the output function when the equations have one
    auto write_column = [&](double tq, int col) -> int {
        if (!has_out) return s.interpolate(tq, out + (size_t)col * nrow);
        int e = s.interpolate(tq, tmp_nstates.data());
        if (e) return e;
        pr->model.out(tmp_nstates.data(), pr->p.data(), tq, out + (size_t)col * nrow);
        return ST_OK;
    };
    int err = s.set_stop_time(t_eval[nt - 1]);
    if (err) return err;
    int col = 0;
    while (true) {
        StopReason r = s.step(&err);
        if (r == STEP_ERROR) return err;
        if (r == ROOT_FOUND) {
            const double tr = s.root_t();
            while (col < nt && t_eval[col] <= tr) {
                int e2 = write_column(t_eval[col], col);
                if (e2) return e2;
                ++col;
            }
            int e3 = s.state_mut_back(tr);
            if (e3) return e3;
            if (pr && pr->model.reset) {                    // has_reset (method.rs:783-797): reset, new stop time, go on
                int e4 = s.apply_reset();
                if (e4) return e4;
                if (s.t() < t_eval[nt - 1]) {
                    int e5 = s.set_stop_time(t_eval[nt - 1]);
                    if (e5) return e5;
                    continue;
                }
                break;                                      // TstopReached
            }
            if (col < nt) {                                 // write_state_out at the root, then resize_cols(col + 1)
                if (has_out) pr->model.out(s.y(), pr->p.data(), s.t(), out + (size_t)col * nrow);
                else for (int i = 0; i < n; ++i) out[(size_t)col * n + i] = s.y()[i];
                ++col;
            }
            if (ncols) *ncols = col;
            if (root_t) *root_t = tr;
            if (root_idx) *root_idx = s.root_index();
            return ST_OK;
        }
        while (col < nt && t_eval[col] <= s.t()) {
            int e2 = write_column(t_eval[col], col);
            if (e2) return e2;
            ++col;
        }
        if (r == TSTOP_REACHED) break;
    }
    // columns actually written: nt, unless the stop time was declared reached within its round-off tolerance just
    // below the last t_eval (the reference asserts col == t_eval.len() there, method.rs:771)
    if (ncols) *ncols = col;
    return ST_OK;
}

// fn solve (ode_solver/method.rs:881-961) + write_out (:965-1000) behind OdeSolverMethod::solve(final_time) (:227-258): one
// column per internal step -- (state.t, state.y), or out(state.y, state.t) for equations with an output function -- after
// the initial one; a root ends the solve with the state moved back to the root (no reset function here).
// ST_BAD_ARG when max_cols columns do not hold the run.
int solve_ragged(Method& s, double final_time, int n, const Problem* pr, int max_cols, double* ts, double* ys, int* ncols,
                 double* root_t, int* root_idx) {
    const bool has_out = pr && pr->model.nout > 0;
    const int nrow = has_out ? pr->model.nout : n;
    int col = 0;
    auto write_out = [&]() -> bool {
        if (col >= max_cols) return false;
        ts[col] = s.t();
        if (has_out) pr->model.out(s.y(), pr->p.data(), s.t(), ys + (size_t)col * nrow);
        else for (int i = 0; i < n; ++i) ys[(size_t)col * nrow + i] = s.y()[i];
        ++col;
        return true;
    };
    if (root_idx) *root_idx = -1;
    if (pr && pr->model.reset) return ST_BAD_ARG;
    if (!write_out()) return ST_BAD_ARG;
    int err = s.set_stop_time(final_time);
    while (!err) {
        StopReason r = s.step(&err);
        if (r == STEP_ERROR) break;
        if (r == ROOT_FOUND) {
            const double tr = s.root_t();
            err = s.state_mut_back(tr);
            if (!err && !write_out()) err = ST_BAD_ARG;
            if (root_t) *root_t = tr;
            if (root_idx) *root_idx = s.root_index();
            break;
        }
        if (!write_out()) { err = ST_BAD_ARG; break; }
        if (r == TSTOP_REACHED) break;
    }
    if (ncols) *ncols = col;
    return err;
}

}  // namespace orc
