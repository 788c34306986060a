// oracle/dsb_oracle_capi.cpp -- TEST INFRASTRUCTURE (see dsb_oracle.hpp).
// extern "C" entry points used through ctypes by tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs.
#include <dlfcn.h>

#include "dsb_oracle.hpp"

#include <atomic>
#include <memory>
#include <thread>

using namespace orc;

// Flat problem description; mirrors what OdeBuilder collects (builder.rs:112-140).
struct orc_problem_desc {
    int32_t model_id;
    int32_t method;        // 0 = Bdf, 1 = Sdirk(tr_bdf2), 2 = Sdirk(esdirk34)
    int32_t use_coloring;
    int32_t powmode;       // 0 = libm pow (reference-literal), 1 = dsb_pow (shared with CUDA)
    double rtol;
    double t0, h0;
    int32_t natol;         // 1 => broadcast (builder default), else == n
    int32_t has_options;   // 0 => OdeSolverOptions::default()
    double atol[64];
    // option overrides (only read when has_options != 0), same order as orc::Options
    int32_t max_nonlinear_solver_iterations, max_error_test_failures, max_nonlinear_solver_failures;
    int32_t update_jacobian_after_steps, update_rhs_jacobian_after_steps, pad0;
    double nonlinear_solver_tolerance, min_timestep;
    double max_timestep_growth, min_timestep_growth, max_timestep_shrink, min_timestep_shrink;
    double threshold_to_update_jacobian, threshold_to_update_rhs_jacobian;
    double pi_control_proportional, pi_control_integral;
    // forward sensitivities (problem.bdf_sens()): sens != 0 integrates one sensitivity vector per parameter; sens_natol = 0
    // leaves them out of the error test (builder.rs turn_off_sensitivities_error_control), 1 broadcasts, else == n
    int32_t sens, sens_natol;
    double sens_rtol;
    double sens_atol[64];
};

static int build_problem(const orc_problem_desc* d, const double* p, int np, Problem* pr) {
    if (!model_by_id(d->model_id, &pr->model)) return ST_BAD_ARG;
    const int n = pr->model.n;
    if (np != pr->model.np) return ST_BAD_ARG;
    pr->p.assign(p, p + np);
    pr->rtol = d->rtol; pr->t0 = d->t0; pr->h0 = d->h0;
    pr->atol.resize(n);
    if (d->natol == 1) for (int i = 0; i < n; ++i) pr->atol[i] = d->atol[0];
    else if (d->natol == n && n <= 64) for (int i = 0; i < n; ++i) pr->atol[i] = d->atol[i];
    else return ST_BAD_ARG;
    pr->math.powmode = d->powmode;
    if (d->has_options) {
        Options& o = pr->opt;
        o.max_nonlinear_solver_iterations = d->max_nonlinear_solver_iterations;
        o.max_error_test_failures = d->max_error_test_failures;
        o.max_nonlinear_solver_failures = d->max_nonlinear_solver_failures;
        o.update_jacobian_after_steps = d->update_jacobian_after_steps;
        o.update_rhs_jacobian_after_steps = d->update_rhs_jacobian_after_steps;
        o.nonlinear_solver_tolerance = d->nonlinear_solver_tolerance;
        o.min_timestep = d->min_timestep;
        o.max_timestep_growth = d->max_timestep_growth; o.min_timestep_growth = d->min_timestep_growth;
        o.max_timestep_shrink = d->max_timestep_shrink; o.min_timestep_shrink = d->min_timestep_shrink;
        o.threshold_to_update_jacobian = d->threshold_to_update_jacobian;
        o.threshold_to_update_rhs_jacobian = d->threshold_to_update_rhs_jacobian;
        o.pi_control_proportional = d->pi_control_proportional;
        o.pi_control_integral = d->pi_control_integral;
    }
    pr->use_coloring = d->use_coloring != 0;
    if (pr->use_coloring) pr->build_coloring();
    pr->sens = d->sens != 0;
    if (pr->sens) {
        if (!pr->model.sens_mul) return ST_BAD_ARG;
        pr->sens_error_control = d->sens_natol != 0;
        pr->sens_rtol = d->sens_rtol;
        pr->sens_atol.assign(n, 0.0);
        if (d->sens_natol == 1) for (int i = 0; i < n; ++i) pr->sens_atol[i] = d->sens_atol[0];
        else if (d->sens_natol == n && n <= 64) for (int i = 0; i < n; ++i) pr->sens_atol[i] = d->sens_atol[i];
        else if (d->sens_natol != 0) return ST_BAD_ARG;
    }
    return ST_OK;
}

static Method* make_method(const Problem& pr, int method, int* err) {
    switch (method) {
        case 0: return new_bdf(pr, err);
        case 1: return new_sdirk(pr, 0, err);
        case 2: return new_sdirk(pr, 1, err);
    }
    *err = ST_BAD_ARG;
    return nullptr;
}

static void export_stats(const Problem& pr, const Method* m, int64_t* stats) {
    if (!stats) return;
    for (int i = 0; i < S_COUNT; ++i) stats[i] = m ? m->stats().v[i] : 0;
    stats[S_RHS_CALLS] = pr.n_calls;
    stats[S_RHS_JAC_MULS] = pr.n_jac_muls;
    stats[S_RHS_MATRIX_EVALS] = pr.n_matrix_evals;
}

extern "C" {

int orc_model_dims(int model_id, int* n, int* np, int* has_mass) {
    Model m;
    if (!model_by_id(model_id, &m)) return ST_BAD_ARG;
    *n = m.n; *np = m.np; *has_mass = m.has_mass ? 1 : 0;
    return ST_OK;
}

// rows of the solve_dense result: the number of outputs of the model's out function, else the number of states
int orc_model_nout(int model_id) {
    Model m;
    if (!model_by_id(model_id, &m)) return -1;
    return m.ncols_out();
}
// root (event) functions of a model at (y, p, t): g[nroots]; returns the number of root functions
int orc_model_root(int model_id, const double* y, const double* p, double t, double* g) {
    Model m;
    if (!model_by_id(model_id, &m)) return -1;
    if (m.nroots > 0) m.root(y, p, t, g);
    return m.nroots;
}
// the shared deterministic elementary functions of csrc/dsb_math.h: which = 0 exp, 1 log, 2 tanh, 3 asinh
double orc_math(int which, double x) {
    switch (which) {
        case 0: return dsb_exp(x);
        case 1: return dsb_log(x);
        case 2: return dsb_tanh(x);
        default: return dsb_asinh(x);
    }
}

// `problem.bdf::<LS>()?.solve_dense(t_eval)`: out is n x nt column-major; fin = {t, h, order}; root (may be NULL) =
// {t_root, root index (-1: the integration did not stop on a root), number of columns written}
static int solve_dense_one(const orc_problem_desc* d, const double* p, int np, const double* t_eval, int nt,
                           double* out, int64_t* stats, double* fin, double* root) {
    Problem pr;
    int err = build_problem(d, p, np, &pr);
    if (err) return err;
    std::unique_ptr<Method> m(make_method(pr, d->method, &err));
    if (!m) { export_stats(pr, nullptr, stats); return err; }
    int ncols = nt, ridx = -1; double rt = 0.0;
    err = solve_dense(*m, t_eval, nt, pr.n(), out, &ncols, &rt, &ridx, &pr);
    export_stats(pr, m.get(), stats);
    if (fin) { fin[0] = m->t(); fin[1] = m->h(); fin[2] = (double)m->cur_order(); }
    if (root) { root[0] = rt; root[1] = (double)ridx; root[2] = (double)ncols; }
    return err;
}
int orc_solve_dense(const orc_problem_desc* d, const double* p, int np, const double* t_eval, int nt,
                    double* out, int64_t* stats, double* fin) {
    return solve_dense_one(d, p, np, t_eval, nt, out, stats, fin, nullptr);
}

// The reference's test harness `test_ode_solver(.., use_tstop = false)` (ode_solver/mod.rs:104-194):
// for each point: step while |t| < |t_point|, then interpolate(t_point); a step that reports RootFound ends the loop with
// interpolate(t_root) in that point's column (the later columns keep what the caller put there).  out is n x npts.
int orc_harness(const orc_problem_desc* d, const double* p, int np, const double* t_points, int npts,
                double* out, int64_t* stats, double* fin) {
    Problem pr;
    int err = build_problem(d, p, np, &pr);
    if (err) return err;
    std::unique_ptr<Method> m(make_method(pr, d->method, &err));
    if (!m) { export_stats(pr, nullptr, stats); return err; }
    const int n = pr.n();
    bool ended_on_root = false;
    for (int k = 0; k < npts && !err && !ended_on_root; ++k) {
        while (std::fabs(m->t()) < std::fabs(t_points[k])) {
            StopReason r = m->step(&err);
            if (r == STEP_ERROR) break;
            if (r == ROOT_FOUND) {
                // `if let RootFound(t, _) = method.step() { return method.interpolate(t) }` (ode_solver/mod.rs:134-141):
                // the state at the root takes the place of the point the loop was stepping towards, and the test ends
                err = m->interpolate(m->root_t(), out + (size_t)k * n);
                ended_on_root = true;
                break;
            }
        }
        if (err || ended_on_root) break;
        err = m->interpolate(t_points[k], out + (size_t)k * n);
    }
    export_stats(pr, m.get(), stats);
    if (fin) { fin[0] = m->t(); fin[1] = m->h(); fin[2] = (double)m->cur_order(); }
    return err;
}

// test_ode_solver(.., solve_for_sensitivities = true) (ode_solver/mod.rs:104-194): the same loop, plus interpolate_sens at
// every point.  out is n x npts, sens_out is n x np x npts (point-major, then parameter).
int orc_harness_sens(const orc_problem_desc* d, const double* p, int np, const double* t_points, int npts,
                     double* out, double* sens_out, int64_t* stats, double* fin) {
    Problem pr;
    int err = build_problem(d, p, np, &pr);
    if (err) return err;
    if (!pr.sens) return ST_BAD_ARG;
    std::unique_ptr<Method> m(make_method(pr, d->method, &err));
    if (!m) { export_stats(pr, nullptr, stats); return err; }
    const int n = pr.n();
    for (int k = 0; k < npts && !err; ++k) {
        while (std::fabs(m->t()) < std::fabs(t_points[k])) {
            StopReason r = m->step(&err);
            if (r == STEP_ERROR) break;
        }
        if (err) break;
        err = m->interpolate(t_points[k], out + (size_t)k * n);
        if (!err) err = m->interpolate_sens(t_points[k], sens_out + (size_t)k * n * np);
    }
    export_stats(pr, m.get(), stats);
    if (fin) { fin[0] = m->t(); fin[1] = m->h(); fin[2] = (double)m->cur_order(); }
    return err;
}

// OdeSolverMethod::solve(final_time) for every row of params (instance-major): instance b's columns at ts[b * max_cols ..),
// ys[b * max_cols * nrow ..) (nrow = the out function's outputs, else n), their number in ncols[b]; roots[b] = (t_root, index)
int orc_batch_solve_ragged(const orc_problem_desc* d, const double* params, int np, int64_t nbatch, double final_time, int max_cols,
                           int nthreads, double* ts, double* ys, int32_t* ncols, int64_t* stats, int32_t* status, double* roots) {
    Model mm;
    if (!model_by_id(d->model_id, &mm)) return ST_BAD_ARG;
    const int nrow = mm.ncols_out();
    int nt_use = nthreads > 0 ? nthreads : (int)std::thread::hardware_concurrency();
    if (nt_use < 1) nt_use = 1;
    std::atomic<int64_t> next(0);
    auto worker = [&]() {
        while (true) {
            const int64_t b0 = next.fetch_add(16);
            if (b0 >= nbatch) break;
            const int64_t b1 = b0 + 16 < nbatch ? b0 + 16 : nbatch;
            for (int64_t b = b0; b < b1; ++b) {
                Problem pr;
                int err = build_problem(d, params + (size_t)b * np, np, &pr);
                int nc = 0, ridx = -1; double rt = 0.0;
                std::unique_ptr<Method> m;
                if (!err) m.reset(make_method(pr, d->method, &err));
                if (m) err = solve_ragged(*m, final_time, pr.n(), &pr, max_cols, ts + (size_t)b * max_cols, ys + (size_t)b * max_cols * nrow,
                                          &nc, &rt, &ridx);
                export_stats(pr, m.get(), stats + (size_t)b * S_COUNT);
                ncols[b] = nc; status[b] = err;
                if (roots) { roots[b * 2] = rt; roots[b * 2 + 1] = (double)ridx; }
            }
        }
    };
    if (nt_use == 1) { worker(); return ST_OK; }
    std::vector<std::thread> pool;
    for (int k = 0; k < nt_use; ++k) pool.emplace_back(worker);
    for (auto& th : pool) th.join();
    return ST_OK;
}

// fn solve_dense_sensitivities (ode_solver/sensitivities.rs:205-262) + dense_write_out_sensitivities (:360-397) for equations
// without output or root functions: a stop time at t_eval[nt - 1], step(), interpolate + interpolate_sens at every point
// passed.  out is n x nt, sens_out is n x np x nt (point-major, then parameter).
static int solve_dense_sens_one(const orc_problem_desc* d, const double* p, int np, const double* t_eval, int nt,
                                double* out, double* sens_out, int64_t* stats) {
    Problem pr;
    int err = build_problem(d, p, np, &pr);
    if (err) return err;
    if (!pr.sens || pr.model.nroots > 0 || pr.model.nout > 0) return ST_BAD_ARG;
    std::unique_ptr<Method> m(make_method(pr, d->method, &err));
    if (!m) { export_stats(pr, nullptr, stats); return err; }
    const int n = pr.n();
    err = m->set_stop_time(t_eval[nt - 1]);
    int col = 0;
    while (!err) {
        StopReason r = m->step(&err);
        if (r == STEP_ERROR) break;
        while (col < nt && t_eval[col] <= m->t() && !err) {
            err = m->interpolate(t_eval[col], out + (size_t)col * n);
            if (!err) err = m->interpolate_sens(t_eval[col], sens_out + (size_t)col * n * np);
            ++col;
        }
        if (r == TSTOP_REACHED) break;
    }
    export_stats(pr, m.get(), stats);
    return err;
}
// instance b: params[b * np ..), out[b] n x nt, sens_out[b] n x np x nt, stats[b] the 16 counters, status[b] the error code
int orc_batch_solve_dense_sens(const orc_problem_desc* d, const double* params, int np, int64_t nbatch, const double* t_eval, int nt,
                               int nthreads, double* out, double* sens_out, int64_t* stats, int32_t* status) {
    Model mm;
    if (!model_by_id(d->model_id, &mm)) return ST_BAD_ARG;
    const int n = mm.n;
    int nt_use = nthreads > 0 ? nthreads : (int)std::thread::hardware_concurrency();
    if (nt_use < 1) nt_use = 1;
    std::atomic<int64_t> next(0);
    auto worker = [&]() {
        while (true) {
            const int64_t b0 = next.fetch_add(16);
            if (b0 >= nbatch) break;
            const int64_t b1 = b0 + 16 < nbatch ? b0 + 16 : nbatch;
            for (int64_t b = b0; b < b1; ++b)
                status[b] = solve_dense_sens_one(d, params + (size_t)b * np, np, t_eval, nt, out + (size_t)b * n * nt,
                                                 sens_out + (size_t)b * n * np * nt, stats + (size_t)b * S_COUNT);
        }
    };
    if (nt_use == 1) { worker(); return ST_OK; }
    std::vector<std::thread> pool;
    for (int k = 0; k < nt_use; ++k) pool.emplace_back(worker);
    for (auto& th : pool) th.join();
    return ST_OK;
}

// The same with `use_tstop = true`: set_stop_time(point) then step until TstopReached; the solution
// is state().y (which for Bdf is the predictor).  out is n x npts.
int orc_harness_tstop(const orc_problem_desc* d, const double* p, int np, const double* t_points, int npts,
                      double* out, int64_t* stats, double* fin) {
    Problem pr;
    int err = build_problem(d, p, np, &pr);
    if (err) return err;
    std::unique_ptr<Method> m(make_method(pr, d->method, &err));
    if (!m) { export_stats(pr, nullptr, stats); return err; }
    const int n = pr.n();
    for (int k = 0; k < npts && !err; ++k) {
        int e = m->set_stop_time(t_points[k]);
        if (e == ST_OK) {
            while (true) {
                StopReason r = m->step(&err);
                if (r == STEP_ERROR || r == TSTOP_REACHED) break;
            }
        }
        for (int i = 0; i < n; ++i) out[(size_t)k * n + i] = m->y()[i];
    }
    export_stats(pr, m.get(), stats);
    if (fin) { fin[0] = m->t(); fin[1] = m->h(); fin[2] = (double)m->cur_order(); }
    return err;
}

// nonzeros2graph + color_graph_greedy on a given pattern (the reference's build_coloring test, jacobian/mod.rs:483-513)
int orc_greedy_coloring(const int32_t* rows, const int32_t* cols, int nnz, int n, int32_t* colors) {
    std::vector<std::pair<int, int>> nz;
    for (int k = 0; k < nnz; ++k) nz.push_back({rows[k], cols[k]});
    const std::vector<int> r = greedy_coloring(nz, n);
    for (int j = 0; j < n; ++j) colors[j] = r[j];
    return ST_OK;
}

// The reference's residual-operator unit tests (op/bdf.rs:317-360, op/sdirk.rs:338-388): F(x) and the iteration matrix of
// the method's callable with its scalars and vector set directly.  F is n, A is n x n column-major.
int orc_residual_known_answer(const orc_problem_desc* d, const double* p, int np, double c, double h, const double* vec,
                              const double* x, double t, double* F, double* A) {
    Problem pr;
    int err = build_problem(d, p, np, &pr);
    if (err) return err;
    std::unique_ptr<Method> m(make_method(pr, d->method, &err));
    if (!m) return err;
    return m->residual_known_answer(c, h, vec, x, t, F, A);
}

// The driver loop of the reference's test_ball_bounce (ode_solver/mod.rs:1024-1080): set_stop_time(tstop), step until the
// first root, put the solver's state at the root and apply the model's reset function (the test's hand-written update
// v <- -e v, x <- max(x, eps), dy[0] <- v), then take up to nsteps further steps and record (t, y) after each; a step
// that reaches the stop time is the last one.  rows is nsteps x (1 + n); *ntaken = rows written.
int orc_steps_after_first_root(const orc_problem_desc* d, const double* p, int np, double tstop, int nsteps,
                               double* rows, int* ntaken) {
    *ntaken = 0;
    Problem pr;
    int err = build_problem(d, p, np, &pr);
    if (err) return err;
    std::unique_ptr<Method> m(make_method(pr, d->method, &err));
    if (!m) return err;
    const int n = pr.n();
    err = m->set_stop_time(tstop);
    if (err) return err;
    while (true) {
        StopReason r = m->step(&err);
        if (r == STEP_ERROR) return err;
        if (r == TSTOP_REACHED) break;
        if (r == ROOT_FOUND) {
            err = m->state_mut_back(m->root_t());
            if (!err) err = m->apply_reset();
            if (err) return err;
            break;
        }
    }
    for (int k = 0; k < nsteps; ++k) {
        StopReason r = m->step(&err);
        if (r == STEP_ERROR) return err;
        if (r == ROOT_FOUND) return ST_BAD_ARG;            // "should be an internal timestep but found a root"
        rows[(size_t)k * (1 + n)] = m->t();
        for (int i = 0; i < n; ++i) rows[(size_t)k * (1 + n) + 1 + i] = m->y()[i];
        *ntaken = k + 1;
        if (r == TSTOP_REACHED) break;
    }
    return ST_OK;
}

// Batched driver: instance b uses params[b*np .. (b+1)*np) (instance-major, as the reference lays
// out batched parameters, test_models/exponential_decay.rs:297-304).  out[b] is n x nt col-major,
// stats[b] the 16 counters, status[b] the error code.  threads over instances = the CPU baseline.
int orc_batch_solve_dense_roots(const orc_problem_desc* d, const double* params, int np, int64_t nbatch,
                                const double* t_eval, int nt, int nthreads,
                                double* out, int64_t* stats, int32_t* status, double* roots /* [nbatch][3] or NULL */);
int orc_batch_solve_dense(const orc_problem_desc* d, const double* params, int np, int64_t nbatch,
                          const double* t_eval, int nt, int nthreads,
                          double* out, int64_t* stats, int32_t* status) {
    return orc_batch_solve_dense_roots(d, params, np, nbatch, t_eval, nt, nthreads, out, stats, status, nullptr);
}
int orc_batch_solve_dense_roots(const orc_problem_desc* d, const double* params, int np, int64_t nbatch,
                                const double* t_eval, int nt, int nthreads,
                                double* out, int64_t* stats, int32_t* status, double* roots) {
    Model mm;
    if (!model_by_id(d->model_id, &mm)) return ST_BAD_ARG;
    const int n = mm.ncols_out();                           // rows of every instance's result block
    // std::thread pool with a shared work counter (dynamic schedule, 16 instances per grab);
    // libgomp is not usable in this image.
    int nt_use = nthreads > 0 ? nthreads : (int)std::thread::hardware_concurrency();
    if (nt_use < 1) nt_use = 1;
    std::atomic<int64_t> next(0);
    auto worker = [&]() {
        const int64_t chunk = 16;
        while (true) {
            int64_t b0 = next.fetch_add(chunk);
            if (b0 >= nbatch) break;
            int64_t b1 = b0 + chunk < nbatch ? b0 + chunk : nbatch;
            for (int64_t b = b0; b < b1; ++b) {
                int err = solve_dense_one(d, params + (size_t)b * np, np, t_eval, nt,
                                          out ? out + (size_t)b * n * nt : nullptr,
                                          stats ? stats + (size_t)b * S_COUNT : nullptr, nullptr,
                                          roots ? roots + (size_t)b * 3 : nullptr);
                if (status) status[b] = err;
            }
        }
    };
    if (nt_use == 1) worker();
    else {
        std::vector<std::thread> pool;
        for (int i = 0; i < nt_use; ++i) pool.emplace_back(worker);
        for (auto& th : pool) th.join();
    }
    return ST_OK;
}

// dlopen a model plugin (built by oracle/oracle.py: load_user_model from the user's source) -> model id, or -1
int orc_load_model_plugin(const char* path) {
    void* h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!h) return -1;
    typedef void (*fn_t)(orc::Model*);
    fn_t fn = (fn_t)dlsym(h, "orc_plugin_model");
    if (!fn) return -1;
    orc::Model m;
    fn(&m);
    return orc::register_plugin_model(m);
}

// experiment switch (tools/dmma_rounding_check.py): fused multiply-adds in the LU trailing updates; returns the old value
int orc_set_fused_lu_updates(int on) { const int old = orc::g_fused_lu_updates; orc::g_fused_lu_updates = on; return old; }

int orc_num_threads() { return (int)std::thread::hardware_concurrency(); }

double orc_pow(double x, double y, int powmode) { Math m; m.powmode = powmode; return m.pow(x, y); }
double orc_powi(double x, int n) { return dsb_powi(x, n); }

// Known-answer hooks for the reference's unit tests of the building blocks.
double orc_squared_norm(const double* x, const double* y, const double* atol, double rtol, int n) {
    return squared_norm(x, y, atol, rtol, n);
}
int orc_lu_solve(const double* A, int n, double* b) {
    DenseLU lu; lu.factor(A, n);
    return lu.solve(b) ? 0 : ST_LU_SOLVE_FAILED;
}
int orc_lu_factor(const double* A, int n, double* lu_out, int32_t* piv_out) {
    DenseLU lu; lu.factor(A, n);
    for (size_t i = 0; i < (size_t)n * n; ++i) lu_out[i] = lu.lu[i];
    // expand the permutation sequence into LAPACK-style ipiv (row i swapped with piv[i])
    for (int i = 0; i < n; ++i) piv_out[i] = i;
    for (auto& pq : lu.perm) piv_out[pq.first] = pq.second;
    return 0;
}

}  // extern "C"
