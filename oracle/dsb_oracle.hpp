// oracle/dsb_oracle.hpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A scalar f64 CPU restatement of the one hot path of martinjrobins/diffsol that this repository
// accelerates: the implicit step loop of `Bdf` and `Sdirk` and everything they call.  Only
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may build,
// load or call it.  The product (diffsol_b200/) never links or imports anything from oracle/.
//
// The reference is Rust and cannot be compiled in this image (no rustc/cargo, no vendored crates),
// so this is a restatement, written from the reference sources cited on every function
// (paths relative to /root/reference/crates).  It is PINNED against the reference's own inline
// `insta` statistics snapshots and golden solution tables (tests/test_oracle_golden.py; forward sensitivities:
// tests/test_sensitivities.py, 7 of the reference's 8 nalgebra snapshots; `solve(final_time)`: tests/test_solve_ragged.py):
// every integer of OdeSolverStatistics / OpStatistics must match.
//
// Third-party arithmetic the reference delegates to and that is restated here from the published
// algorithms: nalgebra 0.35 `DMatrix::lu()` / `LU::solve_mut` (partial pivoting, first max,
// reciprocal-pivot scaling, column-axpy substitution) and `gemm`/`gemv`/`axpy` evaluation order.
//
// pow mode: 0 = libm `pow` (what Rust's f64::powf calls on Linux -- the reference-literal mode),
//           1 = dsb_pow (the deterministic pow shared with the CUDA kernels, csrc/dsb_math.h).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <limits>
#include <utility>
#include <vector>

#include "../diffsol_b200/csrc/dsb_math.h"
#include "../diffsol_b200/csrc/dsb_models.h"

namespace orc {

typedef std::vector<double> Vec;

// ---- error codes (mirror diffsol's OdeSolverError / NonLinearSolverError variants) ------------
enum Status : int {
    ST_OK = 0,
    ST_STEP_SIZE_TOO_SMALL = 1,             // OdeSolverError::StepSizeTooSmall
    ST_TOO_MANY_ERROR_TEST_FAILURES = 2,    // OdeSolverError::TooManyErrorTestFailures
    ST_TOO_MANY_NONLINEAR_FAILURES = 3,     // OdeSolverError::TooManyNonlinearSolverFailures
    ST_STOP_TIME_BEFORE_CURRENT = 4,        // OdeSolverError::StopTimeBeforeCurrentTime
    ST_STOP_TIME_AT_CURRENT = 5,            // OdeSolverError::StopTimeAtCurrentTime
    ST_INITIAL_CONDITION_DID_NOT_CONVERGE = 6,
    ST_LINESEARCH_FAILED = 7,
    ST_INTERPOLATION_TIME_AFTER_CURRENT = 8,
    ST_LU_SOLVE_FAILED = 9,
    ST_BAD_ARG = 10,
};

// ---- statistics layout shared with the C ABI (include/diffsol_b200.h) -------------------------
enum StatIdx {
    S_SETUPS = 0, S_SETUPS_CHECKPOINT, S_SETUPS_FIRST_FAIL, S_SETUPS_SECOND_FAIL, S_SETUPS_ERROR_FAIL,
    S_SETUPS_SUCCESS, S_STEPS, S_ERROR_TEST_FAILS, S_NL_ITERS, S_NL_FAILS,
    S_RHS_CALLS, S_RHS_JAC_MULS, S_RHS_MATRIX_EVALS, S_RESERVED0, S_RESERVED1, S_RESERVED2,
    S_COUNT = 16
};

struct Math {
    int powmode = 0;
    double pow(double x, double y) const { return powmode == 0 ? ::pow(x, y) : dsb_pow(x, y); }
    // f64::powi -> __powidf2 in both modes (it is plain multiplication)
    double powi(double x, int n) const { return dsb_powi(x, n); }
};

// ---- the user equations as function pointers (Closure / LinearClosure) -------------------------
struct Model {
    int n = 0, np = 0;
    bool has_mass = false;
    void (*rhs)(const double*, const double*, double, double*) = nullptr;
    void (*jac_mul)(const double*, const double*, double, const double*, double*) = nullptr;
    void (*mass)(const double*, const double*, double, double, double*) = nullptr;
    void (*init)(const double*, double, double*) = nullptr;
    int nroots = 0;                                                       // OdeEquations::root (ode_equations/mod.rs)
    void (*root)(const double*, const double*, double, double*) = nullptr;
    int nout = 0;                                                         // OdeEquations::out: 0 = none (solve_dense returns the states)
    void (*out)(const double*, const double*, double, double*) = nullptr;
    int ncols_out() const { return nout > 0 ? nout : n; }                 // rows of the solve_dense result
    void (*reset)(const double*, const double*, double, double*) = nullptr;   // OdeEquations::reset: y <- reset(y, t) at a root
    // OdeEquationsImplicitSens: f_p(x, p, t) v and (d y0 / d p) v (op/closure_with_sens.rs:181-183, op/constant_closure_with_sens.rs)
    void (*sens_mul)(const double*, const double*, double, const double*, double*) = nullptr;
    void (*init_sens)(const double*, double, const double*, double*) = nullptr;
};
template <class M, int NR = dsb_model_nroots<M>::value> struct RootOf {
    static void set(Model& m) {          // M::root may be a template over the state accessor (component-wise models)
        m.nroots = NR;
        m.root = [](const double* x, const double* p, double t, double* g) { M::root(x, p, t, g); };
    }
};
template <class M> struct RootOf<M, 0> { static void set(Model&) {} };
template <class M, bool HAS = dsb_model_nout<M>::has_out> struct OutOf {
    static void set(Model& m) {
        m.nout = dsb_model_nout<M>::value;
        m.out = [](const double* x, const double* p, double t, double* o) { M::out(x, p, t, o); };
    }
};
template <class M> struct OutOf<M, false> { static void set(Model&) {} };
template <class M, bool HAS = dsb_model_has_reset<M>::value> struct ResetOf {
    static void set(Model& m) { m.reset = [](const double* x, const double* p, double t, double* y) { M::reset(x, p, t, y); }; }
};
template <class M> struct ResetOf<M, false> { static void set(Model&) {} };
template <class M, bool HAS = dsb_model_has_sens<M>::value> struct SensOf {
    static void set(Model& m) { m.sens_mul = &M::sens_mul; m.init_sens = &M::init_sens; }
};
template <class M> struct SensOf<M, false> { static void set(Model&) {} };
template <class M>
Model make_model() {
    Model m;
    m.n = M::N; m.np = M::NP; m.has_mass = M::HAS_MASS;
    m.rhs = &M::rhs; m.jac_mul = &M::jac_mul; m.mass = &M::mass; m.init = &M::init;
    RootOf<M>::set(m);
    OutOf<M>::set(m);
    ResetOf<M>::set(m);
    SensOf<M>::set(m);
    return m;
}
bool model_by_id(int id, Model* out);
// Equation sets that are not compiled into the oracle: a shared object built from the USER'S source (the same text the
// CUDA library compiles for the device, oracle/oracle.py: load_user_model) exports `void orc_plugin_model(orc::Model*)`;
// its Model is registered under an id >= ORC_PLUGIN_ID0.
constexpr int ORC_PLUGIN_ID0 = 1000;
int register_plugin_model(const Model& m);
std::vector<int> greedy_coloring(const std::vector<std::pair<int, int>>& non_zeros, int n);   // 1-based colour per column

// ---- OdeSolverOptions / InitialConditionSolverOptions (ode_solver/problem.rs:15-152) -----------
struct Options {
    int max_nonlinear_solver_iterations = 10;
    int max_error_test_failures = 40;
    int max_nonlinear_solver_failures = 50;
    double nonlinear_solver_tolerance = 0.2;
    double min_timestep = 1e-13;
    double max_timestep_growth = 2.0;   // BdfConfig/SdirkConfig defaults (config.rs:54-73)
    double min_timestep_growth = 2.0;
    double max_timestep_shrink = 0.9;
    double min_timestep_shrink = 0.5;
    int update_jacobian_after_steps = 20;
    int update_rhs_jacobian_after_steps = 50;
    double threshold_to_update_jacobian = 0.3;
    double threshold_to_update_rhs_jacobian = 0.2;
    double pi_control_proportional = 0.0;
    double pi_control_integral = 0.5;
    // ic_options
    bool ic_use_linesearch = true;
    int ic_max_linesearch_iterations = 10;
    int ic_max_newton_iterations = 10;
    int ic_max_linear_solver_setups = 4;
    double ic_step_reduction_factor = 0.5;
    double ic_armijo_constant = 1e-4;
};

// ---- OdeSolverProblem (what OdeBuilder::build returns) ------------------------------------------
struct Problem {
    Model model;
    Vec p;
    double rtol = 1e-6;        // builder.rs:112-140 defaults
    Vec atol;                  // length n
    double t0 = 0.0, h0 = 1.0;
    bool use_coloring = false;
    // forward sensitivities (problem.bdf_sens(), ode_solver/problem.rs:819-830): one sensitivity vector per parameter;
    // sens_rtol / sens_atol (builder.rs:1682-1716; per state, param_scales = 1) put them into the error test
    bool sens = false;
    bool sens_error_control = false;
    double sens_rtol = 0.0;
    Vec sens_atol;             // length n
    Options opt;
    Math math;
    // colouring data (jacobian/mod.rs:178-214), built by build_coloring()
    std::vector<std::pair<int, int>> non_zeros;              // (row, col)
    std::vector<std::vector<int>> color_inputs;             // per colour: columns seeded with 1
    std::vector<std::vector<std::pair<int, int>>> color_entries;  // per colour: (row, col) written
    // OpStatistics of the rhs closure (op/mod.rs:108-145)
    mutable int64_t n_calls = 0, n_jac_muls = 0, n_matrix_evals = 0;

    int n() const { return model.n; }
    void rhs(const double* x, double t, double* y) const { ++n_calls; model.rhs(x, p.data(), t, y); }
    void jac_mul(const double* x, double t, const double* v, double* y) const {
        ++n_jac_muls; model.jac_mul(x, p.data(), t, v, y);
    }
    // Closure::jacobian_inplace (op/closure.rs:140-147): coloured or default column-by-column
    void jacobian(const double* x, double t, double* J /* n*n col-major */) const;
    // LinearOp::_default_matrix_inplace (op/linear_op.rs:42-51); identity when there is no mass
    void mass_matrix(double t, double* M /* n*n col-major */) const;
    void mass_gemv(const double* x, double t, double beta, double* y) const { model.mass(x, p.data(), t, beta, y); }
    void build_coloring();     // builder.rs:1852-1857 -> calculate_sparsity
};

// ---- Vector::squared_norm (diffsol-la/src/vector/nalgebra_serial.rs:395-408) --------------------
double squared_norm(const double* x, const double* y, const double* atol, double rtol, int n);

// ---- nalgebra 0.35 LU (restated; call sites diffsol-la/src/linear_solver/nalgebra/lu.rs:36,50) --
extern int g_fused_lu_updates;            // experiment switch, see dsb_oracle_core.cpp
struct DenseLU {
    int n = 0;
    Vec lu;                                  // col-major
    std::vector<std::pair<int, int>> perm;   // PermutationSequence
    void factor(const double* A, int n_);
    bool solve(double* b) const;             // false <=> zero on U's diagonal (LuSolveFailed)
};

// ---- Convergence (diffsol-nl/src/convergence.rs) ------------------------------------------------
enum ConvStatus { CONVERGED, DIVERGED, CONTINUE };
struct Convergence {
    double rtol = 0; const double* atol = nullptr; int n = 0;
    double tol = 0.2; int max_iter = 10; int niter = 0;
    bool has_old_norm = false; double old_norm = 0; double eta = 0;
    const Math* math = nullptr;
    void init(double rtol_, const double* atol_, int n_, double tol_, const Math* m);
    void reset_eta() { eta = math->pow(20.0, 1.25); }
    void reset_eta_timestep_change() { eta = math->pow(100.0, 1.25); }
    void reset() { niter = 0; has_old_norm = false; }
    double norm(const double* dy, const double* y) const { return std::sqrt(squared_norm(dy, y, atol, rtol, n)); }
    ConvStatus check_norm(double norm);
    ConvStatus check_new_iteration(double norm);
};

// ---- JacobianUpdate (ode_solver/jacobian_update.rs) ---------------------------------------------
enum SolverState { STEP_SUCCESS, FIRST_CONVERGENCE_FAIL, SECOND_CONVERGENCE_FAIL, ERROR_TEST_FAIL, CHECKPOINT };
struct JacobianUpdate {
    int steps_since_jacobian_eval = 0, steps_since_rhs_jacobian_eval = 0;
    double h_at_last_jacobian_update = 1.0;
    double threshold_to_update_jacobian = 0.3, threshold_to_update_rhs_jacobian = 0.2;
    int update_jacobian_after_steps = 20, update_rhs_jacobian_after_steps = 50;
    void init(const Options& o, double h_at_last);
    void update_jacobian(double h) { steps_since_jacobian_eval = 0; h_at_last_jacobian_update = h; }
    void update_rhs_jacobian(double h) {
        steps_since_rhs_jacobian_eval = 0; steps_since_jacobian_eval = 0; h_at_last_jacobian_update = h;
    }
    void step() { ++steps_since_jacobian_eval; ++steps_since_rhs_jacobian_eval; }
    bool check_jacobian_update(double h, SolverState s) const;
    bool check_rhs_jacobian_update(double h, SolverState s) const;
};

struct Stats {
    int64_t v[S_COUNT];
    Stats() { std::memset(v, 0, sizeof(v)); }
    void record_linear_solver_setup(SolverState s);   // ode_solver/mod.rs:53-68
};

// Result of consistent initialisation + initial step size: what `problem.bdf_state()` /
// `problem.rk_state()` produce (ode_solver/state.rs:969-997).
struct InitialState {
    Vec y, dy; double t = 0, h = 0;
};
int new_and_consistent(const Problem& pr, int solver_order, InitialState* st);
// InitOp + Newton with the backtracking line search for the algebraic rows (state.rs:84-162 for the state, :167-238 for a
// sensitivity vector); a no-op without a singular mass matrix
int consistent_solve(const Problem& pr, const std::function<void(const double*, double, double*)>& eq_rhs,
                     const std::function<void(const double*, double, double*)>& eq_jacobian, Vec& y_state, Vec& dy_state,
                     Convergence* shared_conv, bool zero_dv, int use_linesearch = -1 /* -1: ic_options.use_linesearch */);

enum StopReason { INTERNAL_TIMESTEP = 0, TSTOP_REACHED = 1, STEP_ERROR = 2, ROOT_FOUND = 3 };

// ---- RootFinder (diffsol/src/nonlinear_solver/root.rs:12-160) + Vector::root_finding
// (diffsol-la/src/vector/nalgebra_serial.rs:484-504) ------------------------------------------------
struct RootFinder {
    double t0 = 0.0;
    Vec g0, g1, gmid, ymid;
    void resize(int nroots, int nstates) { g0.assign(nroots, 0.0); g1 = g0; gmid = g0; ymid.assign(nstates, 0.0); }
    // (found_root, max_frac, max_frac_index): g0 = lower end, g1 = upper end
    static void root_finding(const Vec& g0, const Vec& g1, bool* found_root, int* imax);
    // root.rs:34-37
    void init(const Problem& pr, const double* y, double t);
    // root.rs:60-160: true <=> a root was found in (t0, t]; *t_root, *idx
    bool check_root(const Problem& pr, const std::function<int(double, double*)>& interpolate, const double* y,
                    double t, double* t_root, int* idx);
};

// ---- the abstract surface both integrators share (OdeSolverMethod, ode_solver/method.rs:42-618) --
struct Method {
    virtual ~Method() {}
    virtual StopReason step(int* err) = 0;
    virtual int set_stop_time(double tstop) = 0;
    virtual int interpolate(double t, double* y) const = 0;
    virtual double t() const = 0;
    virtual double h() const = 0;
    virtual int cur_order() const = 0;
    virtual const double* y() const = 0;
    virtual const Stats& stats() const = 0;
    // OdeSolverMethod::interpolate_sens (bdf.rs:1162-1190): one vector per parameter, out is n x np (parameter-major)
    virtual int interpolate_sens(double, double*) const { return ST_BAD_ARG; }
    // RootFound(t, index) of the last step() that returned ROOT_FOUND
    virtual double root_t() const { return 0.0; }
    virtual int root_index() const { return -1; }
    // OdeSolverMethod::state_mut_back (bdf.rs:1228-1262): move the state back to t inside the last step
    virtual int state_mut_back(double) { return ST_BAD_ARG; }
    // OdeSolverMethod::apply_reset (method.rs:175-181 -> state.rs:246-270): y <- reset(y, t), dy <- f(y, t)
    virtual int apply_reset() { return ST_BAD_ARG; }
    // Test hook for the reference's residual-operator tests (op/bdf.rs:317-360 test_bdf_callable, op/sdirk.rs:338-388
    // test_sdirk_callable): set the callable's scalars (Bdf: c; Sdirk: c and h) and its vector (Bdf: psi - y0; Sdirk:
    // phi) directly, then evaluate F(x) and the iteration matrix A (n x n col-major) with the SAME member functions
    // step() uses
    virtual int residual_known_answer(double /*c*/, double /*h*/, const double* /*vec*/, const double* /*x*/, double /*t*/,
                                      double* /*F*/, double* /*A*/) { return ST_BAD_ARG; }
};

Method* new_bdf(const Problem& pr, int* err);
Method* new_sdirk(const Problem& pr, int tableau /*0 = tr_bdf2, 1 = esdirk34*/, int* err);

// fn solve_dense (ode_solver/method.rs:721-818) + OdeSolverMethod::solve_dense (:467-505), without checkpointing.
// With a reset function (OdeEquations::reset) a root does not end the solve: the state is moved back to the root, reset,
// the stop time is set again and the integration continues (method.rs:774-805).  When a root stops the integration, the columns up to the root are written, the state at the root
// goes into the next column (when there is one) and *ncols / *root_t / *root_idx say so (ncols = nt otherwise).
// With an output function (OdeEquations::out, dense_write_out method.rs:822-848) every column holds out(y(t), t)
// (nout values) instead of the n states; `pr` supplies it (may be NULL: states).
int solve_dense(Method& s, const double* t_eval, int nt, int n, double* out /* n*nt (or nout*nt) col-major */,
                int* ncols = nullptr, double* root_t = nullptr, int* root_idx = nullptr, const Problem* pr = nullptr);
int solve_ragged(Method& s, double final_time, int n, const Problem* pr, int max_cols, double* ts, double* ys, int* ncols,
                 double* root_t, int* root_idx);

}  // namespace orc
