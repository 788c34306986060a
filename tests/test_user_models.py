"""User equation sets crossing the C ABI as SOURCE TEXT ("user RHS closures and DiffSL-JIT modules drop in"):
the DiffSL symbol table of the reference's external-module test
(/root/reference/crates/diffsol-c/tests/external-dynamic-logistic/src/lib.rs:15-600) restated as C, and a closure-style
functor.  nvcc compiles the text at run time into one instantiation of the library's kernel families
(dsb_model_library_build / dsb_model_library_load); the oracle compiles the SAME text for the host (oracle.load_user_model).

CPU tests: the oracle on the user model against the analytic solution (the reference's own check, ASSERT_TOL = 1e-5,
crates/diffsol-c/tests/common/mod.rs:8), and that the model plugin cross-compiles and registers (no compute).
GPU tests: the CUDA path bit-identical to the oracle on the user model, the analytic solution, the stop at x = 0.5."""
import ctypes
import shutil

import numpy as np
import pytest

# logistic growth x' = r x (1 - x), x(0) = 0.1, input r; out = x; stop = x - 0.5: the symbol table of
# external-dynamic-logistic/src/lib.rs (set_u0 :15, rhs :123, rhs_grad :142, mass :233, calc_out :303, calc_stop :407,
# get_dims :529, set_inputs :576), every function restated with the reference's body
LOGISTIC_DIFFSL = r"""
#define DSB_DIFFSL_STATES 1
#define DSB_DIFFSL_INPUTS 1
#define DSB_DIFFSL_OUTPUTS 1
#define DSB_DIFFSL_DATA 1
#define DSB_DIFFSL_STOP %(stop)d
#define DSB_DIFFSL_HAS_MASS %(has_mass)d
DSB_SYMBOL void set_u0(double* u, double* data, unsigned thread_id, unsigned thread_dim) { if (u) *u = 0.1; }
DSB_SYMBOL void rhs(double t, const double* u, double* data, double* rr, unsigned thread_id, unsigned thread_dim) {
    if (!u || !data || !rr) return;
    const double x = *u, r = *data;
    *rr = r * x * (1.0 - x);
}
DSB_SYMBOL void rhs_grad(double t, const double* u, const double* du, const double* data, double* ddata, const double* rr,
                         double* drr, unsigned thread_id, unsigned thread_dim) {
    if (!u || !du || !data || !ddata || !drr) return;
    const double x = *u, dx = *du, r = *data;
    *drr = r * (1.0 - 2.0 * x) * dx;
    *ddata = x * (1.0 - x);
}
DSB_SYMBOL void mass(double t, const double* v, double* data, double* mv, unsigned thread_id, unsigned thread_dim) { if (v && mv) *mv = *v; }
DSB_SYMBOL void calc_out(double t, const double* u, double* data, double* out, unsigned thread_id, unsigned thread_dim) { if (u && out) *out = *u; }
DSB_SYMBOL void calc_stop(double t, const double* u, double* data, double* root, unsigned thread_id, unsigned thread_dim) { if (u && root) *root = *u - 0.5; }
DSB_SYMBOL void set_inputs(const double* inputs, double* data, unsigned model_index) { if (inputs && data) *data = *inputs; }
DSB_SYMBOL void get_dims(unsigned* states, unsigned* inputs, unsigned* outputs, unsigned* data, unsigned* stop, unsigned* has_mass,
                         unsigned* has_reset) {
    if (states) *states = 1; if (inputs) *inputs = 1; if (outputs) *outputs = 1; if (data) *data = 1; if (stop) *stop = %(stop)d;
    if (has_mass) *has_mass = %(has_mass)d; if (has_reset) *has_reset = 0;
}
"""

# a closure-style functor (builder.rs:192-200 signatures): the Brusselator, n = 2, p = [a, b]
BRUSSELATOR = r"""
struct Brusselator {
    static constexpr int N = 2, NP = 2;
    static constexpr bool HAS_MASS = false;
    DSB_HD static void rhs(const double* x, const double* p, double, double* y) {
        y[0] = p[0] + x[0] * x[0] * x[1] - (p[1] + 1.0) * x[0];
        y[1] = p[1] * x[0] - x[0] * x[0] * x[1];
    }
    DSB_HD static void jac_mul(const double* x, const double* p, double, const double* v, double* y) {
        y[0] = (2.0 * x[0] * x[1] - (p[1] + 1.0)) * v[0] + x[0] * x[0] * v[1];
        y[1] = (p[1] - 2.0 * x[0] * x[1]) * v[0] - x[0] * x[0] * v[1];
    }
    DSB_HD static void mass(const double* x, const double*, double, double beta, double* y) { y[0] = x[0] + beta * y[0]; y[1] = x[1] + beta * y[1]; }
    DSB_HD static void init(const double* p, double, double* y) { y[0] = 1.0 + 0.1 * p[0]; y[1] = 1.0; }
};
"""


def logistic_exact(r, t, x0=0.1):
    return 1.0 / (1.0 + (1.0 / x0 - 1.0) * np.exp(-np.asarray(r)[:, None] * np.asarray(t)[None, :]))


def logistic_params(B):
    from diffsol_b200 import sweeps
    return (0.5 + 2.0 * sweeps.uniform(np.arange(B), 0)).reshape(-1, 1)


@pytest.mark.parametrize("method", ["bdf", "tr_bdf2", "esdirk34"])
def test_oracle_on_the_logistic_symbol_table_matches_the_analytic_solution(oracle, method):
    """The reference's check of its external logistic module (solve_dense against the closed form, 1e-5), on the oracle."""
    name = oracle.load_user_model(LOGISTIC_DIFFSL % dict(stop=0, has_mass=0), kind="diffsl")
    assert oracle.model_dims(name) == (1, 1, False) and oracle.model_nout(name) == 1
    r = logistic_params(16)
    t_eval = np.linspace(0.25, 4.0, 16)
    ys, stats, status = oracle.batch_solve_dense(oracle.make_desc(name, method=method, rtol=1e-8, atol=1e-10), r, t_eval)
    assert (status == 0).all()
    assert np.abs(ys[:, :, 0] - logistic_exact(r[:, 0], t_eval)).max() < 1e-5


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="nvcc cross-compiles the model plugin")
def test_model_plugin_cross_compiles_and_registers():
    """dsb_model_library_build + dsb_model_library_load without a GPU: the plugin builds for sm_100a, loads, and
    dsb_problem_new reports the dimensions the symbol table declares."""
    import diffsol_b200
    from diffsol_b200 import capi
    name = capi.load_model_source(LOGISTIC_DIFFSL % dict(stop=1, has_mass=0), kind="diffsl")
    assert capi.MODELS[name] >= 1000
    prob = diffsol_b200.OdeBuilder().rhs_implicit(name).p(logistic_params(3)).build()
    assert (prob.nstates, prob.nparams, prob.nout) == (1, 1, 1)
    # a second request for the same text is served from the registry
    assert capi.load_model_source(LOGISTIC_DIFFSL % dict(stop=1, has_mass=0), kind="diffsl") == name
    # a source that does not compile reports nvcc's message through dsb_last_error
    with pytest.raises(capi.DiffsolB200Error, match="nvcc failed"):
        capi.load_model_source("this is not C", kind="functor", struct="Nope")


@pytest.fixture(scope="module")
def dsb():
    import diffsol_b200
    from diffsol_b200 import capi
    capi.require_device()
    return diffsol_b200


@pytest.mark.gpu
@pytest.mark.parametrize("method", ["bdf", "tr_bdf2", "esdirk34"])
@pytest.mark.parametrize("has_mass", [0, 1])
def test_logistic_symbol_table_on_the_gpu(dsb, oracle, method, has_mass):
    """The logistic module as a string -> nvcc -> the lane kernels: counters, status and outputs bit-identical to the oracle
    (which compiled the same string for the host), the analytic solution within the reference's 1e-5."""
    src = LOGISTIC_DIFFSL % dict(stop=0, has_mass=has_mass)
    B = 2000
    r = logistic_params(B)
    t_eval = np.linspace(0.25, 4.0, 16)
    prob = dsb.OdeBuilder().rhs_implicit_source(src, kind="diffsl").p(r).rtol(1e-8).atol(1e-10).build()
    solver = getattr(prob, method)()
    ys = solver.solve_dense(t_eval)
    name = oracle.load_user_model(src, kind="diffsl")
    ys_o, stats_o, status_o = oracle.batch_solve_dense(oracle.make_desc(name, method=method, powmode=1, rtol=1e-8, atol=1e-10), r, t_eval)
    assert np.array_equal(solver.status(), status_o) and (status_o == 0).all()
    assert np.array_equal(solver.statistics_array()[:, :13], stats_o[:, :13])
    assert np.array_equal(ys, ys_o)
    assert np.abs(ys[:, :, 0] - logistic_exact(r[:, 0], t_eval)).max() < 1e-5


@pytest.mark.gpu
def test_logistic_stop_function_on_the_gpu(dsb, oracle):
    """calc_stop = x - 0.5: every instance stops at t = ln 9 / r (solve_dense's RootFound branch), bit-identical to the oracle."""
    src = LOGISTIC_DIFFSL % dict(stop=1, has_mass=0)
    B = 1000
    r = logistic_params(B)
    t_eval = np.linspace(0.25, 8.0, 32)
    solver = dsb.OdeBuilder().rhs_implicit_source(src, kind="diffsl").p(r).rtol(1e-8).atol(1e-10).build().bdf()
    ys = solver.solve_dense(t_eval)
    root_idx, ncols = solver.root_info()
    name = oracle.load_user_model(src, kind="diffsl")
    o = oracle.batch_solve_dense_roots(oracle.make_desc(name, powmode=1, rtol=1e-8, atol=1e-10), r, t_eval)
    ys_o, stats_o, status_o, t_root_o, root_idx_o, ncols_o = o
    assert np.array_equal(solver.status(), status_o) and np.array_equal(root_idx, root_idx_o) and np.array_equal(ncols, ncols_o)
    assert np.array_equal(solver.statistics_array()[:, :13], stats_o[:, :13])
    assert np.array_equal(ys, ys_o, equal_nan=True)
    stopped = root_idx >= 0
    assert stopped.sum() > B // 2
    t_fin = solver.final_state()[0]
    assert np.abs(t_fin[stopped] - np.log(9.0) / r[stopped, 0]).max() < 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("method", ["bdf", "tr_bdf2"])
def test_closure_style_functor_on_the_gpu(dsb, oracle, method):
    from diffsol_b200 import sweeps
    B = 3000
    i = np.arange(B)
    p = np.stack([0.5 + sweeps.uniform(i, 0), 1.5 + 2.0 * sweeps.uniform(i, 1)], axis=1)
    t_eval = np.linspace(1.0, 20.0, 20)
    prob = dsb.OdeBuilder().rhs_implicit_source(BRUSSELATOR, kind="functor", struct="Brusselator").p(p).rtol(1e-6).atol(1e-8).build()
    solver = getattr(prob, method)()
    ys = solver.solve_dense(t_eval)
    name = oracle.load_user_model(BRUSSELATOR, kind="functor", struct="Brusselator")
    ys_o, stats_o, status_o = oracle.batch_solve_dense(oracle.make_desc(name, method=method, powmode=1, rtol=1e-6, atol=1e-8), p, t_eval)
    assert np.array_equal(solver.status(), status_o) and (status_o == 0).all()
    assert np.array_equal(solver.statistics_array()[:, :13], stats_o[:, :13])
    assert np.array_equal(ys, ys_o)
