"""ctypes wrapper around oracle/_ref/libdsb_oracle.so -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module.  The product package (diffsol_b200) never does.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "libdsb_oracle.so")

S_NAMES = [
    "number_of_linear_solver_setups",
    "number_of_linear_solver_setups_from_checkpoint",
    "number_of_linear_solver_setups_from_first_convergence_fail",
    "number_of_linear_solver_setups_from_second_convergence_fail",
    "number_of_linear_solver_setups_from_error_test_fail",
    "number_of_linear_solver_setups_from_step_success",
    "number_of_steps",
    "number_of_error_test_failures",
    "number_of_nonlinear_solver_iterations",
    "number_of_nonlinear_solver_fails",
    "rhs_number_of_calls",
    "rhs_number_of_jac_muls",
    "rhs_number_of_matrix_evals",
]
S_COUNT = 16

METHODS = {"bdf": 0, "tr_bdf2": 1, "esdirk34": 2}
MODELS = {
    "exp_decay": 0,
    "exp_decay_algebraic": 1,
    "robertson_dae": 2,
    "robertson_ode": 3,
    "robertson_ode_g3": 4,
    "dydt_y2": 5,
    "gaussian_decay": 6,
    "van_der_pol": 7,
    "van_der_pol_scaled": 8,
    "heat1d_dae_256": 9,
    "heat1d_dae_32": 10,
    "spm": 11,
    "spm99": 12,
    "exp_decay_root": 13,
    "spm_stop": 14,
    "spm99_stop": 15,
    "heat1d_dae_32_bc": 16,
    "exp_decay_reset": 17,
    "heat2d_10": 18,
    "ball_bounce": 19,
    "exp_decay_two_roots": 20,
    "spm_cycle": 21,
    "exp_decay_algebraic_reset": 22,
}


class ProblemDesc(ctypes.Structure):
    _fields_ = [
        ("model_id", ctypes.c_int32),
        ("method", ctypes.c_int32),
        ("use_coloring", ctypes.c_int32),
        ("powmode", ctypes.c_int32),
        ("rtol", ctypes.c_double),
        ("t0", ctypes.c_double),
        ("h0", ctypes.c_double),
        ("natol", ctypes.c_int32),
        ("has_options", ctypes.c_int32),
        ("atol", ctypes.c_double * 64),
        ("max_nonlinear_solver_iterations", ctypes.c_int32),
        ("max_error_test_failures", ctypes.c_int32),
        ("max_nonlinear_solver_failures", ctypes.c_int32),
        ("update_jacobian_after_steps", ctypes.c_int32),
        ("update_rhs_jacobian_after_steps", ctypes.c_int32),
        ("pad0", ctypes.c_int32),
        ("nonlinear_solver_tolerance", ctypes.c_double),
        ("min_timestep", ctypes.c_double),
        ("max_timestep_growth", ctypes.c_double),
        ("min_timestep_growth", ctypes.c_double),
        ("max_timestep_shrink", ctypes.c_double),
        ("min_timestep_shrink", ctypes.c_double),
        ("threshold_to_update_jacobian", ctypes.c_double),
        ("threshold_to_update_rhs_jacobian", ctypes.c_double),
        ("pi_control_proportional", ctypes.c_double),
        ("pi_control_integral", ctypes.c_double),
        ("sens", ctypes.c_int32),
        ("sens_natol", ctypes.c_int32),
        ("sens_rtol", ctypes.c_double),
        ("sens_atol", ctypes.c_double * 64),
    ]


def build(force=False):
    """Compile the oracle with the recipe committed in oracle/Makefile (make rebuilds only when a source is newer than
    the library, so a stale library never outlives an edit of the restatement)."""
    import shutil
    if force or not os.path.exists(_LIB_PATH) or shutil.which("make"):
        subprocess.check_call(["make", "-s", "-C", _HERE] + (["-B"] if force else []))
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        dp = ctypes.POINTER(ctypes.c_double)
        ip = ctypes.POINTER(ctypes.c_int64)
        for name in ("orc_solve_dense", "orc_harness", "orc_harness_tstop"):
            f = getattr(L, name)
            f.restype = ctypes.c_int
            f.argtypes = [ctypes.POINTER(ProblemDesc), dp, ctypes.c_int, dp, ctypes.c_int, dp, ip, dp]
        L.orc_harness_sens.restype = ctypes.c_int
        L.orc_harness_sens.argtypes = [ctypes.POINTER(ProblemDesc), dp, ctypes.c_int, dp, ctypes.c_int, dp, dp, ip, dp]
        L.orc_batch_solve_dense_sens.restype = ctypes.c_int
        L.orc_batch_solve_dense_sens.argtypes = [ctypes.POINTER(ProblemDesc), dp, ctypes.c_int, ctypes.c_int64, dp, ctypes.c_int,
                                                 ctypes.c_int, dp, dp, ip, ctypes.POINTER(ctypes.c_int32)]
        L.orc_batch_solve_ragged.restype = ctypes.c_int
        L.orc_batch_solve_ragged.argtypes = [ctypes.POINTER(ProblemDesc), dp, ctypes.c_int, ctypes.c_int64, ctypes.c_double, ctypes.c_int,
                                             ctypes.c_int, dp, dp, ctypes.POINTER(ctypes.c_int32), ip, ctypes.POINTER(ctypes.c_int32), dp]
        L.orc_batch_solve_dense.restype = ctypes.c_int
        L.orc_batch_solve_dense.argtypes = [
            ctypes.POINTER(ProblemDesc), dp, ctypes.c_int, ctypes.c_int64, dp, ctypes.c_int, ctypes.c_int,
            dp, ip, ctypes.POINTER(ctypes.c_int32)]
        L.orc_batch_solve_dense_roots.restype = ctypes.c_int
        L.orc_batch_solve_dense_roots.argtypes = list(L.orc_batch_solve_dense.argtypes) + [dp]
        L.orc_model_dims.argtypes = [ctypes.c_int] + [ctypes.POINTER(ctypes.c_int)] * 3
        L.orc_pow.restype = ctypes.c_double
        L.orc_pow.argtypes = [ctypes.c_double, ctypes.c_double, ctypes.c_int]
        L.orc_powi.restype = ctypes.c_double
        L.orc_powi.argtypes = [ctypes.c_double, ctypes.c_int]
        L.orc_squared_norm.restype = ctypes.c_double
        L.orc_squared_norm.argtypes = [dp, dp, dp, ctypes.c_double, ctypes.c_int]
        L.orc_load_model_plugin.restype = ctypes.c_int
        L.orc_load_model_plugin.argtypes = [ctypes.c_char_p]
        L.orc_lu_solve.argtypes = [dp, ctypes.c_int, dp]
        L.orc_lu_factor.argtypes = [dp, ctypes.c_int, dp, ctypes.POINTER(ctypes.c_int32)]
        L.orc_num_threads.restype = ctypes.c_int
        L.orc_model_nout.restype = ctypes.c_int
        L.orc_model_nout.argtypes = [ctypes.c_int]
        L.orc_model_root.restype = ctypes.c_int
        L.orc_model_root.argtypes = [ctypes.c_int, dp, dp, ctypes.c_double, dp]
        L.orc_math.restype = ctypes.c_double
        L.orc_math.argtypes = [ctypes.c_int, ctypes.c_double]
        L.orc_steps_after_first_root.restype = ctypes.c_int
        L.orc_steps_after_first_root.argtypes = [ctypes.POINTER(ProblemDesc), dp, ctypes.c_int, ctypes.c_double,
                                                 ctypes.c_int, dp, ctypes.POINTER(ctypes.c_int)]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def model_dims(model):
    n, np_, hm = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    rc = lib().orc_model_dims(MODELS[model], ctypes.byref(n), ctypes.byref(np_), ctypes.byref(hm))
    assert rc == 0
    return n.value, np_.value, bool(hm.value)


def make_desc(model, method="bdf", rtol=1e-6, atol=1e-6, t0=0.0, h0=1.0, use_coloring=False,
              powmode=0, options=None, sens=False, sens_rtol=None, sens_atol=None):
    d = ProblemDesc()
    d.model_id = MODELS[model] if isinstance(model, str) else int(model)
    d.method = METHODS[method] if isinstance(method, str) else int(method)
    d.use_coloring = int(use_coloring)
    d.powmode = int(powmode)
    d.rtol, d.t0, d.h0 = float(rtol), float(t0), float(h0)
    atol = np.atleast_1d(np.asarray(atol, dtype=np.float64))
    d.natol = len(atol)
    for i, a in enumerate(atol):
        d.atol[i] = a
    d.sens = int(bool(sens))
    if sens and sens_rtol is not None and sens_atol is not None:
        sa = np.atleast_1d(np.asarray(sens_atol, dtype=np.float64))
        d.sens_natol, d.sens_rtol = len(sa), float(sens_rtol)
        for i, a in enumerate(sa):
            d.sens_atol[i] = a
    if options:
        d.has_options = 1
        defaults = dict(
            max_nonlinear_solver_iterations=10, max_error_test_failures=40, max_nonlinear_solver_failures=50,
            update_jacobian_after_steps=20, update_rhs_jacobian_after_steps=50,
            nonlinear_solver_tolerance=0.2, min_timestep=1e-13, max_timestep_growth=2.0,
            min_timestep_growth=2.0, max_timestep_shrink=0.9, min_timestep_shrink=0.5,
            threshold_to_update_jacobian=0.3, threshold_to_update_rhs_jacobian=0.2,
            pi_control_proportional=0.0, pi_control_integral=0.5)
        defaults.update(options)
        for k, v in defaults.items():
            setattr(d, k, v)
    return d


def _stats_dict(stats):
    return {name: int(stats[i]) for i, name in enumerate(S_NAMES)}


def _run(fn, desc, p, ts, rows=None):
    n, np_, _ = _dims_by_id(desc.model_id)
    n = rows or n
    p = np.ascontiguousarray(p, dtype=np.float64)
    assert p.size == np_, (p.size, np_)
    ts = np.ascontiguousarray(ts, dtype=np.float64)
    out = np.full((len(ts), n), np.nan)
    stats = np.zeros(S_COUNT, dtype=np.int64)
    fin = np.zeros(3)
    rc = fn(ctypes.byref(desc), _dp(p), int(p.size), _dp(ts), len(ts), _dp(out),
            stats.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), _dp(fin))
    return rc, out, _stats_dict(stats), dict(t=fin[0], h=fin[1], order=int(fin[2]))


def _dims_by_id(model_id):
    n, np_, hm = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    rc = lib().orc_model_dims(int(model_id), ctypes.byref(n), ctypes.byref(np_), ctypes.byref(hm))
    assert rc == 0
    return n.value, np_.value, bool(hm.value)


def solve_dense(desc, p, t_eval):
    """problem.<method>().solve_dense(t_eval) -> (rc, ys[nt, n], stats, final)"""
    return _run(lib().orc_solve_dense, desc, p, t_eval, rows=model_nout(desc.model_id))


def harness(desc, p, t_points, use_tstop=False):
    """The reference's test_ode_solver() loop -> (rc, ys[npts, n], stats, final)"""
    return _run(lib().orc_harness_tstop if use_tstop else lib().orc_harness, desc, p, t_points)


def harness_sens(desc, p, t_points):
    """test_ode_solver(.., solve_for_sensitivities=True) -> (rc, ys[npts, n], sens[npts, np, n], stats, final)"""
    n, np_, _ = _dims_by_id(desc.model_id)
    p = np.ascontiguousarray(p, dtype=np.float64)
    ts = np.ascontiguousarray(t_points, dtype=np.float64)
    out = np.full((len(ts), n), np.nan)
    sens = np.full((len(ts), np_, n), np.nan)
    stats = np.zeros(S_COUNT, dtype=np.int64)
    fin = np.zeros(3)
    rc = lib().orc_harness_sens(ctypes.byref(desc), _dp(p), int(p.size), _dp(ts), len(ts), _dp(out), _dp(sens),
                                stats.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), _dp(fin))
    return rc, out, sens, _stats_dict(stats), dict(t=fin[0], h=fin[1], order=int(fin[2]))


def batch_solve_dense_sens(desc, params, t_eval, nthreads=0):
    """solve_dense_sensitivities for every row of params -> (ys[B, nt, n], sens[B, nt, np, n], stats[B, 16], status[B])"""
    n, np_, _ = _dims_by_id(desc.model_id)
    params = np.ascontiguousarray(params, dtype=np.float64)
    B = params.shape[0]
    ts = np.ascontiguousarray(t_eval, dtype=np.float64)
    ys = np.full((B, len(ts), n), np.nan)
    sens = np.full((B, len(ts), np_, n), np.nan)
    stats = np.zeros((B, S_COUNT), dtype=np.int64)
    status = np.zeros(B, dtype=np.int32)
    rc = lib().orc_batch_solve_dense_sens(ctypes.byref(desc), _dp(params), np_, B, _dp(ts), len(ts), int(nthreads), _dp(ys), _dp(sens),
                                          stats.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)),
                                          status.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)))
    assert rc == 0
    return ys, sens, stats, status


def batch_solve_ragged(desc, params, final_time, max_cols=4096, nthreads=0):
    """OdeSolverMethod::solve(final_time) for every row of params -> (ts[B, max_cols], ys[B, max_cols, nrow], ncols[B],
    stats[B, 16], status[B], roots[B, 2] = (t_root, root index)); instance b's columns are the first ncols[b]."""
    n, np_, _ = _dims_by_id(desc.model_id)
    nrow = lib().orc_model_nout(int(desc.model_id)) or n
    params = np.ascontiguousarray(params, dtype=np.float64).reshape(-1, max(np_, 1))
    B = params.shape[0]
    ts = np.full((B, max_cols), np.nan)
    ys = np.full((B, max_cols, nrow), np.nan)
    ncols = np.zeros(B, dtype=np.int32)
    stats = np.zeros((B, S_COUNT), dtype=np.int64)
    status = np.zeros(B, dtype=np.int32)
    roots = np.zeros((B, 2))
    ip32 = ctypes.POINTER(ctypes.c_int32)
    rc = lib().orc_batch_solve_ragged(ctypes.byref(desc), _dp(params), np_, B, float(final_time), int(max_cols), int(nthreads), _dp(ts),
                                      _dp(ys), ncols.ctypes.data_as(ip32), stats.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)),
                                      status.ctypes.data_as(ip32), _dp(roots))
    assert rc == 0
    return ts, ys, ncols, stats, status, roots


def greedy_coloring(non_zeros, n):
    """nonzeros2graph + color_graph_greedy -> 1-based colour of every column."""
    rows = np.ascontiguousarray([ij[0] for ij in non_zeros], dtype=np.int32)
    cols = np.ascontiguousarray([ij[1] for ij in non_zeros], dtype=np.int32)
    out = np.zeros(n, dtype=np.int32)
    ip = ctypes.POINTER(ctypes.c_int32)
    rc = lib().orc_greedy_coloring(rows.ctypes.data_as(ip), cols.ctypes.data_as(ip), len(rows), int(n), out.ctypes.data_as(ip))
    assert rc == 0
    return out.tolist()


def residual_known_answer(desc, p, c, h, vec, x, t=0.0):
    """BdfCallable / SdirkCallable with c, (h) and psi - y0 / phi set directly -> (rc, F(x)[n], A[n, n])."""
    n, np_, _ = _dims_by_id(desc.model_id)
    p = np.ascontiguousarray(p, dtype=np.float64)
    vec = np.ascontiguousarray(vec, dtype=np.float64)
    x = np.ascontiguousarray(x, dtype=np.float64)
    F = np.zeros(n)
    A = np.zeros((n, n))
    L = lib()
    dp = ctypes.POINTER(ctypes.c_double)
    L.orc_residual_known_answer.restype = ctypes.c_int
    L.orc_residual_known_answer.argtypes = [ctypes.POINTER(ProblemDesc), dp, ctypes.c_int, ctypes.c_double, ctypes.c_double,
                                            dp, dp, ctypes.c_double, dp, dp]
    rc = L.orc_residual_known_answer(ctypes.byref(desc), _dp(p), int(p.size), c, h, _dp(vec), _dp(x), t, _dp(F), _dp(A))
    return rc, F, A.T.copy()          # A comes back column-major


def steps_after_first_root(desc, p, tstop, nsteps):
    """The reference's test_ball_bounce loop (ode_solver/mod.rs:1024-1080) -> (rc, t[k], y[k, n]) for the k <= nsteps
    internal steps taken after the first root + reset."""
    n, np_, _ = _dims_by_id(desc.model_id)
    p = np.ascontiguousarray(p, dtype=np.float64)
    rows = np.full((nsteps, 1 + n), np.nan)
    taken = ctypes.c_int()
    rc = lib().orc_steps_after_first_root(ctypes.byref(desc), _dp(p), int(p.size), ctypes.c_double(tstop), int(nsteps),
                                          _dp(rows), ctypes.byref(taken))
    return rc, rows[:taken.value, 0].copy(), rows[:taken.value, 1:].copy()


def batch_solve_dense(desc, params, t_eval, nthreads=0):
    """params[B, np] instance-major -> (ys[B, nt, n], stats[B, 16], status[B])"""
    n, np_, _ = _dims_by_id(desc.model_id)
    params = np.ascontiguousarray(params, dtype=np.float64).reshape(-1, max(np_, 1))
    B = params.shape[0]
    t_eval = np.ascontiguousarray(t_eval, dtype=np.float64)
    out = np.full((B, len(t_eval), model_nout(desc.model_id)), np.nan)
    stats = np.zeros((B, S_COUNT), dtype=np.int64)
    status = np.zeros(B, dtype=np.int32)
    rc = lib().orc_batch_solve_dense(
        ctypes.byref(desc), _dp(params), np_, B, _dp(t_eval), len(t_eval), int(nthreads), _dp(out),
        stats.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)),
        status.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)))
    assert rc == 0
    return out, stats, status


def batch_solve_dense_roots(desc, params, t_eval, nthreads=0):
    """The same for models with root (event) functions -> (ys, stats, status, t_root[B], root_idx[B], ncols[B]):
    an instance that stops on a root has its state at the root in column ncols - 1 and NaN behind it."""
    n, np_, _ = _dims_by_id(desc.model_id)
    params = np.ascontiguousarray(params, dtype=np.float64).reshape(-1, max(np_, 1))
    B = params.shape[0]
    t_eval = np.ascontiguousarray(t_eval, dtype=np.float64)
    out = np.full((B, len(t_eval), model_nout(desc.model_id)), np.nan)
    stats = np.zeros((B, S_COUNT), dtype=np.int64)
    status = np.zeros(B, dtype=np.int32)
    roots = np.zeros((B, 3))
    rc = lib().orc_batch_solve_dense_roots(
        ctypes.byref(desc), _dp(params), np_, B, _dp(t_eval), len(t_eval), int(nthreads), _dp(out),
        stats.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)),
        status.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), _dp(roots))
    assert rc == 0
    return out, stats, status, roots[:, 0].copy(), roots[:, 1].astype(np.int32), roots[:, 2].astype(np.int32)


def load_user_model(source, kind="functor", struct="UserModel", name=None):
    """Compile a USER equation set for the host and register it with the oracle -> model name usable with make_desc().
    `source` is the same text the CUDA library compiles for the device (diffsol_b200.OdeBuilder.rhs_implicit_source):
    kind "functor": a struct `struct` with the interface of csrc/dsb_models.h; kind "diffsl": the DiffSL symbol table
    (csrc/dsb_diffsl_adapter.h).  TEST INFRASTRUCTURE: the oracle stays the checker of the CUDA path."""
    import hashlib
    build()
    digest = hashlib.sha256((kind + "\0" + struct + "\0" + source).encode()).hexdigest()[:16]
    name = name or "user_" + digest
    if name in MODELS:
        return name
    out_dir = os.path.join(_HERE, "_ref", "plugins")
    os.makedirs(out_dir, exist_ok=True)
    src = os.path.join(out_dir, "m_%s.cpp" % digest)
    so = os.path.join(out_dir, "m_%s.so" % digest)
    csrc = os.path.join(os.path.dirname(_HERE), "diffsol_b200", "csrc")
    if kind == "diffsl":
        body = '#include "%s/dsb_math.h"\n%s\n#include "%s/dsb_diffsl_adapter.h"\ntypedef DsbDiffslModel OrcUserModel;\n' % (csrc, source, csrc)
    else:
        body = '#include "%s/dsb_math.h"\n%s\ntypedef %s OrcUserModel;\n' % (csrc, source, struct)
    with open(src, "w") as f:
        f.write('#include "%s/dsb_oracle.hpp"\n%s\nextern "C" void orc_plugin_model(orc::Model* m) { *m = orc::make_model<OrcUserModel>(); }\n'
                % (_HERE, body))
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call([os.environ.get("CXX", "g++"), "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-mfma",
                               "-shared", src, "-o", so])
    mid = lib().orc_load_model_plugin(so.encode())
    assert mid >= 1000, "oracle model plugin failed to load: " + so
    MODELS[name] = mid
    return name


def num_threads():
    return lib().orc_num_threads()


def model_nout(model):
    """Rows of a solve_dense column: the outputs of the model's out function, else its states."""
    return lib().orc_model_nout(MODELS[model] if isinstance(model, str) else int(model))


def model_root(model, y, p, t=0.0):
    """Root (event) functions of a model at (y, p, t) -> g[nroots]."""
    y = np.ascontiguousarray(y, dtype=np.float64)
    p = np.ascontiguousarray(p, dtype=np.float64)
    g = np.zeros(8)
    nr = lib().orc_model_root(MODELS[model], _dp(y), _dp(p), float(t), _dp(g))
    assert nr >= 0
    return g[:nr].copy()


def math_fn(name, x):
    """dsb_exp / dsb_log / dsb_tanh / dsb_asinh of csrc/dsb_math.h (shared by the oracle and the kernels)."""
    which = {"exp": 0, "log": 1, "tanh": 2, "asinh": 3}[name]
    return np.array([lib().orc_math(which, float(v)) for v in np.atleast_1d(x)])
