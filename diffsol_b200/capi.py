"""ctypes binding of include/diffsol_b200.h (the C-ABI library built by diffsol_b200/build.py).

This is the only place the package talks to native code.  There is no CPU fallback: if the library
cannot be loaded, or no CUDA device is present, calls raise `DiffsolB200Error`.
"""
import ctypes
import os

from . import build as _build

DSB_OK, DSB_ERR, DSB_BAD_ARG = 0, -1, -2
DSB_NSTATS = 16

METHODS = {"bdf": 0, "tr_bdf2": 1, "esdirk34": 2}
MODELS = {
    "exp_decay": 0, "exp_decay_algebraic": 1, "robertson_dae": 2, "robertson_ode": 3,
    "robertson_ode_g3": 4, "dydt_y2": 5, "gaussian_decay": 6, "van_der_pol": 7, "van_der_pol_scaled": 8,
    "heat1d_dae_256": 9, "heat1d_dae_32": 10, "spm": 11, "spm99": 12, "exp_decay_root": 13, "spm_stop": 14, "spm99_stop": 15, "heat1d_dae_32_bc": 16, "exp_decay_reset": 17, "heat2d_10": 18, "ball_bounce": 19, "exp_decay_two_roots": 20, "spm_cycle": 21, "exp_decay_algebraic_reset": 22,
}
STAT_NAMES = [
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
STATUS_NAMES = {
    0: "Ok", 1: "StepSizeTooSmall", 2: "TooManyErrorTestFailures", 3: "TooManyNonlinearSolverFailures",
    4: "StopTimeBeforeCurrentTime", 5: "StopTimeAtCurrentTime", 6: "InitialConditionDidNotConverge",
    7: "LinesearchFailed", 8: "InterpolationTimeAfterCurrentTime", 9: "LuSolveFailed",
}


class DiffsolB200Error(RuntimeError):
    pass


class Options(ctypes.Structure):
    """dsb_options: OdeSolverOptions + InitialConditionSolverOptions + Bdf/SdirkConfig clamps."""
    _fields_ = [
        ("max_nonlinear_solver_iterations", ctypes.c_int32),
        ("max_error_test_failures", ctypes.c_int32),
        ("max_nonlinear_solver_failures", ctypes.c_int32),
        ("update_jacobian_after_steps", ctypes.c_int32),
        ("update_rhs_jacobian_after_steps", ctypes.c_int32),
        ("ic_max_linesearch_iterations", ctypes.c_int32),
        ("ic_max_newton_iterations", ctypes.c_int32),
        ("ic_max_linear_solver_setups", ctypes.c_int32),
        ("ic_use_linesearch", ctypes.c_int32),
        ("reserved0", ctypes.c_int32),
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
        ("ic_step_reduction_factor", ctypes.c_double),
        ("ic_armijo_constant", ctypes.c_double),
    ]


_vp, _i32, _i64, _dbl = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_double
_pd, _pi32, _pi64 = ctypes.POINTER(_dbl), ctypes.POINTER(_i32), ctypes.POINTER(_i64)

# name -> (restype, argtypes); every symbol include/diffsol_b200.h declares
SIGNATURES = {
    "dsb_options_default": (None, [ctypes.POINTER(Options)]),
    "dsb_last_error": (ctypes.c_char_p, []),
    "dsb_version": (ctypes.c_char_p, []),
    "dsb_device_count": (ctypes.c_int, [ctypes.POINTER(ctypes.c_int)]),
    "dsb_problem_new": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(_vp)]),
    "dsb_problem_free": (ctypes.c_int, [_vp]),
    "dsb_problem_dims": (ctypes.c_int, [_vp, _pi32, _pi32, _pi32]),
    "dsb_problem_nout": (ctypes.c_int, [_vp, _pi32]),
    "dsb_problem_set_rtol": (ctypes.c_int, [_vp, _dbl]),
    "dsb_problem_set_atol": (ctypes.c_int, [_vp, _vp, _i32]),
    "dsb_problem_set_t0": (ctypes.c_int, [_vp, _dbl]),
    "dsb_problem_set_h0": (ctypes.c_int, [_vp, _dbl]),
    "dsb_problem_set_use_coloring": (ctypes.c_int, [_vp, _i32]),
    "dsb_problem_set_options": (ctypes.c_int, [_vp, ctypes.POINTER(Options)]),
    "dsb_problem_get_options": (ctypes.c_int, [_vp, ctypes.POINTER(Options)]),
    "dsb_problem_set_sensitivities": (ctypes.c_int, [_vp, _i32, _dbl, _vp, _i32]),
    "dsb_batch_new": (ctypes.c_int, [_vp, _i64, _i32, ctypes.POINTER(_vp)]),
    "dsb_batch_free": (ctypes.c_int, [_vp]),
    "dsb_batch_size": (_i64, [_vp]),
    "dsb_batch_set_execution": (ctypes.c_int, [_vp, _i32]),
    "dsb_model_library_build": (ctypes.c_int, [ctypes.c_char_p, _i32, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p]),
    "dsb_model_library_load": (ctypes.c_int, [ctypes.c_char_p, ctypes.POINTER(ctypes.c_int32)]),
    "dsb_batch_set_params_host": (ctypes.c_int, [_vp, _vp, _i64, _i32]),
    "dsb_batch_set_params_device": (ctypes.c_int, [_vp, _vp, _i64, _i32, _vp]),
    "dsb_batch_solve_dense": (ctypes.c_int, [_vp, _i32, _vp, _i32, _vp, _vp]),
    "dsb_batch_solve_dense_host": (ctypes.c_int, [_vp, _i32, _vp, _i32, _vp, _i32, _vp, _vp, _vp]),
    "dsb_batch_step_and_interpolate": (ctypes.c_int, [_vp, _i32, _vp, _i32, _vp, _vp]),
    "dsb_batch_step_and_interpolate_host": (ctypes.c_int, [_vp, _i32, _vp, _i32, _vp, _i32, _vp, _vp, _vp]),
    "dsb_batch_solve_dense_sensitivities": (ctypes.c_int, [_vp, _i32, _vp, _i32, _vp, _vp, _vp]),
    "dsb_batch_solve_count": (ctypes.c_int, [_vp, _i32, _dbl, _vp]),
    "dsb_batch_solve_offsets": (ctypes.c_int, [_vp, _vp]),
    "dsb_batch_solve_write": (ctypes.c_int, [_vp, _i32, _dbl, _vp, _vp, _vp]),
    "dsb_batch_solve_write_host": (ctypes.c_int, [_vp, _i32, _dbl, _vp, _vp]),
    "dsb_batch_step_and_interpolate_sensitivities": (ctypes.c_int, [_vp, _i32, _vp, _i32, _vp, _vp, _vp]),
    "dsb_batch_solve_dense_sensitivities_host": (ctypes.c_int, [_vp, _i32, _vp, _i32, _vp, _i32, _vp, _vp, _vp, _vp]),
    "dsb_batch_step_and_interpolate_sensitivities_host": (ctypes.c_int, [_vp, _i32, _vp, _i32, _vp, _i32, _vp, _vp, _vp, _vp]),
    "dsb_batch_get_stats": (ctypes.c_int, [_vp, _vp]),
    "dsb_batch_get_stats_device": (ctypes.c_int, [_vp, _vp, _vp]),
    "dsb_batch_get_status": (ctypes.c_int, [_vp, _vp]),
    "dsb_batch_get_final_state": (ctypes.c_int, [_vp, _vp, _vp, _vp]),
    "dsb_batch_device_views": (ctypes.c_int, [_vp, ctypes.POINTER(_vp), ctypes.POINTER(_vp)]),
    "dsb_batch_sum_stat": (ctypes.c_int, [_vp, _i32, _pi64]),
    "dsb_batch_last_kernel_ms": (ctypes.c_int, [_vp, ctypes.POINTER(ctypes.c_float)]),
    "dsb_batch_last_integrator_ms": (ctypes.c_int, [_vp, ctypes.POINTER(ctypes.c_float)]),
    "dsb_batch_last_launch_count": (ctypes.c_int, [_vp, _pi32]),
    "dsb_batch_debug_words": (ctypes.c_int, [_vp, _vp]),
    "dsb_batch_get_root_info": (ctypes.c_int, [_vp, _vp, _vp]),
    "dsb_lu_factor_batched": (ctypes.c_int, [_vp, _i32, _i64, _vp, _vp, _vp]),
    "dsb_lu_solve_batched": (ctypes.c_int, [_vp, _vp, _vp, _i32, _i64, _vp, _vp]),
    "dsb_lu_factor_instance_major": (ctypes.c_int, [_vp, _i32, _i64, _vp, _vp, _vp]),
    "dsb_lu_solve_instance_major": (ctypes.c_int, [_vp, _vp, _vp, _i32, _i64, _vp, _vp]),
}

_lib = None


def library_path():
    return _build.LIB


def lib():
    """Load the C-ABI library, building it first when it is missing or when the sources changed since it was built
    (build() compares a digest of csrc/ with the stamp of the last build and returns at once when they agree).  Without
    nvcc a stale or missing library is an error, not a silent fallback."""
    global _lib
    if _lib is None:
        path = _build.LIB
        if _build.have_nvcc():
            path = _build.build(force=bool(os.environ.get("DSB_REBUILD")))
        elif not os.path.exists(path):
            raise DiffsolB200Error("libdiffsol_b200.so is not built and nvcc is not available (python -m diffsol_b200.build)")
        elif not _build.is_current():
            raise DiffsolB200Error("libdiffsol_b200.so is older than diffsol_b200/csrc and nvcc is not available to rebuild it")
        L = ctypes.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            f = getattr(L, name)          # AttributeError here = header and library disagree
            f.restype, f.argtypes = res, args
        _lib = L
    return _lib


MODEL_SOURCE_KINDS = {"functor": 0, "diffsl": 1}
CSRC_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")


def load_model_source(source, kind="functor", struct="UserModel", cache_dir=None):
    """Compile a user equation set (source text, see include/diffsol_b200.h: dsb_model_library_build) into a model plugin
    -- cached by the digest of the text and of csrc/ -- load it and return its registered name for
    OdeBuilder.rhs_implicit().  Compilation runs nvcc (about a minute for a small system); a cached plugin loads at once."""
    import hashlib
    L = lib()
    h = hashlib.sha256()
    h.update((kind + "\0" + struct + "\0" + source + "\0" + _build._sources_digest()).encode())
    digest = h.hexdigest()[:20]
    name = "user_" + digest
    if name in MODELS:
        return name
    cache_dir = cache_dir or os.path.join(os.path.dirname(_build.LIB), "plugins")
    os.makedirs(cache_dir, exist_ok=True)
    src = os.path.join(cache_dir, name + ".h")
    so = os.path.join(cache_dir, name + ".so")
    if not os.path.exists(so):
        with open(src, "w") as f:
            f.write(source)
        tmp = so + ".tmp.%d" % os.getpid()
        check(L.dsb_model_library_build(src.encode(), MODEL_SOURCE_KINDS[kind], struct.encode(), CSRC_DIR.encode(), tmp.encode()))
        os.replace(tmp, so)
    mid = ctypes.c_int32()
    check(L.dsb_model_library_load(so.encode(), ctypes.byref(mid)))
    MODELS[name] = mid.value
    return name


def check(rc):
    if rc != DSB_OK:
        raise DiffsolB200Error("diffsol_b200 error %d: %s" % (rc, lib().dsb_last_error().decode()))


def device_count():
    c = ctypes.c_int(0)
    rc = lib().dsb_device_count(ctypes.byref(c))
    return c.value if rc == DSB_OK else 0


def require_device():
    c = ctypes.c_int(0)
    check(lib().dsb_device_count(ctypes.byref(c)))
    return c.value
