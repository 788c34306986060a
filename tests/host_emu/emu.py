"""ctypes wrapper around the single-lane HOST build of the product's lane kernels -- TEST INFRASTRUCTURE.

tests/host_emu/dsb_emu.cpp compiles diffsol_b200/csrc/*.cuh with g++ through cuda_shim.h and runs one lane of the
kernels per instance, so the kernels' control flow and arithmetic can be compared bit for bit with the oracle on a
machine without a GPU.  The product package never imports this; it has no host integrator.
"""
import ctypes
import hashlib
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
_CSRC = os.path.join(_ROOT, "diffsol_b200", "csrc")
_OUT = os.path.join(_HERE, "_build")
_LIB = os.path.join(_OUT, "libdsb_emu.so")
NSTATS = 16
METHODS = {"bdf": 0, "tr_bdf2": 1, "esdirk34": 2}
KERNELS = {"lane": 1, "band": 3, "warp": 4}


class Options(ctypes.Structure):
    _fields_ = [(k, ctypes.c_int32) for k in (
        "max_nonlinear_solver_iterations", "max_error_test_failures", "max_nonlinear_solver_failures",
        "update_jacobian_after_steps", "update_rhs_jacobian_after_steps", "ic_max_linesearch_iterations",
        "ic_max_newton_iterations", "ic_max_linear_solver_setups", "ic_use_linesearch", "reserved0")] + [
        (k, ctypes.c_double) for k in (
        "nonlinear_solver_tolerance", "min_timestep", "max_timestep_growth", "min_timestep_growth",
        "max_timestep_shrink", "min_timestep_shrink", "threshold_to_update_jacobian",
        "threshold_to_update_rhs_jacobian", "pi_control_proportional", "pi_control_integral",
        "ic_step_reduction_factor", "ic_armijo_constant")]


def _digest():
    h = hashlib.sha256()
    for d in (_CSRC, _HERE):
        for name in sorted(os.listdir(d)):
            path = os.path.join(d, name)
            if os.path.isfile(path) and name.endswith((".h", ".cuh", ".inc", ".cpp")):
                with open(path, "rb") as f:
                    h.update(name.encode()); h.update(f.read())
    return h.hexdigest()


def build(force=False):
    os.makedirs(_OUT, exist_ok=True)
    stamp = os.path.join(_OUT, "build.sha256")
    digest = _digest()
    if not force and os.path.exists(_LIB) and os.path.exists(stamp) and open(stamp).read().strip() == digest:
        return _LIB
    cxx = os.environ.get("CXX", "g++")
    subprocess.check_call([cxx, "-O1", "-std=c++17", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-mfma", "-shared",
                           os.path.join(_HERE, "dsb_emu.cpp"), "-o", _LIB])
    with open(stamp, "w") as f:
        f.write(digest)
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(build())
        dp = ctypes.POINTER(ctypes.c_double)
        L.emu_solve.restype = ctypes.c_int
        L.emu_solve.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double, dp, ctypes.c_int,
                                ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.POINTER(Options), dp,
                                ctypes.c_int64, dp, ctypes.c_int, ctypes.c_int, dp, ctypes.POINTER(ctypes.c_int64),
                                ctypes.POINTER(ctypes.c_int32), dp, ctypes.POINTER(ctypes.c_int32),
                                ctypes.POINTER(ctypes.c_int32)]
        L.dsb_options_default.argtypes = [ctypes.POINTER(Options)]
        L.emu_solve_ragged.restype = ctypes.c_int
        L.emu_solve_ragged.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_double, dp, ctypes.c_int, ctypes.c_double, ctypes.c_double,
                                       ctypes.POINTER(Options), dp, ctypes.c_int64, ctypes.c_double, ctypes.c_int, dp, dp,
                                       ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int32),
                                       ctypes.POINTER(ctypes.c_int32)]
        L.emu_solve_sens.restype = ctypes.c_int
        L.emu_solve_sens.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_double, dp, ctypes.c_int, ctypes.c_double, ctypes.c_double,
                                     ctypes.POINTER(Options), ctypes.c_double, dp, ctypes.c_int, dp, ctypes.c_int64, dp, ctypes.c_int,
                                     ctypes.c_int, dp, dp, ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int32)]
        L.emu_smem_band_lu.restype = ctypes.c_int
        L.emu_smem_band_lu.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, dp, dp, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def solve(model_id, n, np_, params, t_eval, method="bdf", kernel="lane", rtol=1e-6, atol=1e-6, t0=0.0, h0=1.0,
          use_coloring=False, options=None, free_running=False, nout=None):
    """-> dict(ys[B, nt, n], stats[B, 16], status[B], fin[B, 3], root_idx[B], ncols[B]); raises when the kernel does
    not exist for the model."""
    params = np.ascontiguousarray(params, dtype=np.float64).reshape(-1, max(np_, 1))
    B = params.shape[0]
    t_eval = np.ascontiguousarray(t_eval, dtype=np.float64)
    atol = np.ascontiguousarray(np.atleast_1d(np.asarray(atol, dtype=np.float64)))
    nt = len(t_eval)
    ys = np.full((B, nt, nout or n), np.nan)      # nout: rows of a column for models with an output function
    stats = np.zeros((B, NSTATS), dtype=np.int64)
    status = np.zeros(B, dtype=np.int32)
    fin = np.zeros((B, 3))
    root_idx = np.zeros(B, dtype=np.int32)
    ncols = np.zeros(B, dtype=np.int32)
    opt = None
    if options:
        opt = Options()
        lib().dsb_options_default(ctypes.byref(opt))
        for k, v in options.items():
            setattr(opt, k, v)
    ip32 = ctypes.POINTER(ctypes.c_int32)
    rc = lib().emu_solve(int(model_id), METHODS[method], KERNELS[kernel], float(rtol), _dp(atol), len(atol), float(t0),
                         float(h0), int(use_coloring), ctypes.byref(opt) if opt is not None else None, _dp(params), B,
                         _dp(t_eval), nt, int(free_running), _dp(ys), stats.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)),
                         status.ctypes.data_as(ip32), _dp(fin), root_idx.ctypes.data_as(ip32), ncols.ctypes.data_as(ip32))
    if rc != 0:
        raise RuntimeError("emu_solve: rc = %d (kernel %r not available for this model / method?)" % (rc, kernel))
    return dict(ys=ys, stats=stats, status=status, fin=fin, root_idx=root_idx, ncols=ncols)


def solve_sens(model_id, n, np_, params, t_eval, method="bdf", rtol=1e-6, atol=1e-6, t0=0.0, h0=1.0, sens_rtol=None, sens_atol=None,
               options=None, free_running=False):
    """The sensitivity instantiation of the on-chip BDF lane kernel (DsbWithSens<M>) -> dict(ys[B, nt, n],
    sens[B, nt, np, n], stats[B, 16], status[B])."""
    params = np.ascontiguousarray(params, dtype=np.float64).reshape(-1, np_)
    B = params.shape[0]
    t_eval = np.ascontiguousarray(t_eval, dtype=np.float64)
    atol = np.ascontiguousarray(np.atleast_1d(np.asarray(atol, dtype=np.float64)))
    sa = np.ascontiguousarray(np.atleast_1d(np.asarray(sens_atol if sens_atol is not None else [], dtype=np.float64)))
    nt = len(t_eval)
    ys = np.full((B, nt, n), np.nan)
    sens = np.full((B, nt, np_, n), np.nan)
    stats = np.zeros((B, NSTATS), dtype=np.int64)
    status = np.zeros(B, dtype=np.int32)
    opt = None
    if options:
        opt = Options()
        lib().dsb_options_default(ctypes.byref(opt))
        for k, v in options.items():
            setattr(opt, k, v)
    rc = lib().emu_solve_sens(int(model_id), METHODS[method], float(rtol), _dp(atol), len(atol), float(t0), float(h0),
                              ctypes.byref(opt) if opt is not None else None, float(sens_rtol or 0.0), _dp(sa), len(sa), _dp(params), B,
                              _dp(t_eval), nt, int(free_running), _dp(ys), _dp(sens),
                              stats.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), status.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)))
    if rc != 0:
        raise RuntimeError("emu_solve_sens: rc = %d" % rc)
    return dict(ys=ys, sens=sens, stats=stats, status=status)


def solve_ragged(model_id, n, np_, params, final_time, method="bdf", rtol=1e-6, atol=1e-6, t0=0.0, h0=1.0, nout=None, max_cols=4096):
    """OdeSolverMethod::solve(final_time) through the DsbRagged<M> instantiation of the on-chip lane kernels ->
    dict(ts[B, max_cols], ys[B, max_cols, nout], ncols[B], stats[B, 16], status[B], root_idx[B])."""
    params = np.ascontiguousarray(params, dtype=np.float64).reshape(-1, max(np_, 1))
    B = params.shape[0]
    atol = np.ascontiguousarray(np.atleast_1d(np.asarray(atol, dtype=np.float64)))
    ts = np.full((B, max_cols), np.nan)
    ys = np.full((B, max_cols, nout or n), np.nan)
    ncols = np.zeros(B, dtype=np.int32)
    stats = np.zeros((B, NSTATS), dtype=np.int64)
    status = np.zeros(B, dtype=np.int32)
    root_idx = np.zeros(B, dtype=np.int32)
    ip32 = ctypes.POINTER(ctypes.c_int32)
    rc = lib().emu_solve_ragged(int(model_id), METHODS[method], float(rtol), _dp(atol), len(atol), float(t0), float(h0), None, _dp(params), B,
                                float(final_time), int(max_cols), _dp(ts), _dp(ys), ncols.ctypes.data_as(ip32),
                                stats.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), status.ctypes.data_as(ip32), root_idx.ctypes.data_as(ip32))
    if rc != 0:
        raise RuntimeError("emu_solve_ragged: rc = %d" % rc)
    return dict(ts=ts, ys=ys, ncols=ncols, stats=stats, status=status, root_idx=root_idx)


def greedy_coloring(non_zeros, n):
    """The product's host-side colouring (csrc/dsb_host_setup.h: greedy_coloring) -> 1-based colour of every column."""
    rows = np.ascontiguousarray([ij[0] for ij in non_zeros], dtype=np.int32)
    cols = np.ascontiguousarray([ij[1] for ij in non_zeros], dtype=np.int32)
    out = np.zeros(n, dtype=np.int32)
    ip = ctypes.POINTER(ctypes.c_int32)
    rc = lib().emu_greedy_coloring(rows.ctypes.data_as(ip), cols.ctypes.data_as(ip), len(rows), int(n), out.ctypes.data_as(ip))
    assert rc == 0
    return out.tolist()


def smem_band_lu(A, kl, ku, b, exact=False, mode=None):
    """SmemBandLU (the warp-per-instance kernel's band LU) on the band (kl, ku) of the dense matrix A -> (rc, x, nswaps)."""
    n = A.shape[0]
    Af = np.asfortranarray(A, dtype=np.float64)
    x = np.array(b, dtype=np.float64)
    nsw = ctypes.c_int(0)
    rc = lib().emu_smem_band_lu(n, kl, ku, _dp(Af), _dp(x), mode if mode is not None else (1 if exact else 0), ctypes.byref(nsw))
    return rc, x, nsw.value
