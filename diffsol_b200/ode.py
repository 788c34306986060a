"""Host-side mirror of the reference interface for the batched implicit path.

Names and argument meaning follow diffsol (paths relative to /root/reference/crates/diffsol/src):

    OdeBuilder().p(...).rtol(..).atol(..).t0(..).h0(..).use_coloring(..)      ode_solver/builder.rs:1447-1626
        .rhs_implicit(model).nbatch(B).build()          -> OdeSolverProblem    ode_solver/problem.rs:98-193
    problem.bdf() / problem.tr_bdf2() / problem.esdirk34() -> solver           ode_solver/problem.rs:320-328,649-655
    solver.solve_dense(t_eval) -> ys                                           ode_solver/method.rs:467-505
    solver.get_statistics()                                                    ode_solver/mod.rs:27-69

Differences that the batch imposes: the equations are one of the library's device functors (a Rust
closure or a DiffSL CPU JIT module cannot run inside a kernel; `rhs_implicit` takes the functor's
name), `p` is [nbatch, nparams] (instance-major, as the reference concatenates batched parameters,
test_models/exponential_decay.rs:297-304), results carry a leading batch axis, and an instance that
fails does not raise: it gets a per-instance status (the variant name of the reference's error enum)
and NaN outputs past the failure.
"""
import ctypes

import numpy as np

from . import capi


def _ptr(a):
    return ctypes.c_void_p(a.ctypes.data)


class OdeSolverProblem:
    def __init__(self, handle, model, n, nparams, has_mass, nbatch, params, device, nout=None):
        self._h = handle
        self.model = model
        self.nstates, self.nparams, self.has_mass = n, nparams, has_mass
        self.nout = n if nout is None else nout      # rows of a solve_dense column (the `out` function's outputs, else the states)
        self.nbatch = nbatch
        self.p = params
        self.device = device
        self.has_sens = False

    def __del__(self):
        if getattr(self, "_h", None):
            capi.lib().dsb_problem_free(self._h)
            self._h = None

    def _solver(self, method):
        return BatchedSolver(self, method)

    def bdf(self):
        return self._solver("bdf")

    def tr_bdf2(self):
        return self._solver("tr_bdf2")

    def esdirk34(self):
        return self._solver("esdirk34")

    def bdf_sens(self):
        """`problem.bdf_sens::<LS>()` (ode_solver/problem.rs:819-830): BDF with one forward sensitivity per parameter.  The
        builder must have been given sens_rtol / sens_atol or sensitivities(True)."""
        if not self.has_sens:
            raise ValueError("build the problem with OdeBuilder.sens_rtol(..).sens_atol(..) or .sensitivities()")
        return self._solver("bdf")

    def tr_bdf2_sens(self):
        """`problem.tr_bdf2_sens::<LS>()` (ode_solver/problem.rs): TR-BDF2 with forward sensitivities."""
        if not self.has_sens:
            raise ValueError("build the problem with OdeBuilder.sens_rtol(..).sens_atol(..) or .sensitivities()")
        return self._solver("tr_bdf2")

    def esdirk34_sens(self):
        """`problem.esdirk34_sens::<LS>()`: ESDIRK34 with forward sensitivities."""
        if not self.has_sens:
            raise ValueError("build the problem with OdeBuilder.sens_rtol(..).sens_atol(..) or .sensitivities()")
        return self._solver("esdirk34")


class OdeBuilder:
    """Fluent builder; defaults are the reference's (builder.rs:112-140): t0=0, h0=1, rtol=1e-6, atol=[1e-6]."""

    def __init__(self):
        self._model = None
        self._p = None
        self._rtol, self._atol, self._t0, self._h0 = 1e-6, [1e-6], 0.0, 1.0
        self._coloring = False
        self._nbatch = None
        self._device = 0
        self._opts = {}
        self._sens, self._sens_rtol, self._sens_atol = False, None, None

    def rhs_implicit(self, model):
        if model not in capi.MODELS:
            raise ValueError("unknown equation set %r; available: %s" % (model, sorted(capi.MODELS)))
        self._model = model
        return self

    def rhs_implicit_source(self, source, kind="functor", struct="UserModel"):
        """User equations as SOURCE TEXT (the counterpart of OdeBuilder::rhs_implicit with closures, builder.rs:192-200, and
        of a DiffSL module, ode_equations/diffsl.rs): kind "functor" = a struct `struct` with the closure signatures as
        DSB_HD static functions (csrc/dsb_models.h), kind "diffsl" = the DiffSL symbol table as C (csrc/dsb_diffsl_adapter.h).
        nvcc compiles it for sm_100a at run time into the library's own kernel families (capi.load_model_source)."""
        self._model = capi.load_model_source(source, kind=kind, struct=struct)
        return self

    def p(self, p):
        self._p = np.asarray(p, dtype=np.float64)
        return self

    def rtol(self, v):
        self._rtol = float(v); return self

    def atol(self, v):
        self._atol = [float(x) for x in np.atleast_1d(v)]; return self

    def t0(self, v):
        self._t0 = float(v); return self

    def h0(self, v):
        self._h0 = float(v); return self

    def use_coloring(self, v):
        self._coloring = bool(v); return self

    def sens_rtol(self, v):
        """OdeBuilder::sens_rtol (builder.rs:1454-1464); with sens_atol the sensitivities join the error test."""
        self._sens, self._sens_rtol = True, float(v); return self

    def sens_atol(self, v):
        """OdeBuilder::sens_atol (builder.rs:1466-1477): one entry broadcasts, else one per state."""
        self._sens, self._sens_atol = True, [float(x) for x in np.atleast_1d(v)]; return self

    def sensitivities(self, on=True):
        """Integrate the sensitivities without putting them into the error test (turn_off_sensitivities_error_control)."""
        self._sens = bool(on); return self

    def nbatch(self, b):
        self._nbatch = int(b); return self

    def device(self, d):
        self._device = int(d); return self

    def ode_options(self, **kw):
        """Fields of OdeSolverOptions / BdfConfig / SdirkConfig / InitialConditionSolverOptions by name."""
        self._opts.update(kw); return self

    def build(self):
        if self._model is None:
            raise ValueError("rhs_implicit(model) is required")
        L = capi.lib()
        h = ctypes.c_void_p()
        capi.check(L.dsb_problem_new(capi.MODELS[self._model], ctypes.byref(h)))
        n, npar, hm = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
        capi.check(L.dsb_problem_dims(h, ctypes.byref(n), ctypes.byref(npar), ctypes.byref(hm)))
        nout = ctypes.c_int32()
        capi.check(L.dsb_problem_nout(h, ctypes.byref(nout)))
        try:
            capi.check(L.dsb_problem_set_rtol(h, self._rtol))
            atol = np.asarray(self._atol, dtype=np.float64)
            capi.check(L.dsb_problem_set_atol(h, _ptr(atol), len(atol)))
            capi.check(L.dsb_problem_set_t0(h, self._t0))
            capi.check(L.dsb_problem_set_h0(h, self._h0))
            capi.check(L.dsb_problem_set_use_coloring(h, int(self._coloring)))
            if self._sens:
                if (self._sens_rtol is None) != (self._sens_atol is None):
                    raise ValueError("sens_rtol and sens_atol go together")
                sa = np.asarray(self._sens_atol or [], dtype=np.float64)
                capi.check(L.dsb_problem_set_sensitivities(h, 1, self._sens_rtol or 0.0, _ptr(sa) if len(sa) else None, len(sa)))
            if self._opts:
                o = capi.Options()
                capi.check(L.dsb_problem_get_options(h, ctypes.byref(o)))
                for k, v in self._opts.items():
                    if not hasattr(o, k):
                        raise ValueError("unknown option %r" % k)
                    setattr(o, k, v)
                capi.check(L.dsb_problem_set_options(h, ctypes.byref(o)))
            p = self._p if self._p is not None else np.zeros((0,))
            if npar.value == 0:
                nb = self._nbatch or 1
                p = np.zeros((nb, 0))
            else:
                p = np.ascontiguousarray(p, dtype=np.float64)
                if p.ndim == 1:
                    if p.size != npar.value:
                        raise ValueError("p has %d entries, the equations take %d" % (p.size, npar.value))
                    p = np.tile(p, (self._nbatch or 1, 1))
                if p.ndim != 2 or p.shape[1] != npar.value:
                    raise ValueError("p must be [nbatch, %d]" % npar.value)
                if self._nbatch is not None and p.shape[0] != self._nbatch:
                    raise ValueError("p has %d rows, nbatch is %d" % (p.shape[0], self._nbatch))
        except Exception:
            L.dsb_problem_free(h)
            raise
        prob = OdeSolverProblem(h, self._model, n.value, npar.value, bool(hm.value), p.shape[0],
                                np.ascontiguousarray(p), self._device, nout=nout.value)
        prob.has_sens = self._sens
        return prob


class BatchedSolver:
    """`problem.<method>::<LS>()` over the batch; owns the device state (dsb_batch)."""

    def __init__(self, problem, method):
        self.problem = problem
        self.method = capi.METHODS[method]
        capi.require_device()
        L = capi.lib()
        b = ctypes.c_void_p()
        capi.check(L.dsb_batch_new(problem._h, problem.nbatch, problem.device, ctypes.byref(b)))
        self._b = b
        self._stats = None
        self._status = None

    def __del__(self):
        if getattr(self, "_b", None):
            capi.lib().dsb_batch_free(self._b)
            self._b = None

    def step_and_interpolate(self, t_points):
        """The stepping loop of the reference's tests (ode_solver/mod.rs:132-141), no stop time:
        for each point: while |t| < |t_k|: step(); then interpolate(t_k).  -> ys[nbatch, npts, nstates]."""
        return self._solve_host(t_points, "dsb_batch_step_and_interpolate_host")

    def solve_dense(self, t_eval):
        """-> ys[nbatch, nt, nout] (host): the states, or the outputs of the equations' `out` function when they have one
        (problem.nout).  Per-instance statistics/status via get_statistics()/status()."""
        return self._solve_host(t_eval, "dsb_batch_solve_dense_host")

    def _solve_host(self, t_eval, entry):
        pr = self.problem
        t_eval = np.ascontiguousarray(t_eval, dtype=np.float64)
        nt = len(t_eval)
        ys = np.empty((pr.nbatch, nt, pr.nout))
        stats = np.empty((pr.nbatch, capi.DSB_NSTATS), dtype=np.int64)
        status = np.empty(pr.nbatch, dtype=np.int32)
        capi.check(getattr(capi.lib(), entry)(
            self._b, self.method, _ptr(pr.p) if pr.nparams else None, pr.nparams, _ptr(t_eval), nt,
            _ptr(ys), _ptr(stats), _ptr(status)))
        self._stats, self._status = stats, status
        return ys

    def solve(self, final_time):
        """`solve(final_time)` (ode_solver/method.rs:227-258): every internal step of every instance.  -> (ys[total, nout],
        ts[total], offsets[nbatch + 1]): instance b's columns are rows offsets[b] : offsets[b + 1].  Two passes on the device
        (count, then write: dsb_batch_solve_count / dsb_batch_solve_write_host)."""
        pr = self.problem
        L = capi.lib()
        self.set_params()
        total = ctypes.c_int64()
        capi.check(L.dsb_batch_solve_count(self._b, self.method, float(final_time), ctypes.byref(total)))
        offsets = np.empty(pr.nbatch + 1, dtype=np.int64)
        capi.check(L.dsb_batch_solve_offsets(self._b, _ptr(offsets)))
        ts = np.empty(total.value)
        ys = np.empty((total.value, pr.nout))
        capi.check(L.dsb_batch_solve_write_host(self._b, self.method, float(final_time), _ptr(ts), _ptr(ys)))
        self._stats = self._status = None
        return ys, ts, offsets

    def solve_dense_sensitivities(self, t_eval, free_running=False):
        """`solve_dense_sensitivities(t_eval)` (ode_solver/sensitivities.rs:114-262) -> (ys[nbatch, nt, nstates],
        sens[nbatch, nt, nparams, nstates]); free_running = the step()/interpolate()/interpolate_sens() loop of the
        reference's tests instead (ode_solver/mod.rs:104-194)."""
        pr = self.problem
        t_eval = np.ascontiguousarray(t_eval, dtype=np.float64)
        nt = len(t_eval)
        ys = np.empty((pr.nbatch, nt, pr.nstates))
        sens = np.empty((pr.nbatch, nt, pr.nparams, pr.nstates))
        stats = np.empty((pr.nbatch, capi.DSB_NSTATS), dtype=np.int64)
        status = np.empty(pr.nbatch, dtype=np.int32)
        entry = "dsb_batch_step_and_interpolate_sensitivities_host" if free_running else "dsb_batch_solve_dense_sensitivities_host"
        capi.check(getattr(capi.lib(), entry)(self._b, self.method, _ptr(pr.p), pr.nparams, _ptr(t_eval), nt,
                                              _ptr(ys), _ptr(sens), _ptr(stats), _ptr(status)))
        self._stats, self._status = stats, status
        return ys, sens

    def solve_dense_sensitivities_device(self, t_eval, ys_dev_ptr, sens_dev_ptr, stream=None):
        """Device-resident variant: ys -> [nt][nstates][nbatch], sens -> [nt][nparams][nstates][nbatch] doubles, asynchronous."""
        t_eval = np.ascontiguousarray(t_eval, dtype=np.float64)
        capi.check(capi.lib().dsb_batch_solve_dense_sensitivities(self._b, self.method, _ptr(t_eval), len(t_eval), ctypes.c_void_p(ys_dev_ptr),
                                                                  ctypes.c_void_p(sens_dev_ptr), ctypes.c_void_p(stream or 0)))
        self._stats = self._status = None

    def solve_dense_device(self, t_eval, ys_dev_ptr, stream=None, params_dev_ptr=None):
        """Device-resident variant: ys_dev_ptr -> [nt][nout][nbatch] doubles (batch-major), asynchronous."""
        pr = self.problem
        L = capi.lib()
        if params_dev_ptr is not None:
            capi.check(L.dsb_batch_set_params_device(self._b, ctypes.c_void_p(params_dev_ptr), pr.nbatch, pr.nparams,
                                                     ctypes.c_void_p(stream or 0)))
        t_eval = np.ascontiguousarray(t_eval, dtype=np.float64)
        capi.check(L.dsb_batch_solve_dense(self._b, self.method, _ptr(t_eval), len(t_eval),
                                           ctypes.c_void_p(ys_dev_ptr), ctypes.c_void_p(stream or 0)))
        self._stats = self._status = None

    def set_execution(self, mode):
        """'auto' | 'lane' (one thread per instance, n <= 16) | 'block' (one thread block per instance) |
        'band' (one thread per instance, banded models of n > 16, state in global memory) |
        'warp' (one warp per instance, banded models of n > 16, BDF, state in shared memory)."""
        capi.check(capi.lib().dsb_batch_set_execution(self._b, {"auto": 0, "lane": 1, "block": 2, "band": 3, "warp": 4}[mode]))
        return self

    def set_params(self):
        pr = self.problem
        if pr.nparams:
            capi.check(capi.lib().dsb_batch_set_params_host(self._b, _ptr(pr.p), pr.nbatch, pr.nparams))

    def statistics_array(self):
        if self._stats is None:
            s = np.empty((self.problem.nbatch, capi.DSB_NSTATS), dtype=np.int64)
            capi.check(capi.lib().dsb_batch_get_stats(self._b, _ptr(s)))
            self._stats = s
        return self._stats

    def get_statistics(self, b=0):
        s = self.statistics_array()[b]
        return {name: int(s[i]) for i, name in enumerate(capi.STAT_NAMES)}

    def status(self):
        if self._status is None:
            s = np.empty(self.problem.nbatch, dtype=np.int32)
            capi.check(capi.lib().dsb_batch_get_status(self._b, _ptr(s)))
            self._status = s
        return self._status

    def final_state(self):
        B = self.problem.nbatch
        t, h, o = np.empty(B), np.empty(B), np.empty(B, dtype=np.int32)
        capi.check(capi.lib().dsb_batch_get_final_state(self._b, _ptr(t), _ptr(h), _ptr(o)))
        return t, h, o

    def root_info(self):
        """(root_idx[B], ncols[B]): which root function stopped each instance (-1: none) and how many solve_dense
        columns it wrote; the state at the root is column ncols - 1, its time is final_state()[0]."""
        B = self.problem.nbatch
        r, c = np.empty(B, dtype=np.int32), np.empty(B, dtype=np.int32)
        capi.check(capi.lib().dsb_batch_get_root_info(self._b, _ptr(r), _ptr(c)))
        return r, c

    def sum_statistic(self, name):
        tot = ctypes.c_int64()
        capi.check(capi.lib().dsb_batch_sum_stat(self._b, capi.STAT_NAMES.index(name), ctypes.byref(tot)))
        return tot.value

    def last_kernel_ms(self):
        ms = ctypes.c_float()
        capi.check(capi.lib().dsb_batch_last_kernel_ms(self._b, ctypes.byref(ms)))
        return ms.value

    def last_integrator_ms(self):
        ms = ctypes.c_float()
        capi.check(capi.lib().dsb_batch_last_integrator_ms(self._b, ctypes.byref(ms)))
        return ms.value

    def last_launch_count(self):
        n = ctypes.c_int32()
        capi.check(capi.lib().dsb_batch_last_launch_count(self._b, ctypes.byref(n)))
        return n.value
