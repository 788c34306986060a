//! diffsol-b200: the batched implicit step loop of diffsol's `Bdf` / `Sdirk` on NVIDIA B200, behind the call shapes of
//! diffsol's own API (`OdeBuilder` -> `problem.bdf()` -> `solve_dense` / `solve` / `solve_dense_sensitivities`), over the C ABI
//! of `include/diffsol_b200.h`.
//!
//! One `Problem` describes the equations and tolerances for a whole batch; a `Solver` owns the device state of `nbatch`
//! independent instances, each with its own parameters, adaptive step size, order and counters.  Results come back
//! instance-major: instance `b`'s block is the column-major `nstates x nt` matrix diffsol's `solve_dense` returns.
//!
//! This crate is source only in this repository (no Rust toolchain in the development image); its `extern "C"` block is
//! checked against the header by `tests/test_rust_shim_consistency.py`.
pub mod ffi;

use std::ffi::{CStr, CString};
use std::ptr;

/// Mirrors diffsol's `DiffsolError`: the API-level failure (message from `dsb_last_error`).  A failing *instance* is not an
/// error: see `Solver::status`.
#[derive(Debug, Clone)]
pub struct Error {
    pub code: i32,
    pub message: String,
}
impl std::fmt::Display for Error {
    fn fmt(&self, f: &mut std::fmt::Formatter<'_>) -> std::fmt::Result {
        write!(f, "diffsol_b200 error {}: {}", self.code, self.message)
    }
}
impl std::error::Error for Error {}

fn check(rc: i32) -> Result<(), Error> {
    if rc == ffi::DSB_OK {
        return Ok(());
    }
    let message = unsafe { CStr::from_ptr(ffi::dsb_last_error()) }.to_string_lossy().into_owned();
    Err(Error { code: rc, message })
}

/// `OdeSolverMethod` choice (`problem.bdf()`, `problem.tr_bdf2()`, `problem.esdirk34()`)
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub enum Method {
    Bdf,
    TrBdf2,
    Esdirk34,
}
impl Method {
    fn id(self) -> i32 {
        match self {
            Method::Bdf => ffi::DSB_METHOD_BDF,
            Method::TrBdf2 => ffi::DSB_METHOD_TR_BDF2,
            Method::Esdirk34 => ffi::DSB_METHOD_ESDIRK34,
        }
    }
}

/// How the equations are given: one of the library's built-in equation sets (`enum dsb_model_id`), or source text compiled at
/// run time into the library's kernels (a closure-style functor, or the DiffSL symbol table as C).
pub enum Equations<'a> {
    BuiltIn(i32),
    FunctorSource { text: &'a str, struct_name: &'a str },
    DiffslSource { text: &'a str },
}

/// `OdeBuilder`: same defaults as diffsol's (rtol 1e-6, atol [1e-6], t0 0, h0 1).
pub struct OdeBuilder<'a> {
    equations: Option<Equations<'a>>,
    rtol: f64,
    atol: Vec<f64>,
    t0: f64,
    h0: f64,
    use_coloring: bool,
    sens: Option<(f64, Vec<f64>)>,
    options: Option<ffi::dsb_options>,
    csrc_dir: String,
    work_dir: String,
}
impl<'a> Default for OdeBuilder<'a> {
    fn default() -> Self {
        Self::new()
    }
}
impl<'a> OdeBuilder<'a> {
    pub fn new() -> Self {
        OdeBuilder {
            equations: None,
            rtol: 1e-6,
            atol: vec![1e-6],
            t0: 0.0,
            h0: 1.0,
            use_coloring: false,
            sens: None,
            options: None,
            csrc_dir: concat!(env!("CARGO_MANIFEST_DIR"), "/../../diffsol_b200/csrc").to_string(),
            work_dir: std::env::temp_dir().to_string_lossy().into_owned(),
        }
    }
    pub fn equations(mut self, e: Equations<'a>) -> Self {
        self.equations = Some(e);
        self
    }
    pub fn rtol(mut self, v: f64) -> Self {
        self.rtol = v;
        self
    }
    pub fn atol(mut self, v: &[f64]) -> Self {
        self.atol = v.to_vec();
        self
    }
    pub fn t0(mut self, v: f64) -> Self {
        self.t0 = v;
        self
    }
    pub fn h0(mut self, v: f64) -> Self {
        self.h0 = v;
        self
    }
    pub fn use_coloring(mut self, v: bool) -> Self {
        self.use_coloring = v;
        self
    }
    /// `sens_rtol` + `sens_atol`; an empty `sens_atol` keeps the sensitivities out of the error test
    pub fn sensitivities(mut self, sens_rtol: f64, sens_atol: &[f64]) -> Self {
        self.sens = Some((sens_rtol, sens_atol.to_vec()));
        self
    }
    pub fn ode_options(mut self, o: ffi::dsb_options) -> Self {
        self.options = Some(o);
        self
    }
    pub fn build(self) -> Result<Problem, Error> {
        let model = match self.equations {
            None => return Err(Error { code: ffi::DSB_BAD_ARG, message: "OdeBuilder::build: no equations given".into() }),
            Some(Equations::BuiltIn(id)) => id,
            Some(Equations::FunctorSource { text, struct_name }) => {
                Self::compile(&self.csrc_dir, &self.work_dir, text, ffi::DSB_MODEL_SOURCE_FUNCTOR, Some(struct_name))?
            }
            Some(Equations::DiffslSource { text }) => {
                Self::compile(&self.csrc_dir, &self.work_dir, text, ffi::DSB_MODEL_SOURCE_DIFFSL, None)?
            }
        };
        let mut handle = ptr::null_mut();
        check(unsafe { ffi::dsb_problem_new(model, &mut handle) })?;
        let mut problem = Problem { handle, nstates: 0, nparams: 0, nout: 0, has_mass: false, has_sens: self.sens.is_some() };
        let (mut n, mut np, mut hm, mut nout) = (0i32, 0i32, 0i32, 0i32);
        check(unsafe { ffi::dsb_problem_dims(handle, &mut n, &mut np, &mut hm) })?;
        check(unsafe { ffi::dsb_problem_nout(handle, &mut nout) })?;
        problem.nstates = n as usize;
        problem.nparams = np as usize;
        problem.nout = nout as usize;
        problem.has_mass = hm != 0;
        check(unsafe { ffi::dsb_problem_set_rtol(handle, self.rtol) })?;
        check(unsafe { ffi::dsb_problem_set_atol(handle, self.atol.as_ptr(), self.atol.len() as i32) })?;
        check(unsafe { ffi::dsb_problem_set_t0(handle, self.t0) })?;
        check(unsafe { ffi::dsb_problem_set_h0(handle, self.h0) })?;
        check(unsafe { ffi::dsb_problem_set_use_coloring(handle, self.use_coloring as i32) })?;
        if let Some(o) = self.options {
            check(unsafe { ffi::dsb_problem_set_options(handle, &o) })?;
        }
        if let Some((rtol, atol)) = &self.sens {
            let p = if atol.is_empty() { ptr::null() } else { atol.as_ptr() };
            check(unsafe { ffi::dsb_problem_set_sensitivities(handle, 1, *rtol, p, atol.len() as i32) })?;
        }
        Ok(problem)
    }
    fn compile(csrc_dir: &str, work_dir: &str, text: &str, kind: i32, struct_name: Option<&str>) -> Result<i32, Error> {
        use std::collections::hash_map::DefaultHasher;
        use std::hash::{Hash, Hasher};
        let mut h = DefaultHasher::new();
        text.hash(&mut h);
        kind.hash(&mut h);
        let stem = format!("{}/dsb_model_{:016x}", work_dir, h.finish());
        let (src, so) = (format!("{}.h", stem), format!("{}.so", stem));
        std::fs::write(&src, text).map_err(|e| Error { code: ffi::DSB_ERR, message: e.to_string() })?;
        let c = |s: &str| CString::new(s).unwrap();
        let (src_c, so_c, dir_c) = (c(&src), c(&so), c(csrc_dir));
        let name_c = struct_name.map(c);
        let name_p = name_c.as_ref().map_or(ptr::null(), |n| n.as_ptr());
        check(unsafe { ffi::dsb_model_library_build(src_c.as_ptr(), kind, name_p, dir_c.as_ptr(), so_c.as_ptr()) })?;
        let mut id = 0i32;
        check(unsafe { ffi::dsb_model_library_load(so_c.as_ptr(), &mut id) })?;
        Ok(id)
    }
}

/// `OdeSolverProblem` for a whole batch
pub struct Problem {
    handle: *mut ffi::dsb_problem,
    pub nstates: usize,
    pub nparams: usize,
    pub nout: usize,
    pub has_mass: bool,
    pub has_sens: bool,
}
impl Drop for Problem {
    fn drop(&mut self) {
        unsafe { ffi::dsb_problem_free(self.handle) };
    }
}
impl Problem {
    /// `problem.bdf::<LS>()` etc. for `nbatch` instances on GPU `device`; `params` is instance-major `nbatch x nparams`
    pub fn solver(&self, method: Method, params: &[f64], device: i32) -> Result<Solver<'_>, Error> {
        let np = self.nparams.max(1);
        if self.nparams > 0 && params.len() % np != 0 {
            return Err(Error { code: ffi::DSB_BAD_ARG, message: "params must hold nbatch x nparams values".into() });
        }
        let nbatch = if self.nparams > 0 { params.len() / np } else { params.len().max(1) };
        let mut batch = ptr::null_mut();
        check(unsafe { ffi::dsb_batch_new(self.handle, nbatch as i64, device, &mut batch) })?;
        Ok(Solver { problem: self, batch, method, nbatch, params: if self.nparams > 0 { params.to_vec() } else { Vec::new() } })
    }
    pub fn bdf(&self, params: &[f64]) -> Result<Solver<'_>, Error> {
        self.solver(Method::Bdf, params, 0)
    }
    pub fn tr_bdf2(&self, params: &[f64]) -> Result<Solver<'_>, Error> {
        self.solver(Method::TrBdf2, params, 0)
    }
    pub fn esdirk34(&self, params: &[f64]) -> Result<Solver<'_>, Error> {
        self.solver(Method::Esdirk34, params, 0)
    }
}

/// The 13 counters diffsol reports per solver (`OdeSolverStatistics` + the rhs operator's `OpStatistics`), per instance
#[derive(Clone, Copy, Debug, Default, PartialEq, Eq)]
pub struct Statistics {
    pub number_of_linear_solver_setups: i64,
    pub number_of_steps: i64,
    pub number_of_error_test_failures: i64,
    pub number_of_nonlinear_solver_iterations: i64,
    pub number_of_nonlinear_solver_fails: i64,
    pub rhs_number_of_calls: i64,
    pub rhs_number_of_jac_muls: i64,
    pub rhs_number_of_matrix_evals: i64,
}

pub struct Solver<'p> {
    problem: &'p Problem,
    batch: *mut ffi::dsb_batch,
    method: Method,
    pub nbatch: usize,
    params: Vec<f64>,
}
impl Drop for Solver<'_> {
    fn drop(&mut self) {
        unsafe { ffi::dsb_batch_free(self.batch) };
    }
}
impl Solver<'_> {
    fn params_ptr(&self) -> *const f64 {
        if self.params.is_empty() { ptr::null() } else { self.params.as_ptr() }
    }
    /// `solve_dense(t_eval)`: instance-major `[nbatch][nt][nout]`
    pub fn solve_dense(&mut self, t_eval: &[f64]) -> Result<Vec<f64>, Error> {
        let mut ys = vec![f64::NAN; self.nbatch * t_eval.len() * self.problem.nout];
        check(unsafe {
            ffi::dsb_batch_solve_dense_host(self.batch, self.method.id(), self.params_ptr(), self.problem.nparams as i32, t_eval.as_ptr(),
                                            t_eval.len() as i32, ys.as_mut_ptr(), ptr::null_mut(), ptr::null_mut())
        })?;
        Ok(ys)
    }
    /// the `while t < t_k { step() }; interpolate(t_k)` loop of diffsol's tests
    pub fn step_and_interpolate(&mut self, t_points: &[f64]) -> Result<Vec<f64>, Error> {
        let mut ys = vec![f64::NAN; self.nbatch * t_points.len() * self.problem.nout];
        check(unsafe {
            ffi::dsb_batch_step_and_interpolate_host(self.batch, self.method.id(), self.params_ptr(), self.problem.nparams as i32,
                                                     t_points.as_ptr(), t_points.len() as i32, ys.as_mut_ptr(), ptr::null_mut(), ptr::null_mut())
        })?;
        Ok(ys)
    }
    /// `solve_dense_sensitivities(t_eval)` -> (ys `[nbatch][nt][nstates]`, sens `[nbatch][nt][nparams][nstates]`)
    pub fn solve_dense_sensitivities(&mut self, t_eval: &[f64]) -> Result<(Vec<f64>, Vec<f64>), Error> {
        let (n, np, nt) = (self.problem.nstates, self.problem.nparams, t_eval.len());
        let mut ys = vec![f64::NAN; self.nbatch * nt * n];
        let mut sens = vec![f64::NAN; self.nbatch * nt * np * n];
        check(unsafe {
            ffi::dsb_batch_solve_dense_sensitivities_host(self.batch, self.method.id(), self.params_ptr(), np as i32, t_eval.as_ptr(), nt as i32,
                                                          ys.as_mut_ptr(), sens.as_mut_ptr(), ptr::null_mut(), ptr::null_mut())
        })?;
        Ok((ys, sens))
    }
    /// `solve(final_time)`: every internal step.  -> (ys `[total][nout]`, ts `[total]`, offsets `[nbatch + 1]`)
    pub fn solve(&mut self, final_time: f64) -> Result<(Vec<f64>, Vec<f64>, Vec<i64>), Error> {
        if !self.params.is_empty() {
            check(unsafe { ffi::dsb_batch_set_params_host(self.batch, self.params.as_ptr(), self.nbatch as i64, self.problem.nparams as i32) })?;
        }
        let mut total = 0i64;
        check(unsafe { ffi::dsb_batch_solve_count(self.batch, self.method.id(), final_time, &mut total) })?;
        let mut offsets = vec![0i64; self.nbatch + 1];
        check(unsafe { ffi::dsb_batch_solve_offsets(self.batch, offsets.as_mut_ptr()) })?;
        let mut ts = vec![f64::NAN; total as usize];
        let mut ys = vec![f64::NAN; total as usize * self.problem.nout];
        check(unsafe { ffi::dsb_batch_solve_write_host(self.batch, self.method.id(), final_time, ts.as_mut_ptr(), ys.as_mut_ptr()) })?;
        Ok((ys, ts, offsets))
    }
    /// per-instance `OdeSolverError` of the last solve (0 = Ok; `enum dsb_status`)
    pub fn status(&self) -> Result<Vec<i32>, Error> {
        let mut s = vec![0i32; self.nbatch];
        check(unsafe { ffi::dsb_batch_get_status(self.batch, s.as_mut_ptr()) })?;
        Ok(s)
    }
    /// `get_statistics()` of instance `b`
    pub fn statistics(&self, b: usize) -> Result<Statistics, Error> {
        let mut raw = vec![0i64; self.nbatch * ffi::DSB_NSTATS];
        check(unsafe { ffi::dsb_batch_get_stats(self.batch, raw.as_mut_ptr()) })?;
        let r = &raw[b * ffi::DSB_NSTATS..(b + 1) * ffi::DSB_NSTATS];
        Ok(Statistics {
            number_of_linear_solver_setups: r[0],
            number_of_steps: r[6],
            number_of_error_test_failures: r[7],
            number_of_nonlinear_solver_iterations: r[8],
            number_of_nonlinear_solver_fails: r[9],
            rhs_number_of_calls: r[10],
            rhs_number_of_jac_muls: r[11],
            rhs_number_of_matrix_evals: r[12],
        })
    }
    /// `OdeSolverStopReason::RootFound` per instance: (index of the root function or -1, columns written)
    pub fn root_info(&self) -> Result<(Vec<i32>, Vec<i32>), Error> {
        let (mut idx, mut nc) = (vec![0i32; self.nbatch], vec![0i32; self.nbatch]);
        check(unsafe { ffi::dsb_batch_get_root_info(self.batch, idx.as_mut_ptr(), nc.as_mut_ptr()) })?;
        Ok((idx, nc))
    }
}
