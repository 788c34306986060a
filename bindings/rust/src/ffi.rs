//! `extern "C"` declarations of include/diffsol_b200.h, one for one (checked by tests/test_rust_shim_consistency.py).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
pub struct dsb_problem {
    _private: [u8; 0],
}
#[repr(C)]
pub struct dsb_batch {
    _private: [u8; 0],
}

/// OdeSolverOptions + InitialConditionSolverOptions + the Bdf/Sdirk step-size clamps (same names, same defaults)
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct dsb_options {
    pub max_nonlinear_solver_iterations: i32,
    pub max_error_test_failures: i32,
    pub max_nonlinear_solver_failures: i32,
    pub update_jacobian_after_steps: i32,
    pub update_rhs_jacobian_after_steps: i32,
    pub ic_max_linesearch_iterations: i32,
    pub ic_max_newton_iterations: i32,
    pub ic_max_linear_solver_setups: i32,
    pub ic_use_linesearch: i32,
    pub reserved0: i32,
    pub nonlinear_solver_tolerance: f64,
    pub min_timestep: f64,
    pub max_timestep_growth: f64,
    pub min_timestep_growth: f64,
    pub max_timestep_shrink: f64,
    pub min_timestep_shrink: f64,
    pub threshold_to_update_jacobian: f64,
    pub threshold_to_update_rhs_jacobian: f64,
    pub pi_control_proportional: f64,
    pub pi_control_integral: f64,
    pub ic_step_reduction_factor: f64,
    pub ic_armijo_constant: f64,
}

pub const DSB_OK: c_int = 0;
pub const DSB_ERR: c_int = -1;
pub const DSB_BAD_ARG: c_int = -2;
pub const DSB_METHOD_BDF: i32 = 0;
pub const DSB_METHOD_TR_BDF2: i32 = 1;
pub const DSB_METHOD_ESDIRK34: i32 = 2;
pub const DSB_NSTATS: usize = 16;
pub const DSB_MODEL_SOURCE_FUNCTOR: i32 = 0;
pub const DSB_MODEL_SOURCE_DIFFSL: i32 = 1;

extern "C" {
    pub fn dsb_options_default(opt: *mut dsb_options);
    pub fn dsb_last_error() -> *const c_char;
    pub fn dsb_version() -> *const c_char;
    pub fn dsb_device_count(count: *mut c_int) -> c_int;

    pub fn dsb_problem_new(model: c_int, out: *mut *mut dsb_problem) -> c_int;
    pub fn dsb_problem_free(p: *mut dsb_problem) -> c_int;
    pub fn dsb_problem_dims(p: *const dsb_problem, nstates: *mut i32, nparams: *mut i32, has_mass: *mut i32) -> c_int;
    pub fn dsb_problem_nout(p: *const dsb_problem, nout: *mut i32) -> c_int;
    pub fn dsb_problem_set_rtol(p: *mut dsb_problem, rtol: f64) -> c_int;
    pub fn dsb_problem_set_atol(p: *mut dsb_problem, atol: *const f64, n: i32) -> c_int;
    pub fn dsb_problem_set_t0(p: *mut dsb_problem, t0: f64) -> c_int;
    pub fn dsb_problem_set_h0(p: *mut dsb_problem, h0: f64) -> c_int;
    pub fn dsb_problem_set_use_coloring(p: *mut dsb_problem, use_coloring: i32) -> c_int;
    pub fn dsb_problem_set_options(p: *mut dsb_problem, opt: *const dsb_options) -> c_int;
    pub fn dsb_problem_get_options(p: *const dsb_problem, opt: *mut dsb_options) -> c_int;
    pub fn dsb_problem_set_sensitivities(p: *mut dsb_problem, enable: i32, sens_rtol: f64, sens_atol: *const f64, natol: i32) -> c_int;

    pub fn dsb_model_library_build(source_path: *const c_char, kind: i32, struct_name: *const c_char, csrc_dir: *const c_char,
                                   out_library_path: *const c_char) -> c_int;
    pub fn dsb_model_library_load(library_path: *const c_char, model_out: *mut i32) -> c_int;

    pub fn dsb_batch_new(p: *const dsb_problem, nbatch: i64, device: i32, out: *mut *mut dsb_batch) -> c_int;
    pub fn dsb_batch_free(b: *mut dsb_batch) -> c_int;
    pub fn dsb_batch_size(b: *const dsb_batch) -> i64;
    pub fn dsb_batch_set_execution(b: *mut dsb_batch, mode: i32) -> c_int;
    pub fn dsb_batch_set_params_host(b: *mut dsb_batch, params: *const f64, nbatch: i64, nparams: i32) -> c_int;
    pub fn dsb_batch_set_params_device(b: *mut dsb_batch, params_dev: *const f64, nbatch: i64, nparams: i32, stream: *mut c_void) -> c_int;

    pub fn dsb_batch_solve_dense(b: *mut dsb_batch, method: i32, t_eval: *const f64, nt: i32, ys_dev: *mut f64, stream: *mut c_void) -> c_int;
    pub fn dsb_batch_step_and_interpolate(b: *mut dsb_batch, method: i32, t_points: *const f64, npts: i32, ys_dev: *mut f64,
                                          stream: *mut c_void) -> c_int;
    pub fn dsb_batch_solve_dense_sensitivities(b: *mut dsb_batch, method: i32, t_eval: *const f64, nt: i32, ys_dev: *mut f64,
                                               sens_dev: *mut f64, stream: *mut c_void) -> c_int;
    pub fn dsb_batch_step_and_interpolate_sensitivities(b: *mut dsb_batch, method: i32, t_points: *const f64, npts: i32, ys_dev: *mut f64,
                                                        sens_dev: *mut f64, stream: *mut c_void) -> c_int;
    pub fn dsb_batch_solve_count(b: *mut dsb_batch, method: i32, final_time: f64, total_columns: *mut i64) -> c_int;
    pub fn dsb_batch_solve_offsets(b: *mut dsb_batch, offsets_host: *mut i64) -> c_int;
    pub fn dsb_batch_solve_write(b: *mut dsb_batch, method: i32, final_time: f64, ts_dev: *mut f64, ys_dev: *mut f64, stream: *mut c_void) -> c_int;
    pub fn dsb_batch_solve_write_host(b: *mut dsb_batch, method: i32, final_time: f64, ts_host: *mut f64, ys_host: *mut f64) -> c_int;
    pub fn dsb_batch_solve_dense_sensitivities_host(b: *mut dsb_batch, method: i32, params_host: *const f64, nparams: i32, t_eval: *const f64,
                                                    nt: i32, ys_host: *mut f64, sens_host: *mut f64, stats_host: *mut i64,
                                                    status_host: *mut i32) -> c_int;
    pub fn dsb_batch_step_and_interpolate_sensitivities_host(b: *mut dsb_batch, method: i32, params_host: *const f64, nparams: i32,
                                                             t_points: *const f64, npts: i32, ys_host: *mut f64, sens_host: *mut f64,
                                                             stats_host: *mut i64, status_host: *mut i32) -> c_int;
    pub fn dsb_batch_solve_dense_host(b: *mut dsb_batch, method: i32, params_host: *const f64, nparams: i32, t_eval: *const f64, nt: i32,
                                      ys_host: *mut f64, stats_host: *mut i64, status_host: *mut i32) -> c_int;
    pub fn dsb_batch_step_and_interpolate_host(b: *mut dsb_batch, method: i32, params_host: *const f64, nparams: i32, t_points: *const f64,
                                               npts: i32, ys_host: *mut f64, stats_host: *mut i64, status_host: *mut i32) -> c_int;

    pub fn dsb_batch_get_stats(b: *mut dsb_batch, stats_host: *mut i64) -> c_int;
    pub fn dsb_batch_get_status(b: *mut dsb_batch, status_host: *mut i32) -> c_int;
    pub fn dsb_batch_get_stats_device(b: *mut dsb_batch, stats_dev: *mut i64, stream: *mut c_void) -> c_int;
    pub fn dsb_batch_get_final_state(b: *mut dsb_batch, t_host: *mut f64, h_host: *mut f64, order_host: *mut i32) -> c_int;
    pub fn dsb_batch_get_root_info(b: *mut dsb_batch, root_idx_host: *mut i32, ncols_host: *mut i32) -> c_int;
    pub fn dsb_batch_device_views(b: *mut dsb_batch, stats_dev: *mut *const i32, status_dev: *mut *const i32) -> c_int;
    pub fn dsb_batch_sum_stat(b: *mut dsb_batch, stat: i32, total: *mut i64) -> c_int;
    pub fn dsb_batch_last_kernel_ms(b: *mut dsb_batch, ms: *mut f32) -> c_int;
    pub fn dsb_batch_last_integrator_ms(b: *mut dsb_batch, ms: *mut f32) -> c_int;
    pub fn dsb_batch_last_launch_count(b: *mut dsb_batch, launches: *mut i32) -> c_int;
    pub fn dsb_batch_debug_words(b: *mut dsb_batch, words_host: *mut u64) -> c_int;

    pub fn dsb_lu_factor_batched(a_dev: *mut f64, n: i32, nbatch: i64, piv_dev: *mut i32, info_dev: *mut i32, stream: *mut c_void) -> c_int;
    pub fn dsb_lu_solve_batched(lu_dev: *const f64, piv_dev: *const i32, b_dev: *mut f64, n: i32, nbatch: i64, info_dev: *mut i32,
                                stream: *mut c_void) -> c_int;
    pub fn dsb_lu_factor_instance_major(a_dev: *mut f64, n: i32, nbatch: i64, piv_dev: *mut i32, info_dev: *mut i32, stream: *mut c_void) -> c_int;
    pub fn dsb_lu_solve_instance_major(lu_dev: *const f64, piv_dev: *const i32, b_dev: *mut f64, n: i32, nbatch: i64, info_dev: *mut i32,
                                       stream: *mut c_void) -> c_int;
}
