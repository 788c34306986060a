// diffsol_b200.hpp -- C++17 host-side mirror of the reference's solver interface for the batched implicit path,
// header-only, over the C ABI of diffsol_b200.h (nothing but plain pointers and sizes crosses the library boundary).
//
// The reference is Rust; there is no Rust toolchain in this image, so the host side above the C ABI is written in
// C++ (and in Python, diffsol_b200/ode.py) with the reference's names, argument meaning and error behaviour:
//
//   reference (crates/diffsol/src)                              here
//   ------------------------------------------------------------------------------------------------------------
//   OdeBuilder::<M>::new().rtol().atol().t0().h0().p()           diffsol_b200::OdeBuilder (ode_solver/builder.rs)
//     .rhs_implicit(f, jac).init(..).root(..).reset(..).out(..)  .rhs_implicit("<built-in equation set>"): the device
//                                                                 functor carries rhs / jac_mul / mass / init / root /
//                                                                 reset / out (csrc/dsb_models.h)
//     .use_coloring(b).build()                                   .use_coloring(b).build()  -> OdeSolverProblem
//   problem.bdf::<LS>() / tr_bdf2::<LS>() / esdirk34::<LS>()     problem.bdf() / tr_bdf2() / esdirk34() -> BatchedSolver
//   OdeSolverMethod::solve_dense(&t_eval) -> M                   solver.solve_dense(t_eval) -> DenseBlock per instance
//   OdeSolverMethod::state().t / .h, get_statistics()            solver.final_state(), solver.get_statistics(b)
//   OdeSolverStopReason::RootFound(t, idx)                       solver.root_info()
//   DiffsolError / OdeSolverError                                diffsol_b200::DiffsolError (call failed),
//                                                                 solver.status()[b] (per instance: a failed instance
//                                                                 does not abort the batch)
//
// One OdeSolverProblem holds a BATCH of independent instances of one equation set: p(params) takes nbatch x nparams
// values, instance-major, as the reference lays out batched parameters (test_models/exponential_decay.rs:297-304).
// There is no CPU fallback: every solve runs the CUDA kernels and fails with DiffsolError when no device is present.
#pragma once
#include <cstdint>
#include <cmath>
#include <limits>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "diffsol_b200.h"

namespace diffsol_b200 {

// DiffsolError (crates/diffsol/src/error.rs): what `?` would propagate in the reference
class DiffsolError : public std::runtime_error {
public:
    DiffsolError(int code, const std::string& what) : std::runtime_error(what), code_(code) {}
    int code() const { return code_; }
private:
    int code_;
};

namespace detail {
inline void check(int rc, const char* where) {
    if (rc != DSB_OK) {
        const char* msg = dsb_last_error();
        throw DiffsolError(rc, std::string(where) + ": " + (msg && *msg ? msg : "error"));
    }
}
}  // namespace detail

// OdeSolverStatistics (ode_solver/mod.rs:27-49) + the rhs OpStatistics (op/mod.rs:108-145), one per instance
struct OdeSolverStatistics {
    int64_t number_of_linear_solver_setups = 0;
    int64_t number_of_linear_solver_setups_from_checkpoint = 0;
    int64_t number_of_linear_solver_setups_from_first_convergence_fail = 0;
    int64_t number_of_linear_solver_setups_from_second_convergence_fail = 0;
    int64_t number_of_linear_solver_setups_from_error_test_fail = 0;
    int64_t number_of_linear_solver_setups_from_step_success = 0;
    int64_t number_of_steps = 0;
    int64_t number_of_error_test_failures = 0;
    int64_t number_of_nonlinear_solver_iterations = 0;
    int64_t number_of_nonlinear_solver_fails = 0;
    int64_t rhs_number_of_calls = 0;
    int64_t rhs_number_of_jac_muls = 0;
    int64_t rhs_number_of_matrix_evals = 0;
};

// OdeSolverStopReason (ode_solver/method.rs): how an instance's solve_dense ended
struct StopInfo {
    bool root_found = false;      // RootFound(t, index): the solve ended at a root
    int32_t root_index = -1;
    double t = 0.0;               // final time of the instance (the root time when root_found)
    int32_t ncols = 0;            // columns of the result that were written (the reference resizes its matrix)
};

// The matrix solve_dense returns, for every instance: rows x nt column-major blocks, instance-major.
class DenseBlocks {
public:
    DenseBlocks() = default;
    DenseBlocks(int64_t nbatch, int32_t rows, int32_t nt)
        : nbatch_(nbatch), rows_(rows), nt_(nt),
          data_((size_t)nbatch * rows * nt, std::numeric_limits<double>::quiet_NaN()) {}
    int64_t nbatch() const { return nbatch_; }
    int32_t nrows() const { return rows_; }
    int32_t ncols() const { return nt_; }
    // element (row i, column j) of instance b: ys[(i, j)] of the reference's matrix
    double operator()(int64_t b, int32_t i, int32_t j) const { return data_[((size_t)b * nt_ + j) * rows_ + i]; }
    const double* instance(int64_t b) const { return data_.data() + (size_t)b * nt_ * rows_; }
    double* data() { return data_.data(); }
    const double* data() const { return data_.data(); }
private:
    int64_t nbatch_ = 0;
    int32_t rows_ = 0, nt_ = 0;
    std::vector<double> data_;
};

class BatchedSolver;

// OdeSolverProblem (ode_solver/problem.rs): equations + tolerances + options + the batch's parameters
class OdeSolverProblem {
public:
    OdeSolverProblem(const OdeSolverProblem&) = delete;
    OdeSolverProblem& operator=(const OdeSolverProblem&) = delete;
    OdeSolverProblem(OdeSolverProblem&& o) noexcept { *this = std::move(o); }
    OdeSolverProblem& operator=(OdeSolverProblem&& o) noexcept {
        if (this != &o) { release(); handle_ = o.handle_; o.handle_ = nullptr; nstates_ = o.nstates_; nparams_ = o.nparams_;
                          nout_ = o.nout_; has_mass_ = o.has_mass_; nbatch_ = o.nbatch_; device_ = o.device_;
                          params_ = std::move(o.params_); }
        return *this;
    }
    ~OdeSolverProblem() { release(); }

    int32_t nstates() const { return nstates_; }
    int32_t nparams() const { return nparams_; }
    int32_t nout() const { return nout_; }            // rows of solve_dense: the output function's, else nstates
    bool has_mass() const { return has_mass_; }
    int64_t nbatch() const { return nbatch_; }

    // problem.bdf::<LS>() (ode_solver/problem.rs:649-655), tr_bdf2 / esdirk34 (:320-328): the linear solver is the
    // library's batched dense / band LU (the NalgebraLU arithmetic restated on the device)
    inline BatchedSolver bdf() const;
    inline BatchedSolver tr_bdf2() const;
    inline BatchedSolver esdirk34() const;

private:
    friend class OdeBuilder;
    friend class BatchedSolver;
    OdeSolverProblem() = default;
    void release() { if (handle_) { dsb_problem_free(handle_); handle_ = nullptr; } }
    dsb_problem* handle_ = nullptr;
    int32_t nstates_ = 0, nparams_ = 0, nout_ = 0;
    bool has_mass_ = false;
    int64_t nbatch_ = 1;
    int32_t device_ = 0;
    std::vector<double> params_;
};

// OdeBuilder (ode_solver/builder.rs): same defaults -- rtol = 1e-6, atol = [1e-6], t0 = 0, h0 = 1, no colouring
class OdeBuilder {
public:
    OdeBuilder() { dsb_options_default(&opt_); }
    static const std::map<std::string, int>& models() {
        static const std::map<std::string, int> m = {
            {"exp_decay", DSB_EXP_DECAY}, {"exp_decay_algebraic", DSB_EXP_DECAY_ALGEBRAIC},
            {"robertson_dae", DSB_ROBERTSON_DAE}, {"robertson_ode", DSB_ROBERTSON_ODE},
            {"robertson_ode_g3", DSB_ROBERTSON_ODE_G3}, {"dydt_y2", DSB_DYDT_Y2}, {"gaussian_decay", DSB_GAUSSIAN_DECAY},
            {"van_der_pol", DSB_VAN_DER_POL}, {"van_der_pol_scaled", DSB_VAN_DER_POL_SCALED},
            {"heat1d_dae_256", DSB_HEAT1D_DAE_256}, {"heat1d_dae_32", DSB_HEAT1D_DAE_32}, {"spm", DSB_SPM},
            {"spm99", DSB_SPM99}, {"exp_decay_root", DSB_EXP_DECAY_ROOT}, {"spm_stop", DSB_SPM_STOP},
            {"spm99_stop", DSB_SPM99_STOP}, {"heat1d_dae_32_bc", DSB_HEAT1D_DAE_32_BC},
            {"exp_decay_reset", DSB_EXP_DECAY_RESET}, {"heat2d_10", DSB_HEAT2D_10}, {"ball_bounce", DSB_BALL_BOUNCE},
            {"exp_decay_two_roots", DSB_EXP_DECAY_TWO_ROOTS}, {"spm_cycle", DSB_SPM_CYCLE},
            {"exp_decay_algebraic_reset", DSB_EXP_DECAY_ALGEBRAIC_RESET}};
        return m;
    }
    // .rhs_implicit(f, jac) + .init / .mass / .root / .reset / .out of the reference: one built-in device functor
    OdeBuilder& rhs_implicit(const std::string& model) {
        auto it = models().find(model);
        if (it == models().end()) throw DiffsolError(DSB_BAD_ARG, "OdeBuilder::rhs_implicit: unknown equation set '" + model + "'");
        model_ = it->second; return *this;
    }
    OdeBuilder& rhs_implicit(int model_id) { model_ = model_id; return *this; }
    OdeBuilder& rtol(double v) { rtol_ = v; return *this; }
    OdeBuilder& atol(std::vector<double> v) { atol_ = std::move(v); return *this; }
    OdeBuilder& atol(double v) { atol_.assign(1, v); return *this; }
    OdeBuilder& t0(double v) { t0_ = v; return *this; }
    OdeBuilder& h0(double v) { h0_ = v; return *this; }
    OdeBuilder& use_coloring(bool v) { use_coloring_ = v; return *this; }
    // OdeBuilder::sens_rtol / sens_atol (builder.rs:1454-1477): forward sensitivities in the error test;
    // sensitivities(true) alone integrates them outside it (turn_off_sensitivities_error_control)
    OdeBuilder& sens_rtol(double v) { sens_ = true; sens_rtol_ = v; return *this; }
    OdeBuilder& sens_atol(std::vector<double> v) { sens_ = true; sens_atol_ = std::move(v); return *this; }
    OdeBuilder& sensitivities(bool on) { sens_ = on; return *this; }
    // parameters of the whole batch, instance-major: nbatch x nparams values
    OdeBuilder& p(std::vector<double> v) { p_ = std::move(v); return *this; }
    OdeBuilder& device(int32_t d) { device_ = d; return *this; }
    // OdeSolverOptions / InitialConditionSolverOptions (ode_solver/config.rs), all fields of dsb_options
    dsb_options& ode_options() { return opt_; }

    OdeSolverProblem build() const {
        if (model_ < 0) throw DiffsolError(DSB_BAD_ARG, "OdeBuilder::build: no equations (rhs_implicit) given");
        OdeSolverProblem pr;
        detail::check(dsb_problem_new(model_, &pr.handle_), "dsb_problem_new");
        int32_t n = 0, np = 0, hm = 0, nout = 0;
        detail::check(dsb_problem_dims(pr.handle_, &n, &np, &hm), "dsb_problem_dims");
        detail::check(dsb_problem_nout(pr.handle_, &nout), "dsb_problem_nout");
        pr.nstates_ = n; pr.nparams_ = np; pr.has_mass_ = hm != 0; pr.nout_ = nout > 0 ? nout : n;
        if (np > 0) {
            if (p_.empty() || p_.size() % (size_t)np != 0)
                throw DiffsolError(DSB_BAD_ARG, "OdeBuilder::build: p must hold nbatch x " + std::to_string(np) + " values");
            pr.nbatch_ = (int64_t)(p_.size() / (size_t)np);
        } else {
            pr.nbatch_ = p_.empty() ? 1 : (int64_t)p_.size();       // parameter-free equations: p's length is the batch size
        }
        pr.params_ = np > 0 ? p_ : std::vector<double>();
        pr.device_ = device_;
        detail::check(dsb_problem_set_rtol(pr.handle_, rtol_), "dsb_problem_set_rtol");
        detail::check(dsb_problem_set_atol(pr.handle_, atol_.data(), (int32_t)atol_.size()), "dsb_problem_set_atol");
        detail::check(dsb_problem_set_t0(pr.handle_, t0_), "dsb_problem_set_t0");
        detail::check(dsb_problem_set_h0(pr.handle_, h0_), "dsb_problem_set_h0");
        detail::check(dsb_problem_set_use_coloring(pr.handle_, use_coloring_ ? 1 : 0), "dsb_problem_set_use_coloring");
        detail::check(dsb_problem_set_options(pr.handle_, &opt_), "dsb_problem_set_options");
        if (sens_)
            detail::check(dsb_problem_set_sensitivities(pr.handle_, 1, sens_rtol_, sens_atol_.empty() ? nullptr : sens_atol_.data(),
                                                        (int32_t)sens_atol_.size()), "dsb_problem_set_sensitivities");
        return pr;
    }

private:
    int model_ = -1;
    double rtol_ = 1e-6, t0_ = 0.0, h0_ = 1.0;
    std::vector<double> atol_ = {1e-6};
    bool use_coloring_ = false;
    bool sens_ = false;
    double sens_rtol_ = 0.0;
    std::vector<double> sens_atol_;
    std::vector<double> p_;
    int32_t device_ = 0;
    dsb_options opt_;
};

// The solver object problem.bdf() returns: OdeSolverMethod (ode_solver/method.rs) over the whole batch
class BatchedSolver {
public:
    BatchedSolver(const OdeSolverProblem& problem, int32_t method) : problem_(&problem), method_(method) {
        detail::check(dsb_batch_new(problem.handle_, problem.nbatch_, problem.device_, &batch_), "dsb_batch_new");
    }
    BatchedSolver(const BatchedSolver&) = delete;
    BatchedSolver& operator=(const BatchedSolver&) = delete;
    BatchedSolver(BatchedSolver&& o) noexcept : problem_(o.problem_), method_(o.method_), batch_(o.batch_) { o.batch_ = nullptr; }
    ~BatchedSolver() { if (batch_) dsb_batch_free(batch_); }

    // 0 automatic, 1 one thread per instance (state on chip), 2 one block per instance, 3 banded lane kernels, 4 banded warp-per-instance kernel
    BatchedSolver& set_execution(int32_t mode) { detail::check(dsb_batch_set_execution(batch_, mode), "dsb_batch_set_execution"); return *this; }

    // OdeSolverMethod::solve_dense (ode_solver/method.rs:721-848): tstop = t_eval.back(), dense output at every t_eval;
    // host buffers in, host buffers out (the copies are part of the call)
    DenseBlocks solve_dense(const std::vector<double>& t_eval) {
        if (t_eval.empty()) throw DiffsolError(DSB_BAD_ARG, "solve_dense: t_eval is empty");
        DenseBlocks ys(problem_->nbatch_, problem_->nout_, (int32_t)t_eval.size());
        detail::check(dsb_batch_solve_dense_host(batch_, method_, problem_->params_.empty() ? nullptr : problem_->params_.data(),
                                                 problem_->nparams_, t_eval.data(), (int32_t)t_eval.size(), ys.data(), nullptr, nullptr),
                      "dsb_batch_solve_dense_host");
        return ys;
    }
    // solve_dense_sensitivities (ode_solver/sensitivities.rs:114-262) on a problem built with sens_rtol / sens_atol /
    // sensitivities(true), solver = problem.bdf(): the states, and per instance nt x nparams sensitivity vectors
    // (sens block b, time k, parameter q at rows [q * nstates, (q + 1) * nstates) of column k)
    std::pair<DenseBlocks, DenseBlocks> solve_dense_sensitivities(const std::vector<double>& t_eval) {
        if (t_eval.empty()) throw DiffsolError(DSB_BAD_ARG, "solve_dense_sensitivities: t_eval is empty");
        DenseBlocks ys(problem_->nbatch_, problem_->nstates_, (int32_t)t_eval.size());
        DenseBlocks sens(problem_->nbatch_, problem_->nstates_ * problem_->nparams_, (int32_t)t_eval.size());
        detail::check(dsb_batch_solve_dense_sensitivities_host(batch_, method_, problem_->params_.empty() ? nullptr : problem_->params_.data(),
                                                               problem_->nparams_, t_eval.data(), (int32_t)t_eval.size(), ys.data(), sens.data(),
                                                               nullptr, nullptr),
                      "dsb_batch_solve_dense_sensitivities_host");
        return {std::move(ys), std::move(sens)};
    }
    // The loop of the reference's test harness (ode_solver/mod.rs:104-194): step while |t| < |t_point|, interpolate
    DenseBlocks step_and_interpolate(const std::vector<double>& t_points) {
        if (t_points.empty()) throw DiffsolError(DSB_BAD_ARG, "step_and_interpolate: t_points is empty");
        DenseBlocks ys(problem_->nbatch_, problem_->nout_, (int32_t)t_points.size());   // rows = outputs, as solve_dense
        detail::check(dsb_batch_step_and_interpolate_host(batch_, method_, problem_->params_.empty() ? nullptr : problem_->params_.data(),
                                                          problem_->nparams_, t_points.data(), (int32_t)t_points.size(), ys.data(), nullptr, nullptr),
                      "dsb_batch_step_and_interpolate_host");
        return ys;
    }

    // per-instance OdeSolverError of the last solve (DSB_STATUS_*; 0 = Ok)
    std::vector<int32_t> status() const {
        std::vector<int32_t> s((size_t)problem_->nbatch_);
        detail::check(dsb_batch_get_status(batch_, s.data()), "dsb_batch_get_status");
        return s;
    }
    // OdeSolverMethod::get_statistics + the rhs op's statistics, for instance b
    OdeSolverStatistics get_statistics(int64_t b) const {
        std::vector<int64_t> raw = statistics_array();
        const int64_t* r = raw.data() + (size_t)b * DSB_NSTATS;
        OdeSolverStatistics s;
        s.number_of_linear_solver_setups = r[0];
        s.number_of_linear_solver_setups_from_checkpoint = r[1];
        s.number_of_linear_solver_setups_from_first_convergence_fail = r[2];
        s.number_of_linear_solver_setups_from_second_convergence_fail = r[3];
        s.number_of_linear_solver_setups_from_error_test_fail = r[4];
        s.number_of_linear_solver_setups_from_step_success = r[5];
        s.number_of_steps = r[6];
        s.number_of_error_test_failures = r[7];
        s.number_of_nonlinear_solver_iterations = r[8];
        s.number_of_nonlinear_solver_fails = r[9];
        s.rhs_number_of_calls = r[10];
        s.rhs_number_of_jac_muls = r[11];
        s.rhs_number_of_matrix_evals = r[12];
        return s;
    }
    std::vector<int64_t> statistics_array() const {              // [nbatch][DSB_NSTATS]
        std::vector<int64_t> raw((size_t)problem_->nbatch_ * DSB_NSTATS);
        detail::check(dsb_batch_get_stats(batch_, raw.data()), "dsb_batch_get_stats");
        return raw;
    }
    // how each instance's last solve ended: state().t and OdeSolverStopReason::RootFound(t, idx)
    std::vector<StopInfo> stop_info() const {
        const size_t B = (size_t)problem_->nbatch_;
        std::vector<double> t(B);
        std::vector<int32_t> idx(B), nc(B);
        detail::check(dsb_batch_get_final_state(batch_, t.data(), nullptr, nullptr), "dsb_batch_get_final_state");
        detail::check(dsb_batch_get_root_info(batch_, idx.data(), nc.data()), "dsb_batch_get_root_info");
        std::vector<StopInfo> out(B);
        for (size_t b = 0; b < B; ++b) { out[b].root_found = idx[b] >= 0; out[b].root_index = idx[b]; out[b].t = t[b]; out[b].ncols = nc[b]; }
        return out;
    }
    float last_kernel_ms() const { float ms = 0; detail::check(dsb_batch_last_kernel_ms(batch_, &ms), "dsb_batch_last_kernel_ms"); return ms; }

private:
    const OdeSolverProblem* problem_;
    int32_t method_;
    dsb_batch* batch_ = nullptr;
};

inline BatchedSolver OdeSolverProblem::bdf() const { return BatchedSolver(*this, DSB_METHOD_BDF); }
inline BatchedSolver OdeSolverProblem::tr_bdf2() const { return BatchedSolver(*this, DSB_METHOD_TR_BDF2); }
inline BatchedSolver OdeSolverProblem::esdirk34() const { return BatchedSolver(*this, DSB_METHOD_ESDIRK34); }

}  // namespace diffsol_b200
