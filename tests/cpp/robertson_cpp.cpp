// Host-side C++ mirror (include/diffsol_b200.hpp) exercised the way the reference's own tests and examples use its
// API (ode_solver/bdf.rs:2175-2197 test_bdf_nalgebra_robertson; examples/*): OdeBuilder -> problem.bdf() ->
// solve_dense -> get_statistics.  Prints one line per instance: status, the 13 counters, the last column.
// Without a CUDA device the solver construction throws DiffsolError (no CPU fallback): exit code 3.
//   usage: robertson_cpp <method: bdf|tr_bdf2|esdirk34|sens> <nbatch>
#include <cstdio>
#include <cstdlib>
#include <string>
#include "diffsol_b200.hpp"

using namespace diffsol_b200;

int main(int argc, char** argv) {
    const std::string method = argc > 1 ? argv[1] : "bdf";
    const int nbatch = argc > 2 ? std::atoi(argv[2]) : 4;
    // builder validation happens on the host, before any device is touched
    try {
        OdeBuilder().rhs_implicit("no_such_equations");
        return 10;
    } catch (const DiffsolError& e) { if (e.code() != DSB_BAD_ARG) return 11; }
    try {
        OdeBuilder().rhs_implicit("robertson_dae").p({0.04, 1.0e4}).build();      // 2 values for 3 parameters
        return 12;
    } catch (const DiffsolError& e) { if (e.code() != DSB_BAD_ARG) return 13; }

    if (method == "sens") {
        // problem.bdf_sens()?.solve_dense_sensitivities(t_eval) on the reference's exponential-decay sensitivity problem
        // (test_models/exponential_decay.rs:703-742), k swept over the batch: prints y0, dy0/dk, dy0/dy0 at the last time
        std::vector<double> ps;
        for (int b = 0; b < nbatch; ++b) { ps.push_back(0.1 * (1.0 + 0.25 * b)); ps.push_back(1.0); }
        OdeSolverProblem pr = OdeBuilder().rhs_implicit("exp_decay").p(ps).sens_rtol(1e-6).sens_atol({1e-6, 1e-6}).build();
        try {
            BatchedSolver solver = pr.bdf();
            const std::vector<double> t_eval = {1.0, 2.0, 5.0};
            auto res = solver.solve_dense_sensitivities(t_eval);
            std::vector<int32_t> status = solver.status();
            for (int b = 0; b < nbatch; ++b)
                std::printf("%d %d %.17g %.17g %.17g\n", b, status[b], res.first(b, 0, 2), res.second(b, 0, 2), res.second(b, 2, 2));
        } catch (const DiffsolError& e) {
            std::fprintf(stderr, "DiffsolError(%d): %s\n", e.code(), e.what());
            return 3;
        }
        return 0;
    }
    std::vector<double> p;
    for (int b = 0; b < nbatch; ++b) { p.push_back(0.04 * (1.0 + 0.125 * b)); p.push_back(1.0e4); p.push_back(3.0e7); }
    OdeSolverProblem problem = OdeBuilder().rhs_implicit("robertson_dae").p(p).rtol(1e-4).atol({1e-8, 1e-6, 1e-6}).build();
    if (problem.nstates() != 3 || problem.nparams() != 3 || !problem.has_mass() || problem.nbatch() != nbatch) return 14;
    try {
        BatchedSolver solver = method == "bdf" ? problem.bdf() : method == "tr_bdf2" ? problem.tr_bdf2() : problem.esdirk34();
        const std::vector<double> t_eval = {0.4, 4.0, 40.0, 400.0, 4000.0, 40000.0};
        DenseBlocks ys = solver.solve_dense(t_eval);
        std::vector<int32_t> status = solver.status();
        std::vector<StopInfo> stop = solver.stop_info();
        for (int b = 0; b < nbatch; ++b) {
            OdeSolverStatistics s = solver.get_statistics(b);
            std::printf("%d %d %lld %lld %lld %lld %lld %lld %lld %lld %lld %lld %lld %lld %lld %.17g %.17g %.17g %.17g %d\n", b, status[b],
                        (long long)s.number_of_linear_solver_setups, (long long)s.number_of_linear_solver_setups_from_checkpoint,
                        (long long)s.number_of_linear_solver_setups_from_first_convergence_fail,
                        (long long)s.number_of_linear_solver_setups_from_second_convergence_fail,
                        (long long)s.number_of_linear_solver_setups_from_error_test_fail,
                        (long long)s.number_of_linear_solver_setups_from_step_success, (long long)s.number_of_steps,
                        (long long)s.number_of_error_test_failures, (long long)s.number_of_nonlinear_solver_iterations,
                        (long long)s.number_of_nonlinear_solver_fails, (long long)s.rhs_number_of_calls,
                        (long long)s.rhs_number_of_jac_muls, (long long)s.rhs_number_of_matrix_evals,
                        ys(b, 0, 5), ys(b, 1, 5), ys(b, 2, 5), stop[b].t, stop[b].root_index);
        }
    } catch (const DiffsolError& e) {
        std::fprintf(stderr, "DiffsolError(%d): %s\n", e.code(), e.what());
        return 3;
    }
    return 0;
}
