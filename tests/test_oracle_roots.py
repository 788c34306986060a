"""CPU tests of the oracle's event handling (RootFinder, nonlinear_solver/root.rs; Bdf::step root check,
bdf.rs:1566-1579; fn solve_dense RootFound branch, method.rs:774-805, 493-503) against the reference's own tests of
the same problem: exponential_decay_problem_with_root (test_models/exponential_decay.rs:370-390, root y[0] - 0.6)."""
import numpy as np


def weighted_norm(y, ystar, atol, rtol):
    return float(np.sqrt(np.mean(((y - ystar) / (np.abs(ystar) * rtol + atol)) ** 2)))


def test_dense_solve_stops_on_root(oracle):
    """ode_solver/method.rs:1141-1163 test_dense_solve_stops_on_root, same assertions."""
    k, y0 = 0.1, 1.0
    desc = oracle.make_desc("exp_decay_root")           # builder defaults rtol = atol = 1e-6, libm pow
    t_eval = np.arange(0.0, 11.0)
    ys, stats, status, t_root_found, root_idx, ncols = oracle.batch_solve_dense_roots(desc, [[k, y0]], t_eval)
    assert status[0] == 0 and root_idx[0] == 0           # OdeSolverStopReason::RootFound(_, 0)
    t_root = -np.log(0.6) / k
    assert abs(t_root_found[0] - t_root) < 1e-3
    assert ncols[0] < len(t_eval)
    minus_one = int(np.argmax(t_eval >= t_root)) - 1
    expected = np.full(2, np.exp(-k * t_eval[minus_one]))
    assert weighted_norm(ys[0, minus_one], expected, 1e-6, 1e-6) < 15.0
    assert weighted_norm(ys[0, ncols[0] - 1], np.array([0.6, 0.6]), 1e-6, 1e-6) < 15.0
    assert np.isnan(ys[0, ncols[0]:]).all()


def test_root_finder_known_answer(oracle):
    """nonlinear_solver/root.rs:177-221 test_root through the integrator: the located root is within 1e-10 relative
    of where the interpolant crosses, i.e. the returned state satisfies g = 0 to the interpolation accuracy."""
    desc = oracle.make_desc("exp_decay_root", powmode=1)
    ks = np.linspace(0.05, 0.9, 18)
    p = np.stack([ks, np.full_like(ks, 1.0)], axis=1)
    ys, stats, status, t_root, root_idx, ncols = oracle.batch_solve_dense_roots(desc, p, np.arange(1.0, 41.0))
    assert (status == 0).all() and (root_idx == 0).all()
    assert np.allclose(t_root, -np.log(0.6) / ks, rtol=1e-4)
    at_root = ys[np.arange(len(ks)), ncols - 1, 0]
    assert np.abs(at_root - 0.6).max() < 1e-9            # g(y(t_root)) = 0 on the interpolant


def test_no_root_runs_to_the_end(oracle):
    desc = oracle.make_desc("exp_decay_root", powmode=1)
    t_eval = np.arange(1.0, 6.0)
    ys, stats, status, t_root, root_idx, ncols = oracle.batch_solve_dense_roots(desc, [[0.1, 0.5], [0.01, 1.0]], t_eval)
    assert (status == 0).all() and (root_idx == -1).all() and (ncols == len(t_eval)).all()
    desc2 = oracle.make_desc("exp_decay", powmode=1)
    ys2, stats2, status2 = oracle.batch_solve_dense(desc2, [[0.1, 0.5], [0.01, 1.0]], t_eval)
    assert np.array_equal(ys, ys2) and np.array_equal(stats, stats2)      # the root function does not steer the steps
