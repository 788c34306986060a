"""Resets on DAEs: `state.apply_reset_with_mass` (/root/reference/crates/diffsol/src/ode_solver/state.rs:279-306, reached from
Bdf::apply_reset, bdf.rs:1017-1020, inside solve_dense's RootFound branch, method.rs:783-797): y <- reset(y, t), then
set_consistent with a Newton solver WITHOUT line search, starting from the reset y and the dy that state_mut_back
interpolated at the root (interpolate_derivative_from_diff, bdf.rs:788-810).  Equations: the DAE of the reference's
reset-with-mass test problem (test_models/exponential_decay_with_algebraic.rs:501-560) without its sensitivities --
y' = -k y (twice), 0 = z - y, every state starts at y0, roots y[0] - 0.6 and y[0] - 2, reset y -> y + 2.

CPU: the oracle against the analytic solution of the reset problem, and the kernel SOURCE (host emulation of the on-chip BDF
lane kernel) bit for bit against the oracle.  GPU: the CUDA path against the oracle; the other kernel families refuse."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)


def sweep(B):
    from diffsol_b200 import sweeps
    i = np.arange(B)
    return np.stack([0.05 + 0.3 * sweeps.uniform(i, 0), 0.8 + 1.5 * sweeps.uniform(i, 1)], axis=1)


def analytic(k, y0, t):
    """y0 e^{-k t} until it reaches 0.6 (if it starts above) -> + 2 -> decays to 2.0 -> + 2 -> ...: the state just after the
    last event before t, decayed.  Only used where the instance starts above 0.6 and below 2."""
    y, tc = y0, 0.0
    while True:
        level = 0.6 if y < 2.0 else 2.0            # the next root below the current value
        if y <= level:
            return None
        t_hit = tc + np.log(y / level) / k
        if t_hit > t:
            return y * np.exp(-k * (t - tc))
        y, tc = level + 2.0, t_hit


def test_oracle_against_the_analytic_reset_solution(oracle):
    p = np.array([[0.1, 1.0], [0.2, 1.5], [0.15, 1.9]])
    t_eval = np.array([1.0, 3.0, 6.0, 9.0, 12.0])
    ys, st, status, t_root, ridx, nc = oracle.batch_solve_dense_roots(oracle.make_desc("exp_decay_algebraic_reset"), p, t_eval)
    assert (status == 0).all() and (nc == len(t_eval)).all()
    for b, (k, y0) in enumerate(p):
        for j, t in enumerate(t_eval):
            want = analytic(k, y0, t)
            assert want is not None and np.abs(ys[b, j] - want).max() < 2e-4 * max(1.0, want), (b, j, ys[b, j], want)
    assert np.abs(ys[:, :, 2] - ys[:, :, 1]).max() < 1e-5                  # the algebraic constraint z = y holds after every reset


def test_kernel_source_equals_oracle(oracle):
    from host_emu import emu
    p = sweep(24)
    t_eval = np.linspace(0.5, 12.0, 24)
    ys, st, status, t_root, ridx, nc = oracle.batch_solve_dense_roots(oracle.make_desc("exp_decay_algebraic_reset", powmode=1), p, t_eval)
    r = emu.solve(22, 3, 2, p, t_eval, method="bdf", kernel="lane")
    assert np.array_equal(r["status"], status) and (status == 0).all()
    assert np.array_equal(r["stats"][:, :13], st[:, :13]) and np.array_equal(r["ncols"], nc)
    assert np.array_equal(r["ys"], ys, equal_nan=True)
    assert st[:, 12].max() >= 10                   # several resets per instance: each re-evaluates the Jacobian


@pytest.fixture(scope="module")
def dsb():
    import diffsol_b200
    from diffsol_b200 import capi
    capi.require_device()
    return diffsol_b200


@pytest.mark.gpu
def test_gpu_reset_with_mass_bit_exact(dsb, oracle):
    B = 3000
    p = sweep(B)
    t_eval = np.linspace(0.5, 12.0, 24)
    solver = dsb.OdeBuilder().rhs_implicit("exp_decay_algebraic_reset").p(p).build().bdf()
    ys = solver.solve_dense(t_eval)
    ys_o, st_o, status_o, t_root_o, ridx_o, nc_o = oracle.batch_solve_dense_roots(oracle.make_desc("exp_decay_algebraic_reset", powmode=1), p, t_eval)
    assert np.array_equal(solver.status(), status_o) and (status_o == 0).all()
    assert np.array_equal(solver.statistics_array()[:, :13], st_o[:, :13])
    assert np.array_equal(ys, ys_o, equal_nan=True)
    assert np.abs(ys[:, :, 2] - ys[:, :, 1]).max() < 1e-5


@pytest.mark.gpu
def test_gpu_reset_with_mass_other_families_refuse(dsb):
    from diffsol_b200 import capi
    p = sweep(8)
    t_eval = np.array([1.0, 2.0])
    prob = dsb.OdeBuilder().rhs_implicit("exp_decay_algebraic_reset").p(p).build()
    with pytest.raises(capi.DiffsolB200Error):
        prob.tr_bdf2().solve_dense(t_eval)
    with pytest.raises(capi.DiffsolB200Error):
        prob.bdf().set_execution("block").solve_dense(t_eval)
