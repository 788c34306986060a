"""GPU parity tests of event handling: root finding inside the lane kernels (BDF and (E)SDIRK,
on-chip and banded) against the oracle: the reference's exponential-decay-with-root problem swept over rate and initial
value, and the battery model with its voltage cut-offs (spm.ds `stop_i`) swept over the applied current."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dsb():
    import diffsol_b200
    from diffsol_b200 import capi
    capi.require_device()
    return diffsol_b200


def sweep(B):
    from diffsol_b200 import sweeps
    idx = np.arange(B)
    k = 0.02 * 50.0 ** sweeps.uniform(idx, 0)             # 0.02 .. 1
    y0 = 0.4 + 1.6 * sweeps.uniform(idx, 1)               # 0.4 .. 2: a quarter starts below the root level
    return np.stack([k, y0], axis=1)


@pytest.mark.parametrize("tol", [1e-6, 1e-9])
def test_roots_bit_exact(dsb, oracle, tol):
    B = 3000
    p = sweep(B)
    t_eval = np.arange(1.0, 21.0)
    solver = dsb.OdeBuilder().rhs_implicit("exp_decay_root").p(p).rtol(tol).atol([tol]).build().bdf()
    ys = solver.solve_dense(t_eval)
    root_idx, ncols = solver.root_info()
    t_fin = solver.final_state()[0]
    desc = oracle.make_desc("exp_decay_root", powmode=1, rtol=tol, atol=[tol])
    ys_o, stats_o, status_o, t_root_o, root_idx_o, ncols_o = oracle.batch_solve_dense_roots(desc, p, t_eval)
    assert np.array_equal(solver.status(), status_o) and (status_o == 0).all()
    assert np.array_equal(root_idx, root_idx_o) and np.array_equal(ncols, ncols_o)
    assert np.array_equal(solver.statistics_array()[:, :13], stats_o[:, :13])
    assert np.array_equal(ys, ys_o, equal_nan=True)
    stopped = root_idx_o == 0
    assert 0 < stopped.sum() < B                          # both kinds of instance are present
    assert np.array_equal(t_fin[stopped], t_root_o[stopped])
    # the state at the root sits in the last written column
    assert np.abs(ys[stopped, ncols[stopped] - 1, 0] - 0.6).max() < 1e-5


@pytest.mark.parametrize("method", ["tr_bdf2", "esdirk34"])
def test_roots_sdirk_bit_exact(dsb, oracle, method):
    """Rk::step_accepted's root check (runge_kutta.rs:935-948) in the SDIRK lane kernel."""
    B = 2000
    p = sweep(B)
    t_eval = np.arange(1.0, 21.0)
    solver = getattr(dsb.OdeBuilder().rhs_implicit("exp_decay_root").p(p).build(), method)()
    ys = solver.solve_dense(t_eval)
    root_idx, ncols = solver.root_info()
    t_fin = solver.final_state()[0]
    desc = oracle.make_desc("exp_decay_root", method=method, powmode=1)
    ys_o, stats_o, status_o, t_root_o, root_idx_o, ncols_o = oracle.batch_solve_dense_roots(desc, p, t_eval)
    assert np.array_equal(solver.status(), status_o) and (status_o == 0).all()
    assert np.array_equal(root_idx, root_idx_o) and np.array_equal(ncols, ncols_o)
    assert np.array_equal(solver.statistics_array()[:, :13], stats_o[:, :13])
    assert np.array_equal(ys, ys_o, equal_nan=True)
    stopped = root_idx_o == 0
    assert 0 < stopped.sum() < B
    assert np.array_equal(t_fin[stopped], t_root_o[stopped])


@pytest.mark.parametrize("model,method,B,block", [("spm_stop", "bdf", 600, "768"), ("spm_stop", "bdf", 600, "128"),
                                                   ("spm_stop", "tr_bdf2", 300, "768"), ("spm_stop", "esdirk34", 300, "128"),
                                                   ("spm99_stop", "bdf", 60, "768"), ("spm99_stop", "tr_bdf2", 40, "128")])
def test_battery_voltage_cut_off_bit_exact(dsb, oracle, model, method, B, block, monkeypatch):
    """BASELINE config 5 with the model text's stop function on the banded lane kernels: every instance ends where the
    terminal voltage reaches 3.105 V (or at 3600 s), at bit-identical root times, states and counters."""
    from diffsol_b200 import sweeps
    monkeypatch.setenv("DSB_BAND_BLOCK", block)
    cur = (0.6 + 0.8 * sweeps.uniform(np.arange(B), 0)).reshape(-1, 1)
    t_eval = np.arange(1, 121) * 30.0
    solver = getattr(dsb.OdeBuilder().rhs_implicit(model).p(cur).use_coloring(True).build(), method)()
    ys = solver.solve_dense(t_eval)
    root_idx, ncols = solver.root_info()
    t_fin = solver.final_state()[0]
    desc = oracle.make_desc(model, method=method, powmode=1, use_coloring=True)
    ys_o, stats_o, status_o, t_root_o, root_idx_o, ncols_o = oracle.batch_solve_dense_roots(desc, cur, t_eval)
    assert np.array_equal(solver.status(), status_o) and (status_o == 0).all()
    assert np.array_equal(root_idx, root_idx_o) and np.array_equal(ncols, ncols_o)
    assert np.array_equal(solver.statistics_array()[:, :13], stats_o[:, :13])
    assert np.array_equal(ys, ys_o, equal_nan=True)
    stopped = root_idx_o == 0
    assert 0 < stopped.sum() < B                          # small currents run to 3600 s, the others hit the cut-off
    assert np.array_equal(t_fin[stopped], t_root_o[stopped])
    capacity = cur[stopped, 0] * t_fin[stopped] / 3600.0
    assert np.all((capacity > 0.66) & (capacity < 0.69))
    # the result is the model's output function (terminal voltage): one row per column, 3.105 V at the root
    assert ys.shape == (B, len(t_eval), 1)
    assert np.abs(ys[stopped, ncols[stopped] - 1, 0] - 3.105).max() < 1e-6


@pytest.mark.parametrize("method", ["bdf", "tr_bdf2", "esdirk34"])
@pytest.mark.parametrize("tol", [1e-6, 1e-9])
def test_reset_bit_exact(dsb, oracle, tol, method):
    """Reset functions (OdeEquations::reset): the reference's exponential_decay_with_reset_problem swept over rate and
    initial value -- every root applies y -> 0.4 and the solve runs on to the last t_eval (method.rs:783-797,
    state.rs:246-270, bdf.rs:1291-1318; Rk::start_step runge_kutta.rs:446-464 for the SDIRK methods)."""
    B = 2000
    idx = np.arange(B)
    from diffsol_b200 import sweeps
    p = np.stack([0.02 * 50.0 ** sweeps.uniform(idx, 0), 0.35 + 1.65 * sweeps.uniform(idx, 1)], axis=1)
    t_eval = np.arange(1.0, 41.0)
    solver = getattr(dsb.OdeBuilder().rhs_implicit("exp_decay_reset").p(p).rtol(tol).atol([tol]).build(), method)()
    ys = solver.solve_dense(t_eval)
    root_idx, ncols = solver.root_info()
    desc = oracle.make_desc("exp_decay_reset", method=method, powmode=1, rtol=tol, atol=[tol])
    ys_o, stats_o, status_o, t_root_o, root_idx_o, ncols_o = oracle.batch_solve_dense_roots(desc, p, t_eval)
    assert np.array_equal(solver.status(), status_o) and (status_o == 0).all()
    assert np.array_equal(root_idx, root_idx_o) and (root_idx == -1).all() and np.array_equal(ncols, ncols_o)
    assert np.array_equal(solver.statistics_array()[:, :13], stats_o[:, :13])
    assert np.array_equal(ys, ys_o, equal_nan=True)


@pytest.mark.parametrize("method", ["bdf", "tr_bdf2", "esdirk34"])
def test_ball_bounce_bit_exact(dsb, oracle, method):
    """The reference's bouncing ball (ode_solver/mod.rs:1001-1080; known answers bdf.rs:2691-2697 pinned on the oracle in
    tests/test_oracle_roots.py) swept over gravity, drop height and restitution: several root + reset events per
    instance inside the lane kernel."""
    B = 3000
    idx = np.arange(B)
    from diffsol_b200 import sweeps
    p = np.stack([5.0 + 10.0 * sweeps.uniform(idx, 0), 2.0 + 18.0 * sweeps.uniform(idx, 1), 0.8 + 0.15 * sweeps.uniform(idx, 2)], axis=1)
    t_eval = np.linspace(0.05, 4.0, 80)      # ends before the earliest accumulation point of bounces (Zeno time) in the sweep
    solver = getattr(dsb.OdeBuilder().rhs_implicit("ball_bounce").p(p).build(), method)()
    ys = solver.solve_dense(t_eval)
    root_idx, ncols = solver.root_info()
    desc = oracle.make_desc("ball_bounce", method=method, powmode=1)
    ys_o, stats_o, status_o, t_root_o, root_idx_o, ncols_o = oracle.batch_solve_dense_roots(desc, p, t_eval)
    assert np.array_equal(solver.status(), status_o) and (status_o == 0).all()
    assert np.array_equal(root_idx, root_idx_o) and np.array_equal(ncols, ncols_o)
    assert np.array_equal(solver.statistics_array()[:, :13], stats_o[:, :13])
    assert np.array_equal(ys, ys_o, equal_nan=True)


@pytest.mark.parametrize("method", ["bdf", "tr_bdf2", "esdirk34"])
def test_root_index_bit_exact(dsb, oracle, method):
    """exponential_decay_with_two_roots_problem (test_models/exponential_decay.rs:890-912; test_root_found_index,
    ode_solver/mod.rs:1187-1220): which of the two root functions fired, per instance."""
    B = 3000
    idx = np.arange(B)
    from diffsol_b200 import sweeps
    p = np.stack([0.02 * 50.0 ** sweeps.uniform(idx, 0), 0.1 + 1.4 * sweeps.uniform(idx, 1)], axis=1)
    t_eval = np.arange(1.0, 21.0)
    solver = getattr(dsb.OdeBuilder().rhs_implicit("exp_decay_two_roots").p(p).build(), method)()
    ys = solver.solve_dense(t_eval)
    root_idx, ncols = solver.root_info()
    desc = oracle.make_desc("exp_decay_two_roots", method=method, powmode=1)
    ys_o, stats_o, status_o, t_root_o, root_idx_o, ncols_o = oracle.batch_solve_dense_roots(desc, p, t_eval)
    assert np.array_equal(solver.status(), status_o) and (status_o == 0).all()
    assert np.array_equal(root_idx, root_idx_o) and np.array_equal(ncols, ncols_o)
    assert set(root_idx.tolist()) == {-1, 0, 1}
    assert np.array_equal(solver.statistics_array()[:, :13], stats_o[:, :13])
    assert np.array_equal(ys, ys_o, equal_nan=True)


@pytest.mark.parametrize("model,B,coloring", [("exp_decay_root", 600, False), ("exp_decay_two_roots", 600, False),
                                              ("spm_stop", 120, True), ("spm_stop", 60, False), ("spm99_stop", 12, True)])
def test_roots_and_outputs_on_the_block_per_instance_path(dsb, oracle, model, B, coloring):
    """The root check of Bdf::step (bdf.rs:1566-1579), the RootFound branch of solve_dense (method.rs:774-805) and
    dense_write_out with an output function (method.rs:822-848) in the block-per-instance kernel: every thread of the
    block runs the same scalar root iteration on the shared state.  Bit-identical to the oracle and to the lane kernels."""
    from diffsol_b200 import sweeps
    if model.startswith("spm"):
        p = (0.6 + 0.8 * sweeps.uniform(np.arange(B), 0)).reshape(-1, 1)
        t_eval = np.arange(1, 121) * 30.0
    else:
        idx = np.arange(B)
        p = np.stack([0.02 * 50.0 ** sweeps.uniform(idx, 0), 0.1 + 1.9 * sweeps.uniform(idx, 1)], axis=1)
        t_eval = np.arange(1.0, 21.0)
    prob = dsb.OdeBuilder().rhs_implicit(model).p(p).use_coloring(coloring).build()
    solver = prob.bdf().set_execution("block")
    ys = solver.solve_dense(t_eval)
    root_idx, ncols = solver.root_info()
    t_fin = solver.final_state()[0]
    desc = oracle.make_desc(model, powmode=1, use_coloring=coloring)
    ys_o, stats_o, status_o, t_root_o, root_idx_o, ncols_o = oracle.batch_solve_dense_roots(desc, p, t_eval)
    assert np.array_equal(solver.status(), status_o) and (status_o == 0).all()
    assert np.array_equal(root_idx, root_idx_o) and np.array_equal(ncols, ncols_o)
    assert np.array_equal(solver.statistics_array()[:, :13], stats_o[:, :13])
    assert np.array_equal(ys, ys_o, equal_nan=True)
    stopped = root_idx_o >= 0
    assert stopped.sum() > 0
    assert np.array_equal(t_fin[stopped], t_root_o[stopped])
    # and the default kernel family of the model gives the same bits
    lane = prob.bdf()
    assert np.array_equal(lane.solve_dense(t_eval), ys, equal_nan=True)
    assert np.array_equal(lane.statistics_array(), solver.statistics_array())


@pytest.mark.parametrize("method", ["bdf", "tr_bdf2", "esdirk34"])
@pytest.mark.parametrize("model", ["exp_decay_reset", "ball_bounce"])
def test_resets_on_the_block_per_instance_path(dsb, oracle, model, method):
    """apply_reset + the modified-state branch of Bdf::step (bdf.rs:1291-1318) / Rk::start_step (runge_kutta.rs:446-464)
    in the block-per-instance kernel."""
    from diffsol_b200 import sweeps
    B = 400
    idx = np.arange(B)
    if model == "ball_bounce":
        p = np.stack([5.0 + 10.0 * sweeps.uniform(idx, 0), 2.0 + 18.0 * sweeps.uniform(idx, 1), 0.8 + 0.15 * sweeps.uniform(idx, 2)], axis=1)
        t_eval = np.linspace(0.05, 4.0, 80)
    else:
        p = np.stack([0.02 * 50.0 ** sweeps.uniform(idx, 0), 0.35 + 1.65 * sweeps.uniform(idx, 1)], axis=1)
        t_eval = np.arange(1.0, 41.0)
    prob = dsb.OdeBuilder().rhs_implicit(model).p(p).build()
    solver = getattr(prob, method)().set_execution("block")
    ys = solver.solve_dense(t_eval)
    root_idx, ncols = solver.root_info()
    desc = oracle.make_desc(model, method=method, powmode=1)
    ys_o, stats_o, status_o, t_root_o, root_idx_o, ncols_o = oracle.batch_solve_dense_roots(desc, p, t_eval)
    assert np.array_equal(solver.status(), status_o) and (status_o == 0).all()
    assert np.array_equal(root_idx, root_idx_o) and (root_idx == -1).all() and np.array_equal(ncols, ncols_o)
    assert np.array_equal(solver.statistics_array()[:, :13], stats_o[:, :13])
    assert np.array_equal(ys, ys_o, equal_nan=True)


@pytest.mark.parametrize("method", ["bdf", "tr_bdf2", "esdirk34"])
@pytest.mark.parametrize("execution", ["band", "block"])
def test_battery_cycling_with_resets_bit_exact(dsb, oracle, method, execution):
    """Resets at n > 16: the battery model cycled (spm_cycle: at a voltage cut-off the cell goes back to its charged state
    and the discharge starts again) on the banded lane kernels and on the block-per-instance kernel, all three methods."""
    from diffsol_b200 import sweeps
    B = 96 if execution == "band" else 32
    cur = (0.6 + 0.8 * sweeps.uniform(np.arange(B), 0)).reshape(-1, 1)
    t_eval = np.arange(1, 121) * 60.0
    solver = getattr(dsb.OdeBuilder().rhs_implicit("spm_cycle").p(cur).use_coloring(True).build(), method)().set_execution(execution)
    ys = solver.solve_dense(t_eval)
    root_idx, ncols = solver.root_info()
    desc = oracle.make_desc("spm_cycle", method=method, powmode=1, use_coloring=True)
    ys_o, stats_o, status_o, t_root_o, root_idx_o, ncols_o = oracle.batch_solve_dense_roots(desc, cur, t_eval)
    assert np.array_equal(solver.status(), status_o) and (status_o == 0).all()
    assert np.array_equal(root_idx, root_idx_o) and (root_idx == -1).all() and np.array_equal(ncols, ncols_o)
    assert np.array_equal(solver.statistics_array()[:, :13], stats_o[:, :13])
    assert np.array_equal(ys, ys_o, equal_nan=True)
    assert (np.diff(ys[:, :, 0], axis=1) > 0.3).sum(axis=1).min() >= 1          # every cell was recharged at least once


@pytest.mark.parametrize("method", ["bdf", "tr_bdf2", "esdirk34"])
@pytest.mark.parametrize("execution", ["lane", "block"])
def test_harness_loop_stops_at_the_root(dsb, oracle, method, execution):
    """The step()/interpolate() loop of the reference's harness (ode_solver/mod.rs:132-141): `if let RootFound(t, _) =
    method.step() { return method.interpolate(t) }` -- the state at the root takes the place of the point the loop was
    stepping towards and the loop ends (the solver's own state stays at the end of that step).  step_and_interpolate does
    the same: the later columns stay NaN, root_info() says which root and how many columns."""
    pts = np.arange(0.0, 10.0)
    params = [[0.1, 1.0], [0.2, 1.5], [0.01, 1.0]]          # the last one does not reach its root before t = 9
    solver = getattr(dsb.OdeBuilder().rhs_implicit("exp_decay_root").p(params).build(), method)().set_execution(execution)
    ys = solver.step_and_interpolate(pts)
    root_idx, ncols = solver.root_info()
    desc = oracle.make_desc("exp_decay_root", method=method, powmode=1)
    for b, p in enumerate(params):
        rc, ys_o, stats_o, fin = oracle.harness(desc, p, pts)
        assert rc == 0 and solver.status()[b] == 0
        assert np.array_equal(ys[b], ys_o, equal_nan=True) and solver.get_statistics(b) == stats_o
        assert solver.final_state()[0][b] == fin["t"]
        written = int(np.isfinite(ys_o[:, 0]).sum())
        assert ncols[b] == written and (root_idx[b] >= 0) == (written < len(pts))
    assert list(root_idx >= 0) == [True, True, False]


def test_root_info_without_roots(dsb):
    p = np.tile(np.array([[0.04, 1.0e4, 3.0e7]]), (40, 1))
    solver = dsb.OdeBuilder().rhs_implicit("robertson_ode").p(p).rtol(1e-4).atol([1e-8, 1e-14, 1e-6]).build().bdf()
    solver.solve_dense([0.4, 4.0])
    root_idx, ncols = solver.root_info()
    assert (root_idx == -1).all() and (ncols == 2).all()
