"""GPU parity tests of the banded WARP-per-instance BDF kernel (dsb_wband_bdf_kernel.cuh: the instance's working set in
shared memory, one-lane band LU with the reciprocal-reuse division, bulk-async parking of df/dy and bulk-async result
stores): bit-identical to the oracle's dense LU path and to the one-lane-per-instance banded kernel it replaces for
BASELINE configs 4 (heat-equation DAE, n = 256) and 5 (battery model, n = 42 / 200, with and without its output and
stop functions)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dsb():
    import diffsol_b200
    from diffsol_b200 import capi
    capi.require_device()
    return diffsol_b200


def spm_currents(B):
    from diffsol_b200 import sweeps
    return (0.6 + 0.8 * sweeps.uniform(np.arange(B), 0)).reshape(-1, 1)


def heat_params(B):
    from diffsol_b200 import sweeps
    i = np.arange(B)
    return np.stack([1.0 + sweeps.uniform(i, 0), 0.1 + 0.3 * sweeps.uniform(i, 1), 0.6 + 0.3 * sweeps.uniform(i, 2)], axis=1)


HEAT_T_EVAL = np.arange(1, 101) / 100.0 * 0.99


def check_against_oracle(solver, ys, oracle, model, params, t_eval, **kw):
    desc = oracle.make_desc(model, powmode=1, **kw)
    ys_o, stats_o, status_o = oracle.batch_solve_dense(desc, params, t_eval)
    assert np.array_equal(solver.status(), status_o) and (status_o == 0).all()
    assert np.array_equal(solver.statistics_array()[:, :13], stats_o[:, :13])
    assert np.array_equal(ys, ys_o)


@pytest.mark.parametrize("model,B,coloring", [("spm", 700, False), ("spm", 700, True), ("spm99", 100, True), ("spm99", 40, False)])
def test_battery_model_bit_exact(dsb, oracle, model, B, coloring):
    current = spm_currents(B)
    t_eval = np.arange(1, 13) * 300.0
    solver = dsb.OdeBuilder().rhs_implicit(model).p(current).use_coloring(coloring).build().bdf().set_execution("warp")
    ys = solver.solve_dense(t_eval)
    check_against_oracle(solver, ys, oracle, model, current, t_eval, use_coloring=coloring)


def test_battery_model_tight_tolerances(dsb, oracle):
    """Higher orders and more rescales."""
    current = spm_currents(96)
    t_eval = np.arange(1, 7) * 500.0
    solver = dsb.OdeBuilder().rhs_implicit("spm").p(current).rtol(1e-9).atol(1e-10).build().bdf().set_execution("warp")
    ys = solver.solve_dense(t_eval)
    check_against_oracle(solver, ys, oracle, "spm", current, t_eval, rtol=1e-9, atol=1e-10)


@pytest.mark.parametrize("model,B,coloring", [("heat1d_dae_32", 300, False), ("heat1d_dae_32", 300, True), ("heat1d_dae_256", 40, True),
                                               ("heat1d_dae_32_bc", 100, True)])
def test_heat_dae_bit_exact(dsb, oracle, model, B, coloring):
    """Singular mass: consistent initialisation by dsb_band_init_kernel, M - cJ in band storage (M parked beside df/dy)."""
    p = heat_params(B)
    solver = (dsb.OdeBuilder().rhs_implicit(model).p(p).rtol(1e-6).atol(1e-6).use_coloring(coloring).build().bdf()
              .set_execution("warp"))
    ys = solver.solve_dense(HEAT_T_EVAL)
    check_against_oracle(solver, ys, oracle, model, p, HEAT_T_EVAL, rtol=1e-6, atol=1e-6, use_coloring=coloring)


@pytest.mark.parametrize("model,B,nt", [("spm_stop", 240, 120), ("spm_stop", 64, 1200), ("spm99_stop", 24, 120)])
def test_voltage_output_and_cut_off_bit_exact(dsb, oracle, model, B, nt):
    """out (terminal voltage, up to 32 pending columns evaluated at once, one per lane) and stop (voltage cut-offs: the root
    check of Bdf::step and the RootFound branch of solve_dense)."""
    p = spm_currents(B)
    t_eval = np.arange(1, nt + 1) * (3600.0 / nt)
    solver = dsb.OdeBuilder().rhs_implicit(model).p(p).use_coloring(True).build().bdf().set_execution("warp")
    ys = solver.solve_dense(t_eval)
    root_idx, ncols = solver.root_info()
    t_fin = solver.final_state()[0]
    desc = oracle.make_desc(model, powmode=1, use_coloring=True)
    ys_o, stats_o, status_o, t_root_o, root_idx_o, ncols_o = oracle.batch_solve_dense_roots(desc, p, t_eval)
    assert np.array_equal(solver.status(), status_o) and (status_o == 0).all()
    assert np.array_equal(root_idx, root_idx_o) and np.array_equal(ncols, ncols_o)
    assert np.array_equal(solver.statistics_array()[:, :13], stats_o[:, :13])
    assert np.array_equal(ys, ys_o, equal_nan=True)
    stopped = root_idx_o >= 0
    assert stopped.sum() > 0
    assert np.array_equal(t_fin[stopped], t_root_o[stopped])


@pytest.mark.parametrize("model,B", [("heat1d_dae_256", 1500), ("spm", 20000), ("spm99", 3000), ("spm_stop", 6000)])
def test_equals_the_lane_per_instance_kernel_on_big_batches(dsb, model, B):
    """More instances than resident warps (persistent warps drawing from the work counter, ragged tail): the same bits as
    the one-lane-per-instance banded kernel."""
    if model.startswith("heat"):
        b = dsb.OdeBuilder().rhs_implicit(model).p(heat_params(B)).rtol(1e-6).atol(1e-6)
        t_eval = HEAT_T_EVAL
    else:
        b = dsb.OdeBuilder().rhs_implicit(model).p(spm_currents(B)).use_coloring(True)
        t_eval = np.arange(1, 121) * 30.0
    prob = b.build()
    w = prob.bdf().set_execution("warp")
    l = prob.bdf().set_execution("band")
    yw, yl = w.solve_dense(t_eval), l.solve_dense(t_eval)
    assert np.array_equal(yw, yl, equal_nan=True)
    assert np.array_equal(w.statistics_array(), l.statistics_array())
    assert np.array_equal(w.status(), l.status())
    assert np.array_equal(w.final_state()[0], l.final_state()[0])
    # automatic selection takes the warp-per-instance kernel for BDF on banded models
    a = prob.bdf()
    assert np.array_equal(a.solve_dense(t_eval), yw, equal_nan=True)


def test_exact_solve_path(dsb, oracle, monkeypatch):
    """The back substitution's fall-back (right-hand side rebuilt, plain IEEE divisions) forced on every Newton iteration."""
    monkeypatch.setenv("DSB_WBAND_FORCE_REDO", "1")
    p = heat_params(200)
    solver = dsb.OdeBuilder().rhs_implicit("heat1d_dae_32").p(p).rtol(1e-6).atol(1e-6).build().bdf().set_execution("warp")
    ys = solver.solve_dense(HEAT_T_EVAL)
    check_against_oracle(solver, ys, oracle, "heat1d_dae_32", p, HEAT_T_EVAL, rtol=1e-6, atol=1e-6)


def test_device_resident_entry_point_returns_batch_major(dsb):
    """dsb_batch_solve_dense keeps its [nt][nout][B] contract: the instance-major block the kernel writes is re-laid out."""
    import torch
    B = 500
    p = heat_params(B)
    prob = dsb.OdeBuilder().rhs_implicit("heat1d_dae_32").p(p).rtol(1e-6).atol(1e-6).build()
    s = prob.bdf().set_execution("warp")
    ys_host = s.solve_dense(HEAT_T_EVAL)
    out = torch.empty((len(HEAT_T_EVAL), 32, B), dtype=torch.float64, device="cuda")
    s.set_params()
    s.solve_dense_device(HEAT_T_EVAL, out.data_ptr())
    torch.cuda.synchronize()
    assert np.array_equal(out.permute(2, 0, 1).cpu().numpy(), ys_host)


def test_free_running_loop(dsb, oracle):
    """The step()/interpolate() loop of the reference's harness (ode_solver/mod.rs:132-141) on the warp kernel."""
    B = 64
    current = spm_currents(B)
    pts = np.arange(1, 7) * 500.0
    solver = dsb.OdeBuilder().rhs_implicit("spm").p(current).build().bdf().set_execution("warp")
    ys = solver.step_and_interpolate(pts)
    desc = oracle.make_desc("spm", powmode=1)
    for b in range(0, B, 9):
        rc, ys_o, stats_o, fin = oracle.harness(desc, current[b], pts)
        assert rc == 0 and np.array_equal(ys[b], ys_o)
        assert {n: int(solver.statistics_array()[b, i]) for i, n in enumerate(oracle.S_NAMES)} == stats_o
