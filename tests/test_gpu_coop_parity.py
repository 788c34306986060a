"""GPU parity tests of the block-per-instance (cooperative) BDF path: the reference's small test problems
forced through it, and the heat-equation DAE (BASELINE config 4) at n = 32 and n = 256."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "reference_snapshots.json")) as f:
    GOLD = json.load(f)
BDF_CASES = [c for c in GOLD["cases"] if c["method"] == "bdf" and not c["coloring"]]
SDIRK_CASES = [c for c in GOLD["cases"] if c["method"] != "bdf"]


@pytest.fixture(scope="module")
def dsb():
    import diffsol_b200
    from diffsol_b200 import capi
    capi.require_device()
    return diffsol_b200


def heat_params(idx):
    from diffsol_b200 import sweeps
    return np.stack([1.0 + sweeps.uniform(idx, 0), 0.1 + 0.3 * sweeps.uniform(idx, 1), 0.6 + 0.3 * sweeps.uniform(idx, 2)], axis=1)


HEAT_T_EVAL = np.arange(1, 101) / 100.0 * 0.99


@pytest.mark.parametrize("case", BDF_CASES, ids=[c["name"] for c in BDF_CASES])
def test_reference_snapshots_block_per_instance(dsb, oracle, case):
    from test_oracle_golden import expected_stats, solution_points
    t, ystar = solution_points(case["points"])
    nb = 3
    b = dsb.OdeBuilder().rhs_implicit(case["model"]).rtol(case["rtol"]).atol(case["atol"]).nbatch(nb)
    if len(case["p"]):
        b = b.p(case["p"])
    solver = b.build().bdf().set_execution("block")
    ys = solver.step_and_interpolate(t)
    assert (solver.status() == 0).all()
    for k in range(nb):
        assert solver.get_statistics(k) == expected_stats(case), case["cite"]
    desc = oracle.make_desc(case["model"], rtol=case["rtol"], atol=case["atol"], powmode=1)
    rc, ys_o, stats_o, fin = oracle.harness(desc, case["p"], t)
    assert rc == 0
    for k in range(nb):
        assert np.array_equal(ys[k], ys_o), case["name"]


@pytest.mark.parametrize("case", SDIRK_CASES, ids=[c["name"] for c in SDIRK_CASES])
def test_reference_sdirk_snapshots_block_per_instance(dsb, oracle, case):
    """The six sdirk.rs statistics snapshots (TR-BDF2 / ESDIRK34; exponential decay, its DAE form, Robertson DAE and ODE)
    through the Sdirk form of the block-per-instance kernel: all 13 integers, and the states bitwise vs the oracle."""
    from test_oracle_golden import expected_stats, solution_points
    t, ystar = solution_points(case["points"])
    nb = 3
    b = dsb.OdeBuilder().rhs_implicit(case["model"]).rtol(case["rtol"]).atol(case["atol"]).nbatch(nb)
    if len(case["p"]):
        b = b.p(case["p"])
    solver = getattr(b.build(), case["method"])().set_execution("block")
    ys = solver.step_and_interpolate(t)
    assert (solver.status() == 0).all()
    for k in range(nb):
        assert solver.get_statistics(k) == expected_stats(case), case["cite"]
    desc = oracle.make_desc(case["model"], method=case["method"], rtol=case["rtol"], atol=case["atol"], powmode=1)
    rc, ys_o, stats_o, fin = oracle.harness(desc, case["p"], t)
    assert rc == 0
    for k in range(nb):
        assert np.array_equal(ys[k], ys_o), case["name"]


@pytest.mark.parametrize("method", ["tr_bdf2", "esdirk34"])
@pytest.mark.parametrize("model,B,coloring", [("robertson_dae", 300, False), ("van_der_pol_scaled", 300, False),
                                              ("heat1d_dae_32", 48, False), ("heat1d_dae_32", 48, True),
                                              ("heat1d_dae_256", 6, True), ("spm", 60, True), ("spm_stop", 60, True),
                                              ("exp_decay_two_roots", 300, False)])
def test_sdirk_block_per_instance_bit_exact(dsb, oracle, method, model, B, coloring):
    """Sdirk::step (sdirk.rs:409-543) on the block-per-instance kernel: dense and banded LU, ODEs and singular-mass DAEs,
    coloured and dense Jacobians, output and root functions -- bitwise vs the oracle, and vs the kernel family the model
    runs on by default (lane or banded lane kernels) where SDIRK exists there."""
    from diffsol_b200 import sweeps
    idx = np.arange(B)
    kw = {}
    if model == "robertson_dae":
        p, t_eval, kw = sweeps.robertson_sweep(idx), sweeps.ROBERTSON_T_EVAL, dict(sweeps.ROBERTSON_DAE_TOL)
    elif model == "van_der_pol_scaled":
        p, t_eval, kw = sweeps.van_der_pol_scaled_sweep(idx), sweeps.VAN_DER_POL_T_EVAL, dict(sweeps.VAN_DER_POL_TOL)
    elif model.startswith("heat1d"):
        p, t_eval = heat_params(idx), HEAT_T_EVAL[9::10]
    elif model.startswith("spm"):
        p, t_eval = (0.6 + 0.8 * sweeps.uniform(idx, 0)).reshape(-1, 1), np.arange(1, 41) * 90.0
    else:
        p, t_eval = np.stack([0.02 * 50.0 ** sweeps.uniform(idx, 0), 0.1 + 1.9 * sweeps.uniform(idx, 1)], axis=1), np.arange(1.0, 21.0)
    b = dsb.OdeBuilder().rhs_implicit(model).p(p).use_coloring(coloring)
    if kw:
        b = b.rtol(kw["rtol"]).atol(kw["atol"])
    prob = b.build()
    solver = getattr(prob, method)().set_execution("block")
    ys = solver.solve_dense(t_eval)
    desc = oracle.make_desc(model, method=method, powmode=1, use_coloring=coloring, **kw)
    ys_o, stats_o, status_o, t_root_o, root_idx_o, ncols_o = oracle.batch_solve_dense_roots(desc, p, t_eval)
    assert np.array_equal(solver.status(), status_o)
    assert np.array_equal(solver.statistics_array()[:, :13], stats_o[:, :13])
    ok = status_o == 0
    assert ok.sum() > 0
    assert np.array_equal(ys[ok], ys_o[ok], equal_nan=True)
    root_idx, ncols = solver.root_info()
    assert np.array_equal(root_idx[ok], root_idx_o[ok]) and np.array_equal(ncols[ok], ncols_o[ok])
    other = getattr(prob, method)()
    assert np.array_equal(other.solve_dense(t_eval)[ok], ys[ok], equal_nan=True)
    assert np.array_equal(other.statistics_array(), solver.statistics_array())


@pytest.mark.parametrize("model,tol", [("robertson_ode", "ROBERTSON_ODE_TOL"), ("robertson_dae", "ROBERTSON_DAE_TOL")])
def test_robertson_sweep_block_equals_lane_equals_oracle(dsb, oracle, model, tol):
    from diffsol_b200 import sweeps
    B = 600
    tolkw = getattr(sweeps, tol)
    p = sweeps.robertson_sweep(np.arange(B))
    prob = dsb.OdeBuilder().rhs_implicit(model).p(p).rtol(tolkw["rtol"]).atol(tolkw["atol"]).build()
    s_block = prob.bdf().set_execution("block")
    ys_b = s_block.solve_dense(sweeps.ROBERTSON_T_EVAL)
    s_lane = prob.bdf().set_execution("lane")
    ys_l = s_lane.solve_dense(sweeps.ROBERTSON_T_EVAL)
    desc = oracle.make_desc(model, powmode=1, **tolkw)
    ys_o, stats_o, status_o = oracle.batch_solve_dense(desc, p, sweeps.ROBERTSON_T_EVAL)
    for s, ys in ((s_block, ys_b), (s_lane, ys_l)):
        assert np.array_equal(s.status(), status_o)
        assert np.array_equal(s.statistics_array()[:, :13], stats_o[:, :13])
        assert np.array_equal(ys, ys_o)


@pytest.mark.parametrize("model,B", [("heat1d_dae_32", 96), ("heat1d_dae_256", 12)])
def test_heat_dae_sweep_bit_exact(dsb, oracle, model, B):
    """BASELINE config 4 (1-D heat equation, boundary rows algebraic, initial-condition sweep) at sizes the
    oracle finishes in seconds: counters, status and all 100 dense-output columns bitwise."""
    p = heat_params(np.arange(B))
    solver = dsb.OdeBuilder().rhs_implicit(model).p(p).rtol(1e-6).atol(1e-6).build().bdf().set_execution("block")
    ys = solver.solve_dense(HEAT_T_EVAL)
    desc = oracle.make_desc(model, powmode=1, rtol=1e-6, atol=1e-6)
    ys_o, stats_o, status_o = oracle.batch_solve_dense(desc, p, HEAT_T_EVAL)
    assert np.array_equal(solver.status(), status_o) and (status_o == 0).all()
    assert np.array_equal(solver.statistics_array()[:, :13], stats_o[:, :13])
    assert np.array_equal(ys, ys_o)
    # physics sanity: heat is lost through the cold boundaries, nothing grows
    assert (ys[:, -1].max(axis=1) < p[:, 0]).all()


def test_unsupported_combinations_fail_loudly(dsb):
    p = heat_params(np.arange(2))
    prob = dsb.OdeBuilder().rhs_implicit("heat1d_dae_32").p(p).build()
    with pytest.raises(dsb.DiffsolB200Error):
        prob.bdf().set_execution("lane").solve_dense([0.5])


def test_spm_battery_sweep_bit_exact(dsb, oracle):
    """BASELINE config 5 (single-particle battery model of the reference's battery example, n = 42, applied
    current sweep I = 0.6 + 0.8 u) at a size the oracle finishes in seconds.  States only: the model's
    output / stop functions are outside the implicit step loop."""
    from diffsol_b200 import sweeps
    B = 300
    current = (0.6 + 0.8 * sweeps.uniform(np.arange(B), 0)).reshape(-1, 1)
    t_eval = np.arange(1, 13) * 300.0
    solver = dsb.OdeBuilder().rhs_implicit("spm").p(current).build().bdf().set_execution("block")
    ys = solver.solve_dense(t_eval)
    desc = oracle.make_desc("spm", powmode=1)
    ys_o, stats_o, status_o = oracle.batch_solve_dense(desc, current, t_eval)
    assert np.array_equal(solver.status(), status_o) and (status_o == 0).all()
    assert np.array_equal(solver.statistics_array()[:, :13], stats_o[:, :13])
    assert np.array_equal(ys, ys_o)
    # discharge capacity is I t / 3600 exactly up to the integration tolerance
    assert np.allclose(ys[:, -1, 0], current[:, 0] * 3600.0 / 3600.0, rtol=1e-5)


@pytest.mark.parametrize("model,B", [("heat1d_dae_32", 40), ("heat1d_dae_256", 6), ("spm", 40)])
def test_coloured_jacobian_block_per_instance(dsb, oracle, model, B):
    """use_coloring(true) on the block-per-instance path: 3 jac_mul calls per Jacobian instead of n (plus the n
    sparsity probes the reference counts once), same trajectory."""
    from diffsol_b200 import sweeps
    if model == "spm":
        p = (0.6 + 0.8 * sweeps.uniform(np.arange(B), 0)).reshape(-1, 1)
        t_eval = np.arange(1, 5) * 300.0
    else:
        p = heat_params(np.arange(B))
        t_eval = HEAT_T_EVAL[:20]
    solver = (dsb.OdeBuilder().rhs_implicit(model).p(p).rtol(1e-6).atol(1e-6).use_coloring(True).build().bdf()
              .set_execution("block"))
    ys = solver.solve_dense(t_eval)
    desc = oracle.make_desc(model, powmode=1, rtol=1e-6, atol=1e-6, use_coloring=True)
    ys_o, stats_o, status_o = oracle.batch_solve_dense(desc, p, t_eval)
    assert np.array_equal(solver.status(), status_o) and (status_o == 0).all()
    assert np.array_equal(solver.statistics_array()[:, :13], stats_o[:, :13])
    assert np.array_equal(ys, ys_o)
    n = ys.shape[2]
    me = solver.statistics_array()[:, 12]
    assert (solver.statistics_array()[:, 11] == 3 * me + n).all()      # 3 colours + n probes


def test_reference_snapshot_heat2d_on_gpu(dsb, oracle):
    """The reference's only statistics snapshot with n > 16 (bdf.rs:2424-2446: 2-D heat equation DAE, 10 x 10 grid,
    n = 100, coloured Jacobian) on the block-per-instance path (band LU in shared memory: kl = ku = 10), through the
    stepping loop of the reference's harness: all 13 integers, for every instance of a batch, and the states
    bit-identical to the oracle."""
    import json
    import os
    from test_oracle_golden import expected_stats
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_snapshots.json")) as f:
        case = json.load(f)["heat2d_10"]
    nb = 9
    solver = (dsb.OdeBuilder().rhs_implicit("heat2d_10").nbatch(nb).rtol(case["rtol"]).atol(case["atol"])
              .use_coloring(True).build().bdf())
    ys = solver.step_and_interpolate(case["t"])
    assert (solver.status() == 0).all()
    for b in range(nb):
        assert solver.get_statistics(b) == expected_stats(case), case["cite"]
    desc = oracle.make_desc("heat2d_10", rtol=case["rtol"], atol=case["atol"], use_coloring=True, powmode=1)
    rc, ys_o, stats_o, fin = oracle.harness(desc, [], case["t"])
    assert rc == 0
    for b in range(nb):
        assert np.array_equal(ys[b], ys_o)
    out = (np.sqrt((ys[0] ** 2).sum(axis=1)) / 9.0) ** 2
    expected = np.array(case["out"])
    assert (np.abs(out - expected) / (np.abs(expected) * 1e-5 + 1e-5)).max() < 20.0
