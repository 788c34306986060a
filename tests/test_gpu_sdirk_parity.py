"""GPU parity tests of the (E)SDIRK path (TR-BDF2, ESDIRK34) through the C ABI against the CPU oracle and,
through `step_and_interpolate`, directly against the reference's sdirk.rs statistics snapshots."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "reference_snapshots.json")) as f:
    GOLD = json.load(f)
RK_CASES = [c for c in GOLD["cases"] if c["method"] != "bdf"]


@pytest.fixture(scope="module")
def dsb():
    import diffsol_b200
    from diffsol_b200 import capi
    capi.require_device()
    return diffsol_b200


def _solver(dsb, model, method, p, rtol, atol, **opts):
    b = dsb.OdeBuilder().rhs_implicit(model).rtol(rtol).atol(atol).p(p)
    if opts:
        b = b.ode_options(**opts)
    return getattr(b.build(), method)()


@pytest.mark.parametrize("case", RK_CASES, ids=[c["name"] for c in RK_CASES])
def test_reference_snapshots_on_gpu(dsb, oracle, case):
    from test_oracle_golden import expected_stats, solution_points
    t, ystar = solution_points(case["points"])
    nb = 3
    solver = _solver(dsb, case["model"], case["method"], np.tile(case["p"], (nb, 1)), case["rtol"], case["atol"])
    ys = solver.step_and_interpolate(t)
    assert (solver.status() == 0).all()
    for b in range(nb):
        assert solver.get_statistics(b) == expected_stats(case), case["cite"]
    desc = oracle.make_desc(case["model"], method=case["method"], rtol=case["rtol"], atol=case["atol"], powmode=1)
    rc, ys_o, stats_o, fin = oracle.harness(desc, case["p"], t)
    assert rc == 0
    for b in range(nb):
        assert np.array_equal(ys[b], ys_o), case["name"]
    tf, hf, of = solver.final_state()
    assert tf[0] == fin["t"] and hf[0] == fin["h"] and of[0] == fin["order"]
    n = ystar.shape[1]
    atol = np.array(case["atol"] * n if len(case["atol"]) == 1 else case["atol"])
    w = np.abs(ystar) * case["rtol"] + atol
    assert (np.sqrt(np.mean(((ys[0] - ystar) / w) ** 2, axis=1)) < 20.0).all()


@pytest.mark.parametrize("method", ["tr_bdf2", "esdirk34"])
@pytest.mark.parametrize("model,tol", [("robertson_ode", "ROBERTSON_ODE_TOL"), ("robertson_dae", "ROBERTSON_DAE_TOL")])
def test_robertson_sweep_bit_exact(dsb, oracle, model, tol, method):
    from diffsol_b200 import sweeps
    B = 1500
    tolkw = getattr(sweeps, tol)
    p = sweeps.robertson_sweep(np.arange(B))
    solver = _solver(dsb, model, method, p, **tolkw)
    ys = solver.solve_dense(sweeps.ROBERTSON_T_EVAL)
    desc = oracle.make_desc(model, method=method, powmode=1, **tolkw)
    ys_o, stats_o, status_o = oracle.batch_solve_dense(desc, p, sweeps.ROBERTSON_T_EVAL)
    assert np.array_equal(solver.status(), status_o)
    assert np.array_equal(solver.statistics_array()[:, :13], stats_o[:, :13])
    assert np.array_equal(ys, ys_o, equal_nan=True)


@pytest.mark.parametrize("method", ["tr_bdf2", "esdirk34", "bdf"])
def test_van_der_pol_sweep_bit_exact(dsb, oracle, method):
    """BASELINE config 3 (mu in [1, 1e6], per-instance end time folded into scaled time) at a size the
    oracle finishes in seconds.  Many of the stiffest instances exhaust the reference's budget of 50
    Newton failures: their status codes and partial outputs must match too."""
    from diffsol_b200 import sweeps
    B = 2000
    p = sweeps.van_der_pol_scaled_sweep(np.arange(B))
    solver = _solver(dsb, "van_der_pol_scaled", method, p, **sweeps.VAN_DER_POL_TOL)
    ys = solver.solve_dense(sweeps.VAN_DER_POL_T_EVAL)
    desc = oracle.make_desc("van_der_pol_scaled", method=method, powmode=1, **sweeps.VAN_DER_POL_TOL)
    ys_o, stats_o, status_o = oracle.batch_solve_dense(desc, p, sweeps.VAN_DER_POL_T_EVAL)
    assert np.array_equal(solver.status(), status_o)
    assert (status_o == 0).sum() > 100
    assert np.array_equal(solver.statistics_array()[:, :13], stats_o[:, :13])
    assert np.array_equal(ys, ys_o, equal_nan=True)


def test_sdirk_edge_cases(dsb, oracle):
    from diffsol_b200 import sweeps
    tolkw = sweeps.ROBERTSON_ODE_TOL
    desc = oracle.make_desc("robertson_ode", method="tr_bdf2", powmode=1, **tolkw)
    solver = _solver(dsb, "robertson_ode", "tr_bdf2", [0.04, 1e4, 3e7], **tolkw)
    for t_eval in ([40.0], [0.0, 1.0, 1.0, 2.5], [1e-9, 5.0]):
        ys = solver.solve_dense(t_eval)
        rc, ys_o, st_o, _ = oracle.solve_dense(desc, [0.04, 1e4, 3e7], t_eval)
        assert rc == 0 and np.array_equal(ys[0], ys_o) and solver.get_statistics(0) == st_o
    ys = solver.solve_dense([0.0])
    assert solver.status()[0] == 5 and np.isnan(ys).all()
