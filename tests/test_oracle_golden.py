"""Pins the CPU oracle to the reference's own golden vectors (CPU only).

Every integer of `OdeSolverStatistics` / `OpStatistics` that the reference's inline insta snapshots
hold for the hot path (tests/golden/reference_snapshots.json, transcribed from
crates/diffsol/src/ode_solver/{bdf,sdirk}.rs) must be reproduced exactly, through the same harness
loop as the reference's `test_ode_solver` (ode_solver/mod.rs:104-194), and the states must pass the
reference's acceptance test sqrt(||y - y*||^2_w) < 20 (mod.rs:164-173) against the same solution
tables (SUNDIALS IDA/CVODE printouts, analytic solutions).
"""
import json
import math
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "reference_snapshots.json")) as f:
    GOLD = json.load(f)


def solution_points(kind):
    """The `OdeSolverSolution` tables of the reference's test models."""
    if kind == "exp_decay":            # exponential_decay.rs:264-286: y0 * exp(-k t), t = 0..9
        t = np.arange(10.0)
        return t, np.stack([np.exp(-0.1 * t)] * 2, axis=1)
    if kind == "exp_decay_algebraic":  # exponential_decay_with_algebraic.rs:298-304: t = 0, 0.1, .. 0.9
        t = np.arange(10) / 10.0
        return t, np.stack([np.exp(-0.1 * t)] * 3, axis=1)
    if kind == "robertson_dae":
        g = GOLD["robertson_dae_points"]
        return np.array(g["t"]), np.array(g["y"])
    if kind in ("robertson_ode", "robertson_ode_g3"):
        g = GOLD["robertson_ode_points"]
        y = np.array(g["y"])
        return np.array(g["t"]), (np.tile(y, (1, 3)) if kind.endswith("g3") else y)
    if kind == "dydt_y2":              # dydt_y2.rs:38-50: y0 / (1 - y0 t), t = 0, 2, .. 20
        t = np.arange(11) * 2.0
        return t, np.stack([-200.0 / (1.0 + 200.0 * t)] * 10, axis=1)
    if kind == "gaussian_decay":       # gaussian_decay.rs:41-53: exp(-a t^2 / 2), t = 0..9
        t = np.arange(10.0)
        return t, np.stack([np.exp(0.1 * t * t / -2.0)] * 10, axis=1)
    raise KeyError(kind)


def run_case(orc, case, powmode):
    t, ystar = solution_points(case["points"])
    desc = orc.make_desc(case["model"], method=case["method"], rtol=case["rtol"], atol=case["atol"],
                         use_coloring=case["coloring"], powmode=powmode)
    rc, ys, stats, fin = orc.harness(desc, case["p"], t)
    return rc, t, ys, ystar, stats


def expected_stats(case):
    s, r = case["setups"], case["rhs"]
    return {
        "number_of_linear_solver_setups": s[0],
        "number_of_linear_solver_setups_from_checkpoint": s[1],
        "number_of_linear_solver_setups_from_first_convergence_fail": s[2],
        "number_of_linear_solver_setups_from_second_convergence_fail": s[3],
        "number_of_linear_solver_setups_from_error_test_fail": s[4],
        "number_of_linear_solver_setups_from_step_success": s[5],
        "number_of_steps": case["steps"],
        "number_of_error_test_failures": case["etf"],
        "number_of_nonlinear_solver_iterations": case["nli"],
        "number_of_nonlinear_solver_fails": case["nlf"],
        "rhs_number_of_calls": r[0],
        "rhs_number_of_jac_muls": r[1],
        "rhs_number_of_matrix_evals": r[2],
    }


@pytest.mark.parametrize("powmode", [0, 1], ids=["libm_pow", "dsb_pow"])
@pytest.mark.parametrize("case", GOLD["cases"], ids=[c["name"] for c in GOLD["cases"]])
def test_reference_snapshot(oracle, case, powmode):
    rc, t, ys, ystar, stats = run_case(oracle, case, powmode)
    if rc == 10 and case["method"] != "bdf":
        pytest.skip("SDIRK restatement not built into the oracle yet")
    assert rc == 0
    assert stats == expected_stats(case), case["cite"]
    # the reference's acceptance test on the state (ode_solver/mod.rs:164-173)
    n = ystar.shape[1]
    atol = np.array(case["atol"] * n if len(case["atol"]) == 1 else case["atol"])
    for k in range(len(t)):
        w = np.abs(ystar[k]) * case["rtol"] + atol
        err = math.sqrt(np.mean(((ys[k] - ystar[k]) / w) ** 2))
        assert err < 20.0, (case["name"], t[k], ys[k], ystar[k])


def test_robertson_ode_single_group_equals_three_groups(oracle):
    """robertson_ode(ngroups=3) is three decoupled copies: the controller trace of the n=3 problem
    (BASELINE config 2's model) is pinned by the same snapshot (bdf.rs:2299-2321), jac_muls / 3."""
    case = dict(next(c for c in GOLD["cases"] if c["name"] == "bdf_robertson_ode_g3"))
    case.update(model="robertson_ode", atol=case["atol"][:3], points="robertson_ode")
    rc, t, ys, ystar, stats = run_case(oracle, case, 0)
    exp = expected_stats(case)
    exp["rhs_number_of_jac_muls"] = 27
    assert rc == 0 and stats == exp


def test_squared_norm_known_answer(oracle):
    """vector/mod.rs:530-541 test_squared_norm: ((1/(1*.1+.1))^2 + (2/(2*.1+.2))^2 + (3/(3*.1+.3))^2)/3 = 25."""
    import ctypes
    dp = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
    x = np.array([1.0, 2.0, 3.0]); y = np.array([1.0, 2.0, 3.0]); atol = np.array([0.1, 0.2, 0.3])
    r = oracle.lib().orc_squared_norm(dp(x), dp(y), dp(atol), 0.1, 3)
    assert abs(r - 25.0) < 1e-12


def test_lu_known_answer(oracle):
    """linear_solver/nalgebra/lu.rs:66-84: diag(2, 2) x = [2, 4] -> [1, 2] (tol 1e-10)."""
    import ctypes
    dp = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
    A = np.array([2.0, 0.0, 0.0, 2.0]); b = np.array([2.0, 4.0])
    assert oracle.lib().orc_lu_solve(dp(A), 2, dp(b)) == 0
    assert np.allclose(b, [1.0, 2.0], atol=1e-10)
    # a random system against numpy
    rng = np.random.default_rng(0)
    for n in (3, 5, 9):
        A = rng.standard_normal((n, n)); x = rng.standard_normal(n)
        b = A @ x
        Af = np.asfortranarray(A).ravel(order="F").copy()
        assert oracle.lib().orc_lu_solve(dp(Af), n, dp(b)) == 0
        assert np.allclose(b, x, rtol=1e-9, atol=1e-9)


def test_dsb_pow_accuracy(oracle):
    """The deterministic pow shared with the CUDA kernels stays within 1 ulp of libm pow on the
    ranges the controller uses, and powi is compiler-rt's square-and-multiply."""
    rng = np.random.default_rng(1)
    xs = np.concatenate([10.0 ** rng.uniform(-30, 30, 20000), rng.uniform(0, 2, 20000)])
    ys = np.concatenate([rng.uniform(-3, 3, 20000), rng.choice([0.8, 1.25, -0.5 / 3, -0.25, 1.0 / 3], 20000)])
    nbad = 0
    for x, y in zip(xs, ys):
        a, b = oracle.lib().orc_pow(x, y, 0), oracle.lib().orc_pow(x, y, 1)
        if a != b:
            nbad += 1
            assert abs(a - b) <= np.spacing(abs(a)) * 1.0000001, (x, y, a, b)
    assert nbad / len(xs) < 0.03
    assert oracle.lib().orc_powi(0.5, 9) == 0.5 ** 9
    assert oracle.lib().orc_powi(3.0, 0) == 1.0


@pytest.mark.parametrize("powmode", [0, 1], ids=["libm_pow", "dsb_pow"])
def test_reference_snapshot_heat2d(oracle, powmode):
    """bdf.rs:2424-2446: the 2-D heat equation DAE on a 10 x 10 grid (n = 100, boundary rows algebraic, Jacobian
    colouring) -- the reference's only statistics snapshot with n > 16.  It was taken with faer's sparse LU; the dense
    partial-pivoting restatement reproduces all 13 integers, and the output function (dx ||u||_2)^2 passes the
    reference's acceptance test against the SUNDIALS idaHeat2D table."""
    case = GOLD["heat2d_10"]
    desc = oracle.make_desc(case["model"], rtol=case["rtol"], atol=case["atol"], use_coloring=case["coloring"], powmode=powmode)
    rc, ys, stats, fin = oracle.harness(desc, case["p"], case["t"])
    assert rc == 0
    assert stats == expected_stats(case), case["cite"]
    out = (np.sqrt((ys ** 2).sum(axis=1)) / 9.0) ** 2           # heat2d_out (test_models/heat2d.rs:206-211)
    expected = np.array(case["out"])
    err = np.abs(out - expected) / (np.abs(expected) * case["out_rtol"] + case["out_atol"])
    assert err.max() < 20.0


@pytest.mark.parametrize("method,c,h,vec,F_expected", [("bdf", 0.1, 0.0, [1.1, 1.2], [2.11, 2.21]),
                                                        ("tr_bdf2", 0.1, 1.0, [1.1, 1.2], [1.12, 1.13])])
def test_residual_operator_known_answers(oracle, method, c, h, vec, F_expected):
    """test_bdf_callable (op/bdf.rs:317-360): F(y) = M (y - y0 + psi) - c f(y) = [2.11, 2.21], J = M - c f'(y) =
    diag(1.01), J v = [1.01, 1.01]; test_sdirk_callable (op/sdirk.rs:338-388): F(y) = M y - h f(phi + c y) =
    [1.12, 1.13], J = M - c h f'(phi + c y) = diag(1.01).  Both on the exponential decay problem (k = 0.1), evaluated
    with the member functions the oracle's step() uses."""
    desc = oracle.make_desc("exp_decay", method=method)
    rc, F, A = oracle.residual_known_answer(desc, [0.1, 1.0], c, h, vec, [1.0, 1.0])
    assert rc == 0
    assert np.abs(F - F_expected).max() < 1e-10
    assert A[0, 0] == 1.01 and A[1, 1] == 1.01 and A[0, 1] == 0.0 and A[1, 0] == 0.0      # assert_eq! in the reference
    assert np.abs(A @ np.ones(2) - [1.01, 1.01]).max() < 1e-10
