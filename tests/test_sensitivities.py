"""Forward sensitivities: `problem.bdf_sens::<LS>()` -- Bdf::sensitivity_solve
(/root/reference/crates/diffsol/src/ode_solver/bdf.rs:934-989), the sensitivity terms of the error test (:844-858, 908-919),
interpolate_sens (:1162-1215) and solve_dense_sensitivities (ode_solver/sensitivities.rs:205-262).

CPU tests pin the ORACLE: every counter of SEVEN of the reference's eight nalgebra sensitivity snapshots -- BDF: bdf_test_
nalgebra_exponential_decay_sens (bdf.rs:1812-1834), test_bdf_nalgebra_exponential_decay_algebraic_sens (:2118-2140),
test_bdf_nalgebra_robertson_sens (:2248-2271, 28 failed Newton solves); TR-BDF2 and ESDIRK34: exponential decay and the
Robertson DAE (sdirk.rs:708-730, 783-805, 895-918, 946-969) -- the reference's acceptance measures for the state (< 20) and
the sensitivities (< 29, ode_solver/mod.rs:164-187) against the analytic solution, and the kernel SOURCE (host emulation of
the DsbWithSens<M> instantiation of the on-chip BDF lane kernel) bit for bit against the oracle.
GPU tests: the CUDA path through the C ABI (dsb_batch_solve_dense_sensitivities_host) bit-identical to the oracle.

The reference's fourth BDF sensitivity snapshot (test_bdf_nalgebra_robertson_ode_sens, bdf.rs:2324-2345) is NOT reproduced
(851 steps here, 840 there): that run is chaotic at the level of one unit in the last place (a relative change of 1e-15 in
rtol moves it between 821 and 997 steps, while the same perturbation leaves the counters of the run WITHOUT sensitivities --
which the oracle does reproduce -- unchanged), so it cannot pin a restatement whose third-party arithmetic (nalgebra's gemm
/ LU evaluation order) is itself restated; test_robertson_ode_sens_snapshot_lies_within_the_one_ulp_scatter records that."""
import json
import math
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
with open(os.path.join(HERE, "golden", "reference_snapshots.json")) as f:
    GOLD = json.load(f)

# bdf.rs:1812-1834, in the order of oracle.S_NAMES
EXP_DECAY_SENS_SNAPSHOT = [14, 1, 0, 0, 1, 12, 56, 1, 175, 0, 60, 123, 2]
# bdf.rs:2324-2345
ROBERTSON_ODE_SENS_SNAPSHOT = [364, 1, 18, 0, 226, 119, 840, 226, 5099, 18, 1357, 3859, 28]


def exp_decay_sens_desc(oracle, powmode):
    # exponential_decay_problem_sens (test_models/exponential_decay.rs:703-742): p = [0.1, 1], sens_rtol 1e-6, sens_atol 1e-6
    return oracle.make_desc("exp_decay", sens=True, sens_rtol=1e-6, sens_atol=[1e-6, 1e-6], powmode=powmode)


@pytest.mark.parametrize("powmode", [0, 1], ids=["libm_pow", "dsb_pow"])
def test_exponential_decay_sens_snapshot(oracle, powmode):
    t = np.arange(10.0)
    rc, ys, sens, stats, fin = oracle.harness_sens(exp_decay_sens_desc(oracle, powmode), [0.1, 1.0], t)
    assert rc == 0
    assert list(stats.values())[:13] == EXP_DECAY_SENS_SNAPSHOT
    # the solution table of the test problem: y = y0 e^{-k t}, dy/dk = -t y0 e^{-k t}, dy/dy0 = e^{-k t}
    y = np.exp(-0.1 * t)
    want = [np.stack([y, y], axis=1), np.stack([-t * y] * 2, axis=1), np.stack([y, y], axis=1)]
    got = [ys, sens[:, 0, :], sens[:, 1, :]]
    for k in range(len(t)):
        for j, (g, w) in enumerate(zip(got, want)):
            err = math.sqrt(np.mean(((g[k] - w[k]) / (np.abs(w[k]) * 1e-6 + 1e-6)) ** 2))
            assert err < (20.0 if j == 0 else 29.0), (k, j, err)


# test_bdf_nalgebra_exponential_decay_algebraic_sens (bdf.rs:2118-2140) and test_bdf_nalgebra_robertson_sens (bdf.rs:2248-2271):
# DAEs, whose sensitivities are made consistent first (set_consistent_augmented, state.rs:167-238); the second runs with the
# sensitivities outside the error test and max_nonlinear_solver_failures = 70, and 28 of its steps fail inside a Newton solve
EXP_DECAY_ALGEBRAIC_SENS_SNAPSHOT = [24, 1, 0, 0, 8, 15, 45, 8, 115, 0, 66, 64, 3]
ROBERTSON_DAE_SENS_SNAPSHOT = [92, 1, 26, 2, 4, 59, 319, 4, 1941, 28, 575, 1522, 31]


@pytest.mark.parametrize("powmode", [0, 1], ids=["libm_pow", "dsb_pow"])
def test_dae_sens_snapshots(oracle, powmode):
    t = np.arange(10) / 10.0
    d = oracle.make_desc("exp_decay_algebraic", sens=True, sens_rtol=1e-6, sens_atol=[1e-6] * 3, powmode=powmode)
    rc, ys, sens, stats, fin = oracle.harness_sens(d, [0.1], t)
    assert rc == 0 and list(stats.values())[:13] == EXP_DECAY_ALGEBRAIC_SENS_SNAPSHOT
    y = np.exp(-0.1 * t)
    for k in range(len(t)):                          # the problem's solution table: y = e^{-k t}, dy/dk = -t e^{-k t} in every row
        for got, want, bound in ((ys[k], np.full(3, y[k]), 20.0), (sens[k, 0], np.full(3, -t[k] * y[k]), 29.0)):
            assert math.sqrt(np.mean(((got - want) / (np.abs(want) * 1e-6 + 1e-6)) ** 2)) < bound
    g = GOLD["robertson_dae_points"]
    t, ystar = np.array(g["t"]), np.array(g["y"])
    d = oracle.make_desc("robertson_dae", rtol=1e-4, atol=[1e-8, 1e-6, 1e-6], sens=True, powmode=powmode,
                         options=dict(max_nonlinear_solver_failures=70))
    rc, ys, sens, stats, fin = oracle.harness_sens(d, [0.04, 1e4, 3e7], t)
    assert rc == 0 and list(stats.values())[:13] == ROBERTSON_DAE_SENS_SNAPSHOT
    for k in range(len(t)):
        w = np.abs(ystar[k]) * 1e-4 + np.array([1e-8, 1e-6, 1e-6])
        assert math.sqrt(np.mean(((ys[k] - ystar[k]) / w) ** 2)) < 20.0
    assert np.abs(sens.sum(axis=-1)).max() < 1e-6 * np.abs(sens).max()      # the constraint y1 + y2 + y3 = 1, differentiated


# (E)SDIRK (Rk::do_stage_sdirk's sensitivity part, runge_kutta.rs:691-745; the sensitivity error, :812-822): sdirk.rs:708-730,
# 783-805 (exponential decay, sensitivities in the error test) and :895-918, 946-969 (Robertson DAE, 30 resp. 10 failed solves)
SDIRK_SENS_SNAPSHOTS = {
    ("tr_bdf2", "exp_decay"): [10, 1, 0, 0, 0, 9, 90, 0, 620, 0, 207, 421, 2],
    ("esdirk34", "exp_decay"): [6, 1, 0, 0, 0, 5, 33, 0, 347, 0, 107, 246, 1],
    ("tr_bdf2", "robertson_dae"): [77, 1, 29, 1, 0, 46, 286, 0, 4146, 30, 1303, 2954, 34],
    ("esdirk34", "robertson_dae"): [68, 1, 8, 2, 0, 57, 333, 0, 6856, 10, 2272, 4644, 17],
}


def sdirk_sens_case(oracle, method, model, powmode):
    if model == "exp_decay":
        d = oracle.make_desc("exp_decay", method=method, sens=True, sens_rtol=1e-6, sens_atol=[1e-6, 1e-6], powmode=powmode)
        return d, [0.1, 1.0], np.arange(10.0)
    opts = dict(max_nonlinear_solver_iterations=10) if method == "tr_bdf2" else None
    d = oracle.make_desc("robertson_dae", method=method, rtol=1e-4, atol=[1e-8, 1e-6, 1e-6], sens=True, powmode=powmode, options=opts)
    return d, [0.04, 1e4, 3e7], np.array(GOLD["robertson_dae_points"]["t"])


@pytest.mark.parametrize("powmode", [0, 1], ids=["libm_pow", "dsb_pow"])
@pytest.mark.parametrize("method,model", sorted(SDIRK_SENS_SNAPSHOTS))
def test_sdirk_sens_snapshots(oracle, method, model, powmode):
    d, p, t = sdirk_sens_case(oracle, method, model, powmode)
    rc, ys, sens, stats, fin = oracle.harness_sens(d, p, t)
    assert rc == 0 and list(stats.values())[:13] == SDIRK_SENS_SNAPSHOTS[(method, model)]
    if model == "exp_decay":
        y = np.exp(-0.1 * t)
        for k in range(len(t)):
            for got, want, bound in ((ys[k], np.full(2, y[k]), 20.0), (sens[k, 0], np.full(2, -t[k] * y[k]), 29.0), (sens[k, 1], np.full(2, y[k]), 29.0)):
                assert math.sqrt(np.mean(((got - want) / (np.abs(want) * 1e-6 + 1e-6)) ** 2)) < bound


def test_sensitivities_without_error_control_follow_the_plain_run(oracle):
    """turn_off_sensitivities_error_control: the sensitivities are integrated beside the state but stay out of the error test.
    The sensitivity residual's c is 0 until the first step-size update (op/bdf.rs:61, bdf.rs:551-553), so the step sequence
    is the plain run's.  The first step's sensitivities are then solved with c = 0 -- i.e. held at their predictor -- and
    nothing rejects that step, so an error of the size of that first step (1e-3 here) stays in d y / d k: the reference's
    behaviour, restated as it is."""
    t = np.arange(10.0)
    rc0, ys0, stats0, fin0 = oracle.harness(oracle.make_desc("exp_decay"), [0.1, 1.0], t)
    rc, ys, sens, stats, fin = oracle.harness_sens(oracle.make_desc("exp_decay", sens=True), [0.1, 1.0], t)
    assert rc == 0 and rc0 == 0
    assert stats["number_of_steps"] == stats0["number_of_steps"] and np.array_equal(ys, ys0)
    assert 1e-4 < np.abs(sens[:, 0, 0] + t * np.exp(-0.1 * t)).max() < 2e-3


def test_robertson_ode_sens_snapshot_lies_within_the_one_ulp_scatter(oracle):
    g = GOLD["robertson_ode_points"]
    t, ystar = np.array(g["t"]), np.array(g["y"])
    rows = []
    for eps in (0.0, 1e-15, 1e-14, 1e-13, 1e-12):
        d = oracle.make_desc("robertson_ode", rtol=1e-4 * (1.0 + eps), atol=[1e-8, 1e-6, 1e-6], sens=True, sens_rtol=1e-6, sens_atol=[1e-6] * 3)
        rc, ys, sens, stats, fin = oracle.harness_sens(d, [0.04, 1e4, 3e7], t)
        assert rc == 0
        rows.append(list(stats.values())[:13])
        if eps == 0.0:                       # the state passes the reference's acceptance test (the table has no sensitivities)
            for k in range(len(t)):
                w = np.abs(ystar[k]) * 1e-4 + np.array([1e-8, 1e-6, 1e-6])
                assert math.sqrt(np.mean(((ys[k] - ystar[k]) / w) ** 2)) < 20.0
    rows = np.array(rows)
    assert len({tuple(r) for r in rows}) >= 4                   # chaotic: (nearly) every perturbation gives another trajectory
    for name, idx in (("setups", 0), ("steps", 6), ("error test failures", 7), ("newton iterations", 8), ("rhs calls", 10), ("jac_muls", 11)):
        assert rows[:, idx].min() <= ROBERTSON_ODE_SENS_SNAPSHOT[idx] <= rows[:, idx].max(), name
    # the same perturbations leave the run WITHOUT sensitivities (reproduced exactly, tests/test_oracle_golden.py) unchanged
    plain = []
    for eps in (0.0, 1e-15, 1e-13):
        rc, ys, stats, fin = oracle.harness(oracle.make_desc("robertson_ode", rtol=1e-4 * (1.0 + eps), atol=[1e-8, 1e-6, 1e-6]), [0.04, 1e4, 3e7], t)
        plain.append(tuple(stats.values()))
    assert len(set(plain)) == 1


# ---- the kernel source on the host (no GPU) ---------------------------------------------------------------------------------
def exp_sweep(B):
    from diffsol_b200 import sweeps
    i = np.arange(B)
    return np.stack([0.02 * 100.0 ** sweeps.uniform(i, 0), 0.5 + 1.5 * sweeps.uniform(i, 1)], axis=1)


def robertson_sweep(B):
    from diffsol_b200 import sweeps
    return sweeps.robertson_sweep(np.arange(B))


@pytest.mark.parametrize("sens_atol", [[1e-6, 1e-6], None], ids=["error_control", "no_error_control"])
@pytest.mark.parametrize("free_running", [False, True], ids=["solve_dense", "step_loop"])
def test_kernel_source_equals_oracle_exponential_decay(oracle, sens_atol, free_running):
    from host_emu import emu
    p = exp_sweep(48)
    t = np.linspace(0.5, 10.0, 20)
    d = oracle.make_desc("exp_decay", sens=True, sens_rtol=1e-6 if sens_atol else None, sens_atol=sens_atol, powmode=1)
    if free_running:
        res = [oracle.harness_sens(d, p[k], t) for k in range(len(p))]
        assert all(r[0] == 0 for r in res)
        ys, se = np.array([r[1] for r in res]), np.array([r[2] for r in res])
        st = np.array([list(r[3].values())[:13] for r in res])
    else:
        ys, se, st, status = oracle.batch_solve_dense_sens(d, p, t)
        assert (status == 0).all()
        st = st[:, :13]
    r = emu.solve_sens(0, 2, 2, p, t, sens_rtol=1e-6 if sens_atol else None, sens_atol=sens_atol, free_running=free_running)
    assert (r["status"] == 0).all()
    assert np.array_equal(r["stats"][:, :13], st)
    assert np.array_equal(r["ys"], ys) and np.array_equal(r["sens"], se)


def test_kernel_source_equals_oracle_robertson(oracle):
    """Newton failures inside the sensitivity solves, Jacobian re-evaluations, orders up to 5 -- on the run that amplifies a
    one-ulp difference into other counters, the kernel source and the oracle agree bit for bit."""
    from host_emu import emu
    p = robertson_sweep(6)
    t = np.array([0.4, 4.0, 40.0, 400.0, 4000.0])
    tol = dict(rtol=1e-4, atol=[1e-8, 1e-6, 1e-6])
    ys, se, st, status = oracle.batch_solve_dense_sens(oracle.make_desc("robertson_ode", sens=True, sens_rtol=1e-6, sens_atol=[1e-6] * 3, powmode=1, **tol), p, t)
    r = emu.solve_sens(3, 3, 3, p, t, sens_rtol=1e-6, sens_atol=[1e-6] * 3, **tol)
    assert np.array_equal(r["status"], status) and (status == 0).all()
    assert np.array_equal(r["stats"][:, :13], st[:, :13]) and st[:, 9].sum() > 0          # Newton failures were exercised
    assert np.array_equal(r["ys"], ys) and np.array_equal(r["sens"], se)


def test_kernel_source_reproduces_the_dae_sens_snapshots(oracle):
    """The DAE form in the kernel: consistent sensitivities (the InitOp solve on SensRhs inside FETCH), the mass matrix in the
    sensitivity residual.  Through the free-running loop the kernel SOURCE gives every counter of bdf.rs:2118-2140 and
    bdf.rs:2248-2271 (28 failed Newton solves), and on a Robertson DAE sweep with the sensitivities in the error test it equals
    the oracle bit for bit."""
    from host_emu import emu
    t = np.array(GOLD["robertson_dae_points"]["t"])
    r = emu.solve_sens(2, 3, 3, np.array([[0.04, 1e4, 3e7]]), t, rtol=1e-4, atol=[1e-8, 1e-6, 1e-6], free_running=True,
                       options=dict(max_nonlinear_solver_failures=70))
    assert r["status"][0] == 0 and r["stats"][0, :13].tolist() == ROBERTSON_DAE_SENS_SNAPSHOT
    r = emu.solve_sens(1, 3, 1, np.array([[0.1]]), np.arange(10) / 10.0, sens_rtol=1e-6, sens_atol=[1e-6] * 3, free_running=True)
    assert r["status"][0] == 0 and r["stats"][0, :13].tolist() == EXP_DECAY_ALGEBRAIC_SENS_SNAPSHOT
    p = robertson_sweep(6)
    te = np.array([0.4, 4.0, 40.0, 400.0])
    tol = dict(rtol=1e-4, atol=[1e-8, 1e-6, 1e-6])
    ys, se, st, status = oracle.batch_solve_dense_sens(
        oracle.make_desc("robertson_dae", sens=True, sens_rtol=1e-5, sens_atol=[1e-7] * 3, powmode=1, **tol), p, te)
    r = emu.solve_sens(2, 3, 3, p, te, sens_rtol=1e-5, sens_atol=[1e-7] * 3, **tol)
    assert np.array_equal(r["status"], status) and (status == 0).all() and np.array_equal(r["stats"][:, :13], st[:, :13])
    assert np.array_equal(r["ys"], ys) and np.array_equal(r["sens"], se)


@pytest.mark.parametrize("method", ["tr_bdf2", "esdirk34"])
def test_sdirk_kernel_source_reproduces_the_sens_snapshots(oracle, method):
    """The (E)SDIRK lane kernel's sensitivity instantiation on the host: every counter of the reference's four TR-BDF2 /
    ESDIRK34 sensitivity snapshots through the free-running loop, and bit-identical to the oracle on Robertson ODE and DAE
    sweeps with the sensitivities in the error test."""
    from host_emu import emu
    r = emu.solve_sens(0, 2, 2, np.array([[0.1, 1.0]]), np.arange(10.0), method=method, sens_rtol=1e-6, sens_atol=[1e-6, 1e-6], free_running=True)
    assert r["status"][0] == 0 and r["stats"][0, :13].tolist() == SDIRK_SENS_SNAPSHOTS[(method, "exp_decay")]
    opts = dict(max_nonlinear_solver_iterations=10) if method == "tr_bdf2" else None
    r = emu.solve_sens(2, 3, 3, np.array([[0.04, 1e4, 3e7]]), np.array(GOLD["robertson_dae_points"]["t"]), method=method, rtol=1e-4,
                       atol=[1e-8, 1e-6, 1e-6], free_running=True, options=opts)
    assert r["status"][0] == 0 and r["stats"][0, :13].tolist() == SDIRK_SENS_SNAPSHOTS[(method, "robertson_dae")]
    p = robertson_sweep(6)
    te = np.array([0.4, 4.0, 40.0, 400.0])
    tol = dict(rtol=1e-4, atol=[1e-8, 1e-6, 1e-6])
    # (TR-BDF2 with the sensitivities in the error test takes millions of steps on this problem -- see the GPU test; short horizon)
    if method == "tr_bdf2":
        te = np.array([0.01, 0.02, 0.04])
    for model, model_id in (("robertson_ode", 3), ("robertson_dae", 2)):
        ys, se, st, status = oracle.batch_solve_dense_sens(
            oracle.make_desc(model, method=method, sens=True, sens_rtol=1e-5, sens_atol=[1e-7] * 3, powmode=1, **tol), p, te)
        r = emu.solve_sens(model_id, 3, 3, p, te, method=method, sens_rtol=1e-5, sens_atol=[1e-7] * 3, **tol)
        assert np.array_equal(r["status"], status) and (status == 0).all() and np.array_equal(r["stats"][:, :13], st[:, :13])
        assert np.array_equal(r["ys"], ys) and np.array_equal(r["sens"], se)


def test_sensitivity_arguments_are_checked_without_a_gpu():
    """dsb_problem_set_sensitivities: argument errors come back as DSB_BAD_ARG with a message (no device needed)."""
    import ctypes
    from diffsol_b200 import capi
    L = capi.lib()
    h = ctypes.c_void_p()
    capi.check(L.dsb_problem_new(capi.MODELS["exp_decay"], ctypes.byref(h)))
    a = np.array([1e-6, 1e-6, 1e-6])
    ptr = ctypes.c_void_p(a.ctypes.data)
    try:
        assert L.dsb_problem_set_sensitivities(h, 1, 1e-6, ptr, 3) != 0 and b"sens_atol" in L.dsb_last_error()
        assert L.dsb_problem_set_sensitivities(h, 1, 1e-6, None, 2) != 0
        assert L.dsb_problem_set_sensitivities(h, 1, -1.0, ptr, 2) != 0
        assert L.dsb_problem_set_sensitivities(h, 1, 1e-6, ptr, 2) == 0
        assert L.dsb_problem_set_sensitivities(h, 1, 0.0, None, 0) == 0
        assert L.dsb_problem_set_sensitivities(h, 0, 0.0, None, 0) == 0
        assert L.dsb_problem_set_sensitivities(None, 1, 0.0, None, 0) != 0
    finally:
        L.dsb_problem_free(h)


# ---- the CUDA path -------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def dsb():
    import diffsol_b200
    from diffsol_b200 import capi
    capi.require_device()
    return diffsol_b200


@pytest.mark.gpu
def test_gpu_reference_snapshot_exponential_decay_sens(dsb):
    """The reference's own test on the GPU: the step / interpolate / interpolate_sens loop over t = 0 .. 9 at p = [0.1, 1]
    gives every counter of bdf.rs:1812-1834 (sparsity probes excluded as everywhere: no colouring here)."""
    solver = dsb.OdeBuilder().rhs_implicit("exp_decay").p(np.array([[0.1, 1.0]] * 64)).sens_rtol(1e-6).sens_atol([1e-6, 1e-6]).build().bdf_sens()
    t = np.arange(10.0)
    ys, sens = solver.solve_dense_sensitivities(t, free_running=True)
    assert (solver.status() == 0).all()
    assert (solver.statistics_array()[:, :13] == np.array(EXP_DECAY_SENS_SNAPSHOT)).all()
    y = np.exp(-0.1 * t)
    assert np.abs(ys[:, :, 0] - y).max() < 1e-5 and np.abs(sens[:, :, 0, 0] + t * y).max() < 1e-5 and np.abs(sens[:, :, 1, 1] - y).max() < 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("sens_atol", [[1e-6, 1e-6], None], ids=["error_control", "no_error_control"])
@pytest.mark.parametrize("free_running", [False, True], ids=["solve_dense", "step_loop"])
def test_gpu_sensitivities_bit_exact_exponential_decay(dsb, oracle, sens_atol, free_running):
    B = 4000 if not free_running else 300
    p = exp_sweep(B)
    t = np.linspace(0.5, 10.0, 20)
    b = dsb.OdeBuilder().rhs_implicit("exp_decay").p(p)
    b = b.sens_rtol(1e-6).sens_atol(sens_atol) if sens_atol else b.sensitivities()
    solver = b.build().bdf_sens()
    ys, sens = solver.solve_dense_sensitivities(t, free_running=free_running)
    d = oracle.make_desc("exp_decay", sens=True, sens_rtol=1e-6 if sens_atol else None, sens_atol=sens_atol, powmode=1)
    if free_running:
        res = [oracle.harness_sens(d, p[k], t) for k in range(B)]
        ys_o, se_o = np.array([r[1] for r in res]), np.array([r[2] for r in res])
        st_o = np.array([list(r[3].values())[:13] for r in res])
        status_o = np.array([r[0] for r in res])
    else:
        ys_o, se_o, st_o, status_o = oracle.batch_solve_dense_sens(d, p, t)
        st_o = st_o[:, :13]
    assert np.array_equal(solver.status(), status_o) and (status_o == 0).all()
    assert np.array_equal(solver.statistics_array()[:, :13], st_o)
    assert np.array_equal(ys, ys_o) and np.array_equal(sens, se_o)
    # and the analytic sensitivities
    ex = p[:, 1, None] * np.exp(-p[:, 0, None] * t[None, :])
    if sens_atol:
        assert np.abs(sens[:, :, 0, 0] + t[None, :] * ex).max() < 2e-4 and np.abs(sens[:, :, 1, 0] - ex / p[:, 1, None]).max() < 2e-4


@pytest.mark.gpu
def test_gpu_sensitivities_bit_exact_robertson(dsb, oracle):
    """The stiff case: Newton failures inside sensitivity solves, hundreds of rejected steps, orders up to 5."""
    B = 2000
    p = robertson_sweep(B)
    t = np.array([0.4, 4.0, 40.0, 400.0, 4000.0])
    tol = dict(rtol=1e-4, atol=[1e-8, 1e-6, 1e-6])
    solver = (dsb.OdeBuilder().rhs_implicit("robertson_ode").p(p).rtol(tol["rtol"]).atol(tol["atol"])
              .sens_rtol(1e-6).sens_atol([1e-6] * 3).build().bdf_sens())
    ys, sens = solver.solve_dense_sensitivities(t)
    ys_o, se_o, st_o, status_o = oracle.batch_solve_dense_sens(
        oracle.make_desc("robertson_ode", sens=True, sens_rtol=1e-6, sens_atol=[1e-6] * 3, powmode=1, **tol), p, t)
    assert np.array_equal(solver.status(), status_o)
    assert np.array_equal(solver.statistics_array()[:, :13], st_o[:, :13]) and st_o[:, 9].sum() > 0
    assert np.array_equal(ys, ys_o, equal_nan=True) and np.array_equal(sens, se_o, equal_nan=True)
    # mass conservation differentiated: sum_i d y_i / d p_q = 0 for every parameter
    ok = status_o == 0
    assert ok.mean() > 0.99
    assert np.abs(sens[ok].sum(axis=-1)).max() < 1e-3 * max(1.0, np.abs(sens[ok]).max())


@pytest.mark.gpu
def test_gpu_reference_snapshots_dae_sens(dsb):
    """The reference's two DAE sensitivity tests on the GPU (free-running loop): every counter of bdf.rs:2118-2140 and
    bdf.rs:2248-2271."""
    t = np.array(GOLD["robertson_dae_points"]["t"])
    solver = (dsb.OdeBuilder().rhs_implicit("robertson_dae").p(np.array([[0.04, 1e4, 3e7]] * 40)).rtol(1e-4).atol([1e-8, 1e-6, 1e-6])
              .sensitivities().ode_options(max_nonlinear_solver_failures=70).build().bdf_sens())
    ys, sens = solver.solve_dense_sensitivities(t, free_running=True)
    assert (solver.status() == 0).all()
    assert (solver.statistics_array()[:, :13] == np.array(ROBERTSON_DAE_SENS_SNAPSHOT)).all()
    ystar = np.array(GOLD["robertson_dae_points"]["y"])
    assert (np.sqrt(np.mean(((ys[0] - ystar) / (np.abs(ystar) * 1e-4 + np.array([1e-8, 1e-6, 1e-6]))) ** 2, axis=1)) < 20.0).all()
    t2 = np.arange(10) / 10.0
    solver = dsb.OdeBuilder().rhs_implicit("exp_decay_algebraic").p(np.array([[0.1]] * 40)).sens_rtol(1e-6).sens_atol([1e-6] * 3).build().bdf_sens()
    ys, sens = solver.solve_dense_sensitivities(t2, free_running=True)
    assert (solver.status() == 0).all()
    assert (solver.statistics_array()[:, :13] == np.array(EXP_DECAY_ALGEBRAIC_SENS_SNAPSHOT)).all()
    assert np.abs(sens[:, :, 0, :] + (t2 * np.exp(-0.1 * t2))[None, :, None]).max() < 1e-5


@pytest.mark.gpu
def test_gpu_sensitivities_bit_exact_robertson_dae(dsb, oracle):
    B = 1500
    p = robertson_sweep(B)
    te = np.array([0.4, 4.0, 40.0, 400.0, 4000.0])
    tol = dict(rtol=1e-4, atol=[1e-8, 1e-6, 1e-6])
    solver = (dsb.OdeBuilder().rhs_implicit("robertson_dae").p(p).rtol(tol["rtol"]).atol(tol["atol"]).sens_rtol(1e-5).sens_atol([1e-7] * 3)
              .build().bdf_sens())
    ys, sens = solver.solve_dense_sensitivities(te)
    ys_o, se_o, st_o, status_o = oracle.batch_solve_dense_sens(
        oracle.make_desc("robertson_dae", sens=True, sens_rtol=1e-5, sens_atol=[1e-7] * 3, powmode=1, **tol), p, te)
    assert np.array_equal(solver.status(), status_o)
    assert np.array_equal(solver.statistics_array()[:, :13], st_o[:, :13])
    assert np.array_equal(ys, ys_o, equal_nan=True) and np.array_equal(sens, se_o, equal_nan=True)
    ok = status_o == 0
    assert ok.mean() > 0.99
    assert np.abs(sens[ok].sum(axis=-1)).max() < 1e-6 * max(1.0, np.abs(sens[ok]).max())      # y1 + y2 + y3 = 1, differentiated


@pytest.mark.gpu
@pytest.mark.parametrize("method", ["tr_bdf2", "esdirk34"])
def test_gpu_sdirk_sensitivities(dsb, oracle, method):
    """problem.tr_bdf2_sens() / esdirk34_sens(): the reference's four (E)SDIRK sensitivity snapshots on the GPU, and sweeps
    (exponential decay, Robertson ODE and DAE) bit-identical to the oracle."""
    prob = dsb.OdeBuilder().rhs_implicit("exp_decay").p(np.array([[0.1, 1.0]] * 40)).sens_rtol(1e-6).sens_atol([1e-6, 1e-6]).build()
    solver = getattr(prob, method + "_sens")()
    ys, sens = solver.solve_dense_sensitivities(np.arange(10.0), free_running=True)
    assert (solver.status() == 0).all()
    assert (solver.statistics_array()[:, :13] == np.array(SDIRK_SENS_SNAPSHOTS[(method, "exp_decay")])).all()
    b = (dsb.OdeBuilder().rhs_implicit("robertson_dae").p(np.array([[0.04, 1e4, 3e7]] * 40)).rtol(1e-4).atol([1e-8, 1e-6, 1e-6]).sensitivities())
    if method == "tr_bdf2":
        b = b.ode_options(max_nonlinear_solver_iterations=10)
    solver = getattr(b.build(), method + "_sens")()
    ys, sens = solver.solve_dense_sensitivities(np.array(GOLD["robertson_dae_points"]["t"]), free_running=True)
    assert (solver.status() == 0).all()
    assert (solver.statistics_array()[:, :13] == np.array(SDIRK_SENS_SNAPSHOTS[(method, "robertson_dae")])).all()
    te = np.array([0.4, 4.0, 40.0, 400.0, 4000.0])
    tol = dict(rtol=1e-4, atol=[1e-8, 1e-6, 1e-6])
    # TR-BDF2's sensitivity error estimate is not filtered through the iteration matrix (runge_kutta.rs:812-822), which on
    # this stiff problem drives the step size down to 2.6 MILLION steps per instance (restated as it is, measured on the
    # oracle); the reference's own Robertson test keeps the sensitivities out of the error test, and so does this sweep
    sens_tol = dict(sens_rtol=1e-5, sens_atol=[1e-7] * 3) if method == "esdirk34" else dict(sens_rtol=None, sens_atol=None)
    for model, B in (("robertson_ode", 1200), ("robertson_dae", 1200)):
        p = robertson_sweep(B)
        b = dsb.OdeBuilder().rhs_implicit(model).p(p).rtol(tol["rtol"]).atol(tol["atol"])
        b = b.sens_rtol(1e-5).sens_atol([1e-7] * 3) if method == "esdirk34" else b.sensitivities()
        solver = getattr(b.build(), method + "_sens")()
        ys, sens = solver.solve_dense_sensitivities(te)
        ys_o, se_o, st_o, status_o = oracle.batch_solve_dense_sens(
            oracle.make_desc(model, method=method, sens=True, powmode=1, **sens_tol, **tol), p, te)
        assert np.array_equal(solver.status(), status_o)
        assert np.array_equal(solver.statistics_array()[:, :13], st_o[:, :13])
        assert np.array_equal(ys, ys_o, equal_nan=True) and np.array_equal(sens, se_o, equal_nan=True)
    p = exp_sweep(2000)
    t = np.linspace(0.5, 10.0, 20)
    solver = getattr(dsb.OdeBuilder().rhs_implicit("exp_decay").p(p).sens_rtol(1e-6).sens_atol([1e-6, 1e-6]).build(), method + "_sens")()
    ys, sens = solver.solve_dense_sensitivities(t)
    ys_o, se_o, st_o, status_o = oracle.batch_solve_dense_sens(
        oracle.make_desc("exp_decay", method=method, sens=True, sens_rtol=1e-6, sens_atol=[1e-6, 1e-6], powmode=1), p, t)
    assert np.array_equal(solver.statistics_array()[:, :13], st_o[:, :13]) and np.array_equal(ys, ys_o) and np.array_equal(sens, se_o)
    ex = p[:, 1, None] * np.exp(-p[:, 0, None] * t[None, :])
    assert np.abs(sens[:, :, 0, 0] + t[None, :] * ex).max() < 2e-4


@pytest.mark.gpu
def test_gpu_sensitivities_against_finite_differences(dsb):
    """Independent of the oracle: d y / d p from the sensitivity equations against central differences of two plain solves."""
    B = 256
    p = robertson_sweep(B)
    t = np.array([0.4, 4.0, 40.0])
    def plain(pp):
        return dsb.OdeBuilder().rhs_implicit("robertson_ode").p(pp).rtol(1e-10).atol([1e-14, 1e-14, 1e-14]).build().bdf().solve_dense(t)
    solver = (dsb.OdeBuilder().rhs_implicit("robertson_ode").p(p).rtol(1e-8).atol([1e-12] * 3).sens_rtol(1e-8).sens_atol([1e-12] * 3)
              .build().bdf_sens())
    ys, sens = solver.solve_dense_sensitivities(t)
    assert (solver.status() == 0).all()
    for q in range(3):
        dp = np.zeros_like(p)
        dp[:, q] = 1e-3 * p[:, q]
        fd = (plain(p + dp) - plain(p - dp)) / (2.0 * dp[:, q])[:, None, None]
        scale = np.abs(fd).max(axis=(1, 2), keepdims=True)
        assert (np.abs(sens[:, :, q, :] - fd) / scale).max() < 5e-3, q     # the difference quotient of two 1e-10 solves is the limit


@pytest.mark.gpu
def test_gpu_sensitivity_errors(dsb):
    from diffsol_b200 import capi
    p = exp_sweep(8)
    t = np.array([1.0, 2.0])
    # an equation set without sens_mul / init_sens
    pv = np.array([[1.0]] * 4)
    with pytest.raises(capi.DiffsolB200Error, match="sensitivities"):
        dsb.OdeBuilder().rhs_implicit("van_der_pol").p(pv).sensitivities().build().bdf_sens().solve_dense_sensitivities(t)
    prob = dsb.OdeBuilder().rhs_implicit("exp_decay").p(p).sensitivities().build()
    # a problem with sensitivities goes through the sensitivity entry point, one without cannot
    with pytest.raises(capi.DiffsolB200Error, match="sensitivities"):
        prob.bdf().solve_dense(t)
    plain = dsb.OdeBuilder().rhs_implicit("exp_decay").p(p).build()
    with pytest.raises(ValueError):
        plain.bdf_sens()
    with pytest.raises(capi.DiffsolB200Error, match="sensitivities"):
        plain.bdf().solve_dense_sensitivities(t)


SENS_FUNCTOR = r"""
struct LogisticSens {
    static constexpr int N = 1, NP = 2;
    static constexpr bool HAS_MASS = false;
    static constexpr bool HAS_SENS = true;
    // x' = r x (1 - x / K), p = [r, K], x(0) = 0.1
    DSB_HD static void rhs(const double* x, const double* p, double, double* y) { y[0] = p[0] * x[0] * (1.0 - x[0] / p[1]); }
    DSB_HD static void jac_mul(const double* x, const double* p, double, const double* v, double* y) { y[0] = p[0] * (1.0 - 2.0 * x[0] / p[1]) * v[0]; }
    DSB_HD static void mass(const double* x, const double*, double, double beta, double* y) { y[0] = x[0] + beta * y[0]; }
    DSB_HD static void init(const double*, double, double* y) { y[0] = 0.1; }
    DSB_HD static void sens_mul(const double* x, const double* p, double, const double* v, double* y) {
        y[0] = x[0] * (1.0 - x[0] / p[1]) * v[0] + p[0] * x[0] * x[0] / (p[1] * p[1]) * v[1];
    }
    DSB_HD static void init_sens(const double*, double, const double*, double* y) { y[0] = 0.0; }
};
"""


@pytest.mark.gpu
def test_gpu_sensitivities_of_a_user_functor(dsb, oracle):
    """A closure-style functor with sens_mul / init_sens (builder.rs rhs_sens_implicit / init_sens) compiled at run time:
    bit-identical to the oracle (which compiled the same text), and the analytic d x / d r, d x / d K of logistic growth."""
    from diffsol_b200 import sweeps
    B = 1500
    i = np.arange(B)
    p = np.stack([0.5 + 2.0 * sweeps.uniform(i, 0), 0.8 + 0.7 * sweeps.uniform(i, 1)], axis=1)
    t = np.linspace(0.25, 4.0, 16)
    solver = (dsb.OdeBuilder().rhs_implicit_source(SENS_FUNCTOR, kind="functor", struct="LogisticSens").p(p).rtol(1e-8).atol(1e-10)
              .sens_rtol(1e-8).sens_atol(1e-10).build().bdf_sens())
    ys, sens = solver.solve_dense_sensitivities(t)
    name = oracle.load_user_model(SENS_FUNCTOR, kind="functor", struct="LogisticSens")
    ys_o, se_o, st_o, status_o = oracle.batch_solve_dense_sens(
        oracle.make_desc(name, rtol=1e-8, atol=1e-10, sens=True, sens_rtol=1e-8, sens_atol=1e-10, powmode=1), p, t)
    assert np.array_equal(solver.status(), status_o) and (status_o == 0).all()
    assert np.array_equal(solver.statistics_array()[:, :13], st_o[:, :13])
    assert np.array_equal(ys, ys_o) and np.array_equal(sens, se_o)
    r, K, x0 = p[:, 0, None], p[:, 1, None], 0.1
    A = (K - x0) / x0
    e = np.exp(-r * t[None, :])
    x = K / (1.0 + A * e)
    dx_dr = K * A * t[None, :] * e / (1.0 + A * e) ** 2
    dx_dK = 1.0 / (1.0 + A * e) - K * (e / x0) / (1.0 + A * e) ** 2
    assert np.abs(ys[:, :, 0] - x).max() < 1e-6
    assert np.abs(sens[:, :, 0, 0] - dx_dr).max() < 1e-5 and np.abs(sens[:, :, 1, 0] - dx_dK).max() < 1e-5


# the reference's external logistic module (crates/diffsol-c/tests/external-dynamic-logistic/src/lib.rs) with its forward-mode
# symbols: rhs_sgrad (:189-208) and set_u0_sgrad (:292-301, empty: u0 does not depend on the input), no out / stop functions
LOGISTIC_DIFFSL_SENS = r"""
#define DSB_DIFFSL_STATES 1
#define DSB_DIFFSL_INPUTS 1
#define DSB_DIFFSL_DATA 1
#define DSB_DIFFSL_HAS_SENS 1
DSB_SYMBOL void set_u0(double* u, double* data, unsigned thread_id, unsigned thread_dim) { if (u) *u = 0.1; }
DSB_SYMBOL void rhs(double t, const double* u, double* data, double* rr, unsigned thread_id, unsigned thread_dim) {
    if (!u || !data || !rr) return;
    const double x = *u, r = *data;
    *rr = r * x * (1.0 - x);
}
DSB_SYMBOL void rhs_grad(double t, const double* u, const double* du, const double* data, double* ddata, const double* rr,
                         double* drr, unsigned thread_id, unsigned thread_dim) {
    if (!u || !du || !data || !ddata || !drr) return;
    const double x = *u, dx = *du, r = *data;
    *drr = r * (1.0 - 2.0 * x) * dx;
    *ddata = x * (1.0 - x);
}
DSB_SYMBOL void rhs_sgrad(double t, const double* u, const double* data, double* ddata, const double* rr, double* drr,
                          unsigned thread_id, unsigned thread_dim) {
    if (!u || !data || !ddata || !drr) return;
    const double x = *u;
    *drr = x * (1.0 - x) * *ddata;              // the module of the reference hard-codes d r = 1; here the seeded direction
}
DSB_SYMBOL void set_u0_sgrad(const double* u, double* du, const double* data, double* ddata, unsigned thread_id, unsigned thread_dim) {}
DSB_SYMBOL void set_inputs(const double* inputs, double* data, unsigned model_index) { if (inputs && data) *data = *inputs; }
"""


@pytest.mark.gpu
def test_gpu_sensitivities_of_a_diffsl_symbol_table(dsb, oracle):
    """A DiffSL module with its forward-mode symbols (rhs_sgrad / set_u0_sgrad) as source text: DiffSlRhs::sens_mul_inplace and
    DiffSlInit::sens_mul_inplace (ode_equations/diffsl.rs:1152-1168, 737-750) through csrc/dsb_diffsl_adapter.h.  Bit-identical
    to the oracle (same text), and d x / d r of logistic growth."""
    from diffsol_b200 import sweeps
    B = 1200
    r = (0.5 + 2.0 * sweeps.uniform(np.arange(B), 0)).reshape(-1, 1)
    t = np.linspace(0.25, 4.0, 16)
    solver = (dsb.OdeBuilder().rhs_implicit_source(LOGISTIC_DIFFSL_SENS, kind="diffsl").p(r).rtol(1e-8).atol(1e-10)
              .sens_rtol(1e-8).sens_atol(1e-10).build().bdf_sens())
    ys, sens = solver.solve_dense_sensitivities(t)
    name = oracle.load_user_model(LOGISTIC_DIFFSL_SENS, kind="diffsl")
    ys_o, se_o, st_o, status_o = oracle.batch_solve_dense_sens(
        oracle.make_desc(name, rtol=1e-8, atol=1e-10, sens=True, sens_rtol=1e-8, sens_atol=1e-10, powmode=1), r, t)
    assert np.array_equal(solver.status(), status_o) and (status_o == 0).all()
    assert np.array_equal(solver.statistics_array()[:, :13], st_o[:, :13])
    assert np.array_equal(ys, ys_o) and np.array_equal(sens, se_o)
    A, e = 9.0, np.exp(-r * t[None, :])
    assert np.abs(sens[:, :, 0, 0] - A * t[None, :] * e / (1.0 + A * e) ** 2).max() < 1e-5


def test_oracle_sensitivities_of_a_diffsl_symbol_table(oracle):
    """The same module on the oracle (CPU): the analytic d x / d r."""
    name = oracle.load_user_model(LOGISTIC_DIFFSL_SENS, kind="diffsl")
    r = np.linspace(0.5, 2.5, 9).reshape(-1, 1)
    t = np.linspace(0.25, 4.0, 16)
    ys, se, st, status = oracle.batch_solve_dense_sens(
        oracle.make_desc(name, rtol=1e-8, atol=1e-10, sens=True, sens_rtol=1e-8, sens_atol=1e-10), r, t)
    assert (status == 0).all()
    A, e = 9.0, np.exp(-r * t[None, :])
    assert np.abs(ys[:, :, 0] - 1.0 / (1.0 + A * e)).max() < 1e-6
    assert np.abs(se[:, :, 0, 0] - A * t[None, :] * e / (1.0 + A * e) ** 2).max() < 1e-5
