"""The product's lane kernels, compiled for the HOST and run one lane at a time (tests/host_emu: the kernel source of
diffsol_b200/csrc/*.cuh behind a CUDA shim), against the oracle: counters, status and states bit-identical.

This checks the kernels' control flow and arithmetic on machines without a GPU; it does not replace the `-m gpu`
parity tests (those run the real sm_100a build through the C ABI, many lanes per warp).  The product itself has no
host integrator: nothing under diffsol_b200/ can reach this build."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from host_emu import emu  # noqa: E402

from diffsol_b200 import sweeps  # noqa: E402


def run_both(oracle, model, params, t_eval, method="bdf", kernel="lane", **kw):
    n, np_, _ = oracle.model_dims(model)
    desc = oracle.make_desc(model, method=method, powmode=1, **kw)
    ys_o, stats_o, status_o = oracle.batch_solve_dense(desc, params, t_eval)
    r = emu.solve(oracle.MODELS[model], n, np_, params, t_eval, method=method, kernel=kernel, **kw)
    return r, ys_o, stats_o, status_o


def assert_same(r, ys_o, stats_o, status_o):
    assert np.array_equal(r["status"], status_o)
    assert np.array_equal(r["stats"][:, :13], stats_o[:, :13])
    assert np.array_equal(r["ys"], ys_o, equal_nan=True)


def spm_currents(B):
    return (0.6 + 0.8 * sweeps.uniform(np.arange(B), 0)).reshape(-1, 1)


@pytest.mark.parametrize("model,tol,coloring", [
    ("robertson_ode", sweeps.ROBERTSON_ODE_TOL, False),
    ("robertson_ode", sweeps.ROBERTSON_ODE_TOL, True),
    ("robertson_dae", sweeps.ROBERTSON_DAE_TOL, False),
])
def test_bdf_lane_kernel_on_host(oracle, model, tol, coloring):
    p = sweeps.robertson_sweep(np.arange(48))
    r, *o = run_both(oracle, model, p, sweeps.ROBERTSON_T_EVAL, use_coloring=coloring, **tol)
    assert (o[2] == 0).all()
    assert_same(r, *o)


@pytest.mark.parametrize("method", ["tr_bdf2", "esdirk34"])
def test_sdirk_lane_kernel_on_host(oracle, method):
    p = sweeps.van_der_pol_scaled_sweep(np.arange(48))
    r, *o = run_both(oracle, "van_der_pol_scaled", p, sweeps.VAN_DER_POL_T_EVAL, method=method, **sweeps.VAN_DER_POL_TOL)
    assert_same(r, *o)


@pytest.mark.parametrize("model,B,coloring", [("spm", 12, False), ("spm", 12, True), ("spm99", 3, True)])
def test_band_bdf_lane_kernel_on_host(oracle, model, B, coloring):
    r, *o = run_both(oracle, model, spm_currents(B), np.arange(1, 7) * 600.0, kernel="band", use_coloring=coloring)
    assert (o[2] == 0).all()
    assert_same(r, *o)


def heat_params(B):
    i = np.arange(B)
    return np.stack([1.0 + sweeps.uniform(i, 0), 0.1 + 0.3 * sweeps.uniform(i, 1), 0.6 + 0.3 * sweeps.uniform(i, 2)], axis=1)


@pytest.mark.parametrize("model,B,coloring", [("heat1d_dae_32", 10, False), ("heat1d_dae_32", 10, True), ("heat1d_dae_256", 2, True)])
def test_band_dae_kernels_on_host(oracle, model, B, coloring):
    """Singular mass on the banded lane path: dsb_band_init_kernel (consistent initialisation) + the BDF kernel."""
    r, *o = run_both(oracle, model, heat_params(B), np.arange(1, 101) / 100.0 * 0.99, kernel="band", use_coloring=coloring,
                     rtol=1e-6, atol=1e-6)
    assert (o[2] == 0).all()
    assert_same(r, *o)


@pytest.mark.parametrize("method", ["tr_bdf2", "esdirk34"])
@pytest.mark.parametrize("model,B,coloring", [("spm", 8, False), ("spm", 8, True), ("spm99", 2, True),
                                               ("heat1d_dae_32", 8, False), ("heat1d_dae_256", 2, True)])
def test_band_sdirk_kernel_on_host(oracle, method, model, B, coloring):
    if model.startswith("spm"):
        p, t_eval = spm_currents(B), np.arange(1, 13) * 300.0
    else:
        p, t_eval = heat_params(B), np.arange(1, 101) / 100.0 * 0.99
    r, *o = run_both(oracle, model, p, t_eval, method=method, kernel="band", use_coloring=coloring, rtol=1e-6, atol=1e-6)
    assert (o[2] == 0).all()
    assert_same(r, *o)


def run_both_roots(oracle, model, params, t_eval, method="bdf", kernel="lane", **kw):
    n, np_, _ = oracle.model_dims(model)
    desc = oracle.make_desc(model, method=method, powmode=1, **kw)
    o = oracle.batch_solve_dense_roots(desc, params, t_eval)
    r = emu.solve(oracle.MODELS[model], n, np_, params, t_eval, method=method, kernel=kernel,
                  nout=oracle.model_nout(model), **kw)
    return r, o


def assert_same_roots(r, o):
    ys_o, stats_o, status_o, t_root_o, root_idx_o, ncols_o = o
    assert_same(r, ys_o, stats_o, status_o)
    assert np.array_equal(r["root_idx"], root_idx_o) and np.array_equal(r["ncols"], ncols_o)
    stopped = root_idx_o >= 0
    assert np.array_equal(r["fin"][stopped, 0], t_root_o[stopped])
    return stopped


@pytest.mark.parametrize("method", ["bdf", "tr_bdf2", "esdirk34"])
def test_roots_in_the_on_chip_lane_kernels_on_host(oracle, method):
    idx = np.arange(120)
    p = np.stack([0.02 * 50.0 ** sweeps.uniform(idx, 0), 0.4 + 1.6 * sweeps.uniform(idx, 1)], axis=1)
    r, o = run_both_roots(oracle, "exp_decay_root", p, np.arange(1.0, 21.0), method=method)
    stopped = assert_same_roots(r, o)
    assert 0 < stopped.sum() < len(idx)


@pytest.mark.parametrize("model,method,B", [("spm_stop", "bdf", 24), ("spm_stop", "tr_bdf2", 12), ("spm_stop", "esdirk34", 12),
                                             ("spm99_stop", "bdf", 3), ("spm99_stop", "tr_bdf2", 2)])
def test_battery_voltage_cut_off_in_the_band_kernels_on_host(oracle, model, method, B):
    r, o = run_both_roots(oracle, model, spm_currents(B), np.arange(1, 121) * 30.0, method=method, kernel="band",
                          use_coloring=True)
    stopped = assert_same_roots(r, o)
    assert stopped.sum() > 0


# ---- the warp-per-instance banded kernel (dsb_wband_bdf_kernel.cuh), built with one lane per warp ------------------------
@pytest.mark.parametrize("model,B,coloring", [("spm", 12, False), ("spm", 12, True), ("spm99", 3, True)])
def test_warp_band_bdf_kernel_on_host(oracle, model, B, coloring):
    r, *o = run_both(oracle, model, spm_currents(B), np.arange(1, 7) * 600.0, kernel="warp", use_coloring=coloring)
    assert (o[2] == 0).all()
    assert_same(r, *o)


@pytest.mark.parametrize("model,B,coloring", [("heat1d_dae_32", 10, False), ("heat1d_dae_32", 10, True), ("heat1d_dae_256", 2, True),
                                               ("heat1d_dae_32_bc", 6, True)])
def test_warp_band_dae_kernel_on_host(oracle, model, B, coloring):
    r, *o = run_both(oracle, model, heat_params(B), np.arange(1, 101) / 100.0 * 0.99, kernel="warp", use_coloring=coloring,
                     rtol=1e-6, atol=1e-6)
    assert (o[2] == 0).all()
    assert_same(r, *o)


@pytest.mark.parametrize("n,kl,ku", [(24, 1, 1), (24, 2, 1), (24, 1, 2), (24, 2, 2), (7, 2, 2), (41, 1, 1)])
def test_smem_band_lu_matches_the_dense_restatement(oracle, n, kl, ku):
    """The band LU of the warp-per-instance kernel (row-per-diagonal band storage, one-lane recurrences, reciprocal-reuse back
    substitution) against the oracle's dense nalgebra LU on random banded matrices that DO interchange rows, on diagonally
    dominant ones that do not (the interchange-free forward sweep), with right-hand sides that span the whole exponent range
    (the in-place plain division for numerators outside the fast quotient's proven range, signed zeros), and singular ones."""
    import ctypes
    rng = np.random.default_rng(1000 * n + 10 * kl + ku)
    dp = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
    seen_swaps = seen_noswaps = seen_few = 0
    for trial in range(300):
        A = np.zeros((n, n))
        for i in range(n):
            for j in range(max(0, i - kl), min(n, i + ku + 1)):
                A[i, j] = rng.standard_normal() * 10.0 ** rng.integers(-3, 4)
        if trial % 3 == 0:
            A[np.arange(n), np.arange(n)] += 1e5                      # no interchanges
        if trial % 3 == 1:                                            # only a few interchanges: the segmented forward sweep
            A[np.arange(n), np.arange(n)] += 1e5
            for r in rng.integers(0, n - 1, 3):
                A[r, r] = 1e-3
        b = rng.standard_normal(n) * 10.0 ** rng.integers(-8, 9, n)
        if trial % 4 == 1:
            b = b * 10.0 ** rng.choice([-310.0, -250.0, -160.0, 0.0, 120.0, 250.0, 290.0], n)
            b[rng.integers(0, n, 3)] = [0.0, -0.0, 5e-324]
        if trial % 50 == 49:
            A[:, n // 2] = 0.0                                          # a zero column: LuSolveFailed
        Af = np.asfortranarray(A)
        x_o = b.copy()
        with np.errstate(all="ignore"):
            rc_o = oracle.lib().orc_lu_solve(dp(Af), n, dp(x_o))
        for mode in (0, 1, 2):                                          # fast (forward sweep by interchange count), exact, fast with selects
            exact = mode == 1
            rc, x, nsw = emu.smem_band_lu(A, kl, ku, b, mode=mode)
            if rc_o != 0:
                assert rc == 0
                continue
            assert rc in (1, 2)
            if rc == 1:
                assert np.array_equal(x.view(np.int64), x_o.view(np.int64)), (trial, exact)
            assert not (exact and rc == 2)
            seen_swaps += nsw > 8
            seen_few += 0 < nsw <= 8
            seen_noswaps += nsw == 0
    assert (seen_swaps > 100 or n <= 8) and seen_noswaps > 100 and seen_few > 100


def test_warp_band_exact_solve_path_on_host(oracle, monkeypatch):
    """The back substitution's fall-back (right-hand side rebuilt, plain IEEE divisions) forced on every Newton iteration."""
    monkeypatch.setenv("DSB_WBAND_FORCE_REDO", "1")
    r, *o = run_both(oracle, "heat1d_dae_32", heat_params(6), np.arange(1, 101) / 100.0 * 0.99, kernel="warp", rtol=1e-6, atol=1e-6)
    assert_same(r, *o)
    r, *o = run_both(oracle, "spm", spm_currents(6), np.arange(1, 7) * 600.0, kernel="warp", use_coloring=True)
    assert_same(r, *o)


@pytest.mark.parametrize("model,B", [("spm_stop", 24), ("spm99_stop", 3)])
def test_battery_voltage_cut_off_in_the_warp_band_kernel_on_host(oracle, model, B):
    r, o = run_both_roots(oracle, model, spm_currents(B), np.arange(1, 121) * 30.0, method="bdf", kernel="warp", use_coloring=True)
    stopped = assert_same_roots(r, o)
    assert stopped.sum() > 0


@pytest.mark.parametrize("method", ["bdf", "tr_bdf2"])
@pytest.mark.parametrize("coloring", [False, True])
def test_band_dae_inconsistent_initial_values_on_host(oracle, method, coloring):
    """Boundary rows 0 = u - height / 4 with u = 0 initially: dsb_band_init_kernel has to move the algebraic components
    (InitOp Newton + backtracking line search on the band LU) before the integrator starts."""
    p = heat_params(8)
    t_eval = np.arange(1, 11) / 10.0 * 0.99
    r, *o = run_both(oracle, "heat1d_dae_32_bc", p, t_eval, method=method, kernel="band", use_coloring=coloring, rtol=1e-6, atol=1e-6)
    assert (o[2] == 0).all()
    assert_same(r, *o)
    assert np.abs(r["ys"][:, :, 0] - 0.25 * p[:, :1]).max() < 1e-12 and np.abs(r["ys"][:, :, -1] - 0.25 * p[:, :1]).max() < 1e-12


@pytest.mark.parametrize("method", ["bdf", "tr_bdf2", "esdirk34"])
@pytest.mark.parametrize("tol", [1e-6, 1e-9])
def test_reset_in_the_on_chip_lane_kernels_on_host(oracle, tol, method):
    """Equations with a reset function: a root does not end solve_dense (method.rs:783-797).  The BDF lane kernel applies
    the reset and re-initialises to first order (bdf.rs:1291-1318); the SDIRK lane kernel applies it and re-initialises
    the root finder and the stop time (Rk::start_step, runge_kutta.rs:446-464); both integrate on, many resets per
    instance."""
    idx = np.arange(100)
    p = np.stack([0.02 * 50.0 ** sweeps.uniform(idx, 0), 0.35 + 1.65 * sweeps.uniform(idx, 1)], axis=1)
    r, o = run_both_roots(oracle, "exp_decay_reset", p, np.arange(1.0, 41.0), method=method, rtol=tol, atol=[tol])
    assert_same_roots(r, o)
    assert (o[2] == 0).all() and (o[4] == -1).all()
    assert r["ys"][:, -1, 0].max() <= 0.6 + 1e-6          # nobody stays above the first root level


@pytest.mark.parametrize("method", ["bdf", "tr_bdf2", "esdirk34"])
def test_ball_bounce_in_the_on_chip_lane_kernels_on_host(oracle, method):
    """The reference's bouncing ball (ode_solver/mod.rs:1001-1080) swept over gravity, drop height and restitution:
    several bounces per instance, each one a root + reset inside the lane kernel; the window ends before the earliest
    accumulation point of bounces in the sweep (t_b (1 + e) / (1 - e) >= 4.6)."""
    idx = np.arange(60)
    p = np.stack([5.0 + 10.0 * sweeps.uniform(idx, 0), 2.0 + 18.0 * sweeps.uniform(idx, 1), 0.8 + 0.15 * sweeps.uniform(idx, 2)], axis=1)
    r, o = run_both_roots(oracle, "ball_bounce", p, np.linspace(0.05, 4.0, 80), method=method)
    assert_same_roots(r, o)
    assert (o[2] == 0).all() and (o[4] == -1).all()
    assert (o[5] >= 79).all()
    assert np.nanmin(r["ys"][:, :, 0]) > -1e-3             # the ball stays above the ground


@pytest.mark.parametrize("method", ["bdf", "tr_bdf2", "esdirk34"])
def test_backward_integration_on_host(oracle, method):
    """negative_exponential_decay_problem (test_models/exponential_decay.rs:168-197: h0 = -1, points 0, -1, .., -9)
    through the step()/interpolate() loop of the reference's harness, as bdf.rs:1729-1733 / sdirk.rs:669-673 run it."""
    k, y0 = 0.1, np.exp(-1.0)
    pts = -np.arange(0.0, 10.0)
    desc = oracle.make_desc("exp_decay", method=method, powmode=1, h0=-1.0)
    rc, ys_o, stats_o, fin = oracle.harness(desc, [k, y0], pts)
    assert rc == 0
    r = emu.solve(oracle.MODELS["exp_decay"], 2, 2, [[k, y0]], pts, method=method, h0=-1.0, free_running=True)
    assert r["status"][0] == 0 and np.array_equal(r["ys"][0], ys_o)
    assert {n: int(r["stats"][0, i]) for i, n in enumerate(oracle.S_NAMES)} == stats_o
    assert r["fin"][0, 0] == fin["t"] and r["fin"][0, 1] == fin["h"] and fin["h"] < 0
    exact = y0 * np.exp(-k * pts)
    err = np.sqrt(np.mean(((ys_o - exact[:, None]) / (np.abs(exact[:, None]) * 1e-6 + 1e-6)) ** 2, axis=1))
    assert err.max() < (30.0 if method == "esdirk34" else 20.0)      # the reference's acceptance thresholds


@pytest.mark.parametrize("method", ["bdf", "tr_bdf2", "esdirk34"])
def test_root_index_in_the_on_chip_lane_kernels_on_host(oracle, method):
    """exponential_decay_with_two_roots_problem (test_models/exponential_decay.rs:890-912) swept over rate and initial
    value: instances started above 0.6 stop on root 0, those between 0.3 and 0.6 on root 1, the rest run to the end."""
    idx = np.arange(120)
    p = np.stack([0.02 * 50.0 ** sweeps.uniform(idx, 0), 0.1 + 1.4 * sweeps.uniform(idx, 1)], axis=1)
    r, o = run_both_roots(oracle, "exp_decay_two_roots", p, np.arange(1.0, 21.0), method=method)
    assert_same_roots(r, o)
    fired = o[4]
    assert set(fired.tolist()) == {-1, 0, 1}
    assert (fired[p[:, 1] > 0.6] != 1).all() and (fired[p[:, 1] < 0.6] != 0).all() and (fired[p[:, 1] < 0.3] == -1).all()


@pytest.mark.parametrize("method", ["bdf", "tr_bdf2", "esdirk34"])
def test_reset_in_the_band_kernels_on_host(oracle, method):
    """Resets on the banded lane kernels: the battery model cycled (spm_cycle: at a voltage cut-off the cell goes back
    to its charged state and the discharge starts again), two to three discharges per instance."""
    B = 10
    r, o = run_both_roots(oracle, "spm_cycle", spm_currents(B), np.arange(1, 121) * 60.0, method=method, kernel="band",
                          use_coloring=True)
    assert_same_roots(r, o)
    assert (o[2] == 0).all() and (o[4] == -1).all()
    v = r["ys"][:, :, 0]
    assert (np.diff(v, axis=1) > 0.3).sum(axis=1).min() >= 1          # every instance was recharged at least once
    assert v.min() > 3.0 and v.max() < 4.2


@pytest.mark.parametrize("method", ["bdf", "tr_bdf2", "esdirk34"])
def test_harness_loop_stops_at_the_root_on_host(oracle, method):
    """The step()/interpolate() loop of the reference's harness (ode_solver/mod.rs:132-141) returns interpolate(t_root) for
    the point it was stepping towards when a step reports RootFound, and ends; without a root before the last point it
    runs through.  step_and_interpolate does the same."""
    pts = np.arange(0.0, 10.0)
    desc = oracle.make_desc("exp_decay_root", method=method, powmode=1)
    for p, stops in (([0.1, 1.0], True), ([0.01, 1.0], False)):
        rc, ys_o, stats_o, fin = oracle.harness(desc, p, pts)
        assert rc == 0
        r = emu.solve(oracle.MODELS["exp_decay_root"], 2, 2, [p], pts, method=method, free_running=True)
        assert r["status"][0] == 0 and np.array_equal(r["ys"][0], ys_o, equal_nan=True)
        assert {n: int(r["stats"][0, i]) for i, n in enumerate(oracle.S_NAMES)} == stats_o
        assert r["fin"][0, 0] == fin["t"]
        written = int(np.isfinite(ys_o[:, 0]).sum())
        assert (written < len(pts)) == stops and r["ncols"][0] == written and (r["root_idx"][0] >= 0) == stops


COLORING_CASES = [      # jacobian/mod.rs:483-513 build_coloring: (row, col) patterns of 2 x 2 operators and their colourings
    ([(0, 0), (1, 1)], [1, 1]),
    ([(0, 0), (0, 1), (1, 1)], [1, 2]),
    ([(1, 1)], [1, 1]),
    ([(0, 0), (1, 0), (0, 1), (1, 1)], [1, 2]),
]


@pytest.mark.parametrize("non_zeros,expected", COLORING_CASES)
def test_greedy_coloring_known_answers(oracle, non_zeros, expected):
    """nonzeros2graph + color_graph_greedy (jacobian/coloring.rs:27-47, greedy_coloring.rs:14-34) with the reference's own
    expected colourings, for the oracle's restatement and for the product's host-side one (csrc/dsb_host_setup.h)."""
    assert oracle.greedy_coloring(non_zeros, 2) == expected
    assert emu.greedy_coloring(non_zeros, 2) == expected


def test_greedy_coloring_band_patterns(oracle):
    """Tridiagonal and pentadiagonal patterns need 3 and 5 colours; an arrow pattern (dense first row) n; both restatements
    agree on every pattern."""
    rng = np.random.default_rng(3)
    for n, half in ((12, 1), (12, 2)):
        nz = [(i, j) for j in range(n) for i in range(n) if abs(i - j) <= half]
        assert max(oracle.greedy_coloring(nz, n)) == 2 * half + 1
        assert emu.greedy_coloring(nz, n) == oracle.greedy_coloring(nz, n)
    arrow = [(0, j) for j in range(8)] + [(j, j) for j in range(1, 8)]
    assert oracle.greedy_coloring(arrow, 8) == list(range(1, 9)) == emu.greedy_coloring(arrow, 8)
    for _ in range(20):
        n = int(rng.integers(2, 15))
        nz = sorted({(int(rng.integers(0, n)), int(rng.integers(0, n))) for _ in range(3 * n)}, key=lambda ij: (ij[1], ij[0]))
        a, b = oracle.greedy_coloring(nz, n), emu.greedy_coloring(nz, n)
        assert a == b
        for (i, j) in nz:                                  # structurally orthogonal: no two columns of a colour share a row
            for (i2, j2) in nz:
                assert not (i == i2 and j != j2 and a[j] == a[j2])
