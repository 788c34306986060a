"""GPU parity tests: the sm_100a BDF path through the C ABI against the CPU oracle (and, through
`step_and_interpolate`, directly against the reference's statistics snapshots)."""
import ctypes
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "reference_snapshots.json")) as f:
    GOLD = json.load(f)
BDF_CASES = [c for c in GOLD["cases"] if c["method"] == "bdf"]


@pytest.fixture(scope="module")
def dsb():
    import diffsol_b200
    from diffsol_b200 import capi
    capi.require_device()          # fails loudly: no CPU fallback
    return diffsol_b200


def _points(kind):
    from test_oracle_golden import solution_points
    return solution_points(kind)


def _build(dsb, case, nbatch=None, p=None, **kw):
    b = dsb.OdeBuilder().rhs_implicit(case["model"]).rtol(case["rtol"]).atol(case["atol"]).use_coloring(case["coloring"])
    if p is None:
        p = case["p"]
    if len(np.atleast_1d(p)):
        b = b.p(p)
    if nbatch:
        b = b.nbatch(nbatch)
    if kw:
        b = b.ode_options(**kw)
    return b.build()


@pytest.mark.parametrize("case", BDF_CASES, ids=[c["name"] for c in BDF_CASES])
def test_reference_snapshots_on_gpu(dsb, oracle, case):
    """The CUDA path reproduces the reference's inline statistics snapshots (bdf.rs) integer for
    integer, through the same stepping loop as the reference's harness, for every lane of a batch."""
    from test_oracle_golden import expected_stats
    t, ystar = _points(case["points"])
    nb = 5
    solver = _build(dsb, case, nbatch=nb).bdf()
    ys = solver.step_and_interpolate(t)
    assert (solver.status() == 0).all()
    for b in range(nb):
        assert solver.get_statistics(b) == expected_stats(case), case["cite"]
    # bit-identical to the oracle using the shared deterministic pow
    desc = oracle.make_desc(case["model"], rtol=case["rtol"], atol=case["atol"], use_coloring=case["coloring"], powmode=1)
    rc, ys_o, stats_o, fin = oracle.harness(desc, case["p"], t)
    assert rc == 0
    for b in range(nb):
        assert np.array_equal(ys[b], ys_o), case["name"]
    tf, hf, of = solver.final_state()
    assert tf[0] == fin["t"] and hf[0] == fin["h"] and of[0] == fin["order"]
    # the reference's acceptance test on the state (ode_solver/mod.rs:164-173)
    n = ystar.shape[1]
    atol = np.array(case["atol"] * n if len(case["atol"]) == 1 else case["atol"])
    w = np.abs(ystar) * case["rtol"] + atol
    err = np.sqrt(np.mean(((ys[0] - ystar) / w) ** 2, axis=1))
    assert (err < 20.0).all()


@pytest.mark.parametrize("model,tol", [("robertson_ode", "ROBERTSON_ODE_TOL"), ("robertson_dae", "ROBERTSON_DAE_TOL")])
def test_robertson_sweep_bit_exact(dsb, oracle, model, tol):
    """BASELINE config 2 at a size the oracle finishes in seconds: every instance of the rate-constant
    sweep has the same 13 counters, status and solve_dense output (bitwise) as the oracle."""
    from diffsol_b200 import sweeps
    B = 3001                                   # not a multiple of the block size on purpose
    tolkw = getattr(sweeps, tol)
    p = sweeps.robertson_sweep(np.arange(B))
    case = dict(model=model, coloring=False, **tolkw)
    solver = _build(dsb, case, p=p).bdf()
    ys = solver.solve_dense(sweeps.ROBERTSON_T_EVAL)
    desc = oracle.make_desc(model, powmode=1, **tolkw)
    ys_o, stats_o, status_o = oracle.batch_solve_dense(desc, p, sweeps.ROBERTSON_T_EVAL)
    assert np.array_equal(solver.status(), status_o)
    assert (status_o == 0).all()
    assert np.array_equal(solver.statistics_array()[:, :13], stats_o[:, :13])
    assert np.array_equal(ys, ys_o)
    # against the reference-literal libm pow (what diffsol's own powf calls): a last-ulp difference in pow very occasionally
    # flips a step-size decision -- measured 1e-4 of the instances; the bound is 1e-3, on a sample big enough
    # to resolve it
    B2 = 20000
    p2 = sweeps.robertson_sweep(np.arange(B2))
    s2 = _build(dsb, case, p=p2).bdf()
    ys2 = s2.solve_dense(sweeps.ROBERTSON_T_EVAL)
    desc0 = oracle.make_desc(model, powmode=0, **tolkw)
    ys_0, stats_0, _ = oracle.batch_solve_dense(desc0, p2, sweeps.ROBERTSON_T_EVAL)
    frac = np.mean((s2.statistics_array()[:, :13] != stats_0[:, :13]).any(axis=1))
    print("fraction of instances whose counters differ from the libm-pow oracle: %.5f" % frac)
    assert frac <= 1e-3
    ys, ys_0 = ys2, ys_0
    w = np.abs(ys_0) * tolkw["rtol"] + np.array(tolkw["atol"])
    assert (np.abs(ys - ys_0) <= 20 * w).all()
    # device-side reduction of a statistic agrees with the per-instance array
    assert solver.sum_statistic("number_of_nonlinear_solver_iterations") == int(stats_o[:, 8].sum())


def test_host_entry_point_pipelines_big_batches_in_chunks(dsb, monkeypatch):
    """dsb_batch_solve_dense_host splits big batches into child batches on their own streams (copies of one chunk under the
    kernels of another): same trajectories, counters, status and getters as the direct path."""
    from diffsol_b200 import sweeps
    B = 4 * 65536 + 1237
    p = sweeps.robertson_sweep(np.arange(B))
    case = dict(model="robertson_ode", coloring=False, **sweeps.ROBERTSON_ODE_TOL)
    monkeypatch.setenv("DSB_HOST_CHUNKS", "1")
    a = _build(dsb, case, p=p).bdf()
    ya = a.solve_dense(sweeps.ROBERTSON_T_EVAL)
    sa, sta, fa = a.statistics_array().copy(), a.status().copy(), a.final_state()
    monkeypatch.setenv("DSB_HOST_CHUNKS", "4")
    b = _build(dsb, case, p=p).bdf()
    yb = b.solve_dense(sweeps.ROBERTSON_T_EVAL)
    assert np.array_equal(ya, yb) and np.array_equal(sa, b.statistics_array()) and np.array_equal(sta, b.status())
    assert all(np.array_equal(u, v) for u, v in zip(fa, b.final_state()))
    assert b.sum_statistic("number_of_steps") == int(sa[:, 6].sum())
    assert b.last_kernel_ms() > 0.0


def test_failed_instances_keep_status(dsb, oracle):
    """An instance that fails does not abort the batch: same status code as the oracle, NaN outputs
    past the failure, neighbours unaffected."""
    from diffsol_b200 import sweeps
    p = sweeps.robertson_sweep(np.arange(8))
    p[3] = [0.04, 1e4, 3e7]
    case = dict(model="robertson_ode", coloring=False, **sweeps.ROBERTSON_ODE_TOL)
    kw = dict(max_nonlinear_solver_failures=2)      # the stiff transient needs more than that
    solver = _build(dsb, case, p=p, **kw).bdf()
    ys = solver.solve_dense(sweeps.ROBERTSON_T_EVAL)
    desc = oracle.make_desc("robertson_ode", powmode=1, options=kw, **sweeps.ROBERTSON_ODE_TOL)
    ys_o, stats_o, status_o = oracle.batch_solve_dense(desc, p, sweeps.ROBERTSON_T_EVAL)
    assert np.array_equal(solver.status(), status_o)
    assert (status_o != 0).any()
    assert np.array_equal(solver.statistics_array()[:, :13], stats_o[:, :13])
    assert np.array_equal(np.isnan(ys), np.isnan(ys_o))
    assert np.array_equal(np.nan_to_num(ys), np.nan_to_num(ys_o))


def test_edge_cases(dsb, oracle):
    from diffsol_b200 import sweeps
    case = dict(model="robertson_ode", coloring=False, **sweeps.ROBERTSON_ODE_TOL)
    desc = oracle.make_desc("robertson_ode", powmode=1, **sweeps.ROBERTSON_ODE_TOL)
    # a single instance, a single output time
    solver = _build(dsb, case, p=[0.04, 1e4, 3e7]).bdf()
    ys = solver.solve_dense([40.0])
    rc, ys_o, st_o, _ = oracle.solve_dense(desc, [0.04, 1e4, 3e7], [40.0])
    assert rc == 0 and np.array_equal(ys[0], ys_o) and solver.get_statistics(0) == st_o
    # repeated output times and a first output at t0: solve_dense writes every column
    t_eval = [0.0, 1.0, 1.0, 2.5]
    ys = solver.solve_dense(t_eval)
    rc, ys_o, st_o, _ = oracle.solve_dense(desc, [0.04, 1e4, 3e7], t_eval)
    assert rc == 0 and np.array_equal(ys[0], ys_o) and solver.get_statistics(0) == st_o
    # stop time equal to t0 is an error in the reference (StopTimeAtCurrentTime, bdf.rs:1593-1597)
    ys = solver.solve_dense([0.0])
    assert solver.status()[0] == 5 and np.isnan(ys).all()
    # argument errors come back as DSB_BAD_ARG with a message
    with pytest.raises(dsb.DiffsolB200Error):
        solver.solve_dense([2.0, 1.0])
    with pytest.raises(ValueError):
        dsb.OdeBuilder().rhs_implicit("robertson_ode").p([1.0, 2.0]).build()


def test_non_default_options(dsb, oracle):
    """Options travel to the kernels: a PI controller with a proportional term and tighter refresh
    thresholds change the trace, identically on both sides."""
    from diffsol_b200 import sweeps
    kw = dict(pi_control_proportional=0.2, update_jacobian_after_steps=5, threshold_to_update_jacobian=0.1,
              max_nonlinear_solver_iterations=6)
    p = sweeps.robertson_sweep(np.arange(64))
    case = dict(model="robertson_ode", coloring=False, **sweeps.ROBERTSON_ODE_TOL)
    solver = _build(dsb, case, p=p, **kw).bdf()
    ys = solver.solve_dense(sweeps.ROBERTSON_T_EVAL)
    desc = oracle.make_desc("robertson_ode", powmode=1, options=kw, **sweeps.ROBERTSON_ODE_TOL)
    ys_o, stats_o, status_o = oracle.batch_solve_dense(desc, p, sweeps.ROBERTSON_T_EVAL)
    assert np.array_equal(solver.status(), status_o)
    assert np.array_equal(solver.statistics_array()[:, :13], stats_o[:, :13])
    assert np.array_equal(ys, ys_o, equal_nan=True)


@pytest.mark.parametrize("n", [1, 2, 3, 5, 8, 9, 17])
def test_batched_lu_matches_nalgebra_restatement(dsb, oracle, n):
    """dsb_lu_factor_batched / dsb_lu_solve_batched (the LinearSolver pair) give the factors, pivots
    and solutions of the oracle's nalgebra-style LU, bit for bit."""
    import torch
    from diffsol_b200 import capi
    rng = np.random.default_rng(n)
    B = 777
    A = rng.standard_normal((B, n, n))
    A[5] = 0.0                                   # a singular instance
    if n > 1:
        A[6, :, 0] = 0.0                         # a zero column: nalgebra skips it
    rhs = rng.standard_normal((B, n))
    dev = torch.device("cuda:0")
    # batch-major device layouts: a[(j*n + i)*B + b], b[i*B + b]
    a_dev = torch.from_numpy(np.ascontiguousarray(A.transpose(2, 1, 0))).to(dev)      # [j][i][b]
    b_dev = torch.from_numpy(np.ascontiguousarray(rhs.T)).to(dev)
    piv = torch.zeros((n, B), dtype=torch.int32, device=dev)
    info = torch.zeros(B, dtype=torch.int32, device=dev)
    info2 = torch.zeros(B, dtype=torch.int32, device=dev)
    L = capi.lib()
    vp = ctypes.c_void_p
    capi.check(L.dsb_lu_factor_batched(vp(a_dev.data_ptr()), n, B, vp(piv.data_ptr()), vp(info.data_ptr()), None))
    capi.check(L.dsb_lu_solve_batched(vp(a_dev.data_ptr()), vp(piv.data_ptr()), vp(b_dev.data_ptr()), n, B,
                                      vp(info2.data_ptr()), None))
    torch.cuda.synchronize()
    lu_g = a_dev.cpu().numpy().transpose(2, 1, 0)        # [b][i][j]
    piv_g = piv.cpu().numpy().T
    x_g = b_dev.cpu().numpy().T
    info2 = info2.cpu().numpy()
    dp = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
    for b in range(B):
        Af = np.ascontiguousarray(A[b].T).ravel()        # column-major
        lu_o = np.empty(n * n); piv_o = np.empty(n, dtype=np.int32)
        oracle.lib().orc_lu_factor(dp(Af), n, dp(lu_o), piv_o.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)))
        assert np.array_equal(lu_g[b], lu_o.reshape(n, n).T), b
        assert np.array_equal(piv_g[b], piv_o), b
        x_o = rhs[b].copy()
        rc = oracle.lib().orc_lu_solve(dp(Af), n, dp(x_o))
        assert (rc != 0) == (info2[b] != 0), b
        if rc == 0:
            assert np.array_equal(x_g[b], x_o), b


@pytest.mark.parametrize("n", [9, 17, 32, 33, 64, 100, 200, 256])
def test_cooperative_lu_matches_nalgebra_restatement(dsb, oracle, n):
    """dsb_lu_factor_instance_major / dsb_lu_solve_instance_major (one thread block per instance, blocked
    with 32-column panels) still give nalgebra's factors, pivots and solutions bit for bit."""
    import torch
    from diffsol_b200 import capi
    rng = np.random.default_rng(1000 + n)
    B = 37
    A = rng.standard_normal((B, n, n))               # A[b][i][j]
    A[3] = 0.0                                       # singular
    A[4, :, 0] = 0.0                                 # zero first column: skipped by nalgebra
    A[5, :, n // 2] = 0.0                            # zero column in the middle (becomes zero after elimination only if...)
    A[6] = np.triu(A[6])                             # already upper triangular: no swaps needed below diagonal
    A[7] = A[7] * (10.0 ** rng.uniform(-8, 8, size=(n, 1)))   # badly scaled rows: lots of pivoting
    rhs = rng.standard_normal((B, n))
    dev = torch.device("cuda:0")
    a_dev = torch.from_numpy(np.ascontiguousarray(A.transpose(0, 2, 1))).to(dev)      # [b][j][i] = column-major per instance
    b_dev = torch.from_numpy(rhs.copy()).to(dev)
    piv = torch.zeros((B, n), dtype=torch.int32, device=dev)
    info = torch.zeros(B, dtype=torch.int32, device=dev)
    info2 = torch.zeros(B, dtype=torch.int32, device=dev)
    L = capi.lib()
    vp = ctypes.c_void_p
    capi.check(L.dsb_lu_factor_instance_major(vp(a_dev.data_ptr()), n, B, vp(piv.data_ptr()), vp(info.data_ptr()), None))
    capi.check(L.dsb_lu_solve_instance_major(vp(a_dev.data_ptr()), vp(piv.data_ptr()), vp(b_dev.data_ptr()), n, B,
                                             vp(info2.data_ptr()), None))
    torch.cuda.synchronize()
    lu_g = a_dev.cpu().numpy()                       # [b][j][i]
    piv_g = piv.cpu().numpy()
    x_g = b_dev.cpu().numpy()
    info2 = info2.cpu().numpy()
    dp = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
    for b in range(B):
        Af = np.ascontiguousarray(A[b].T).ravel()
        lu_o = np.empty(n * n); piv_o = np.empty(n, dtype=np.int32)
        oracle.lib().orc_lu_factor(dp(Af), n, dp(lu_o), piv_o.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)))
        assert np.array_equal(piv_g[b], piv_o), b
        assert np.array_equal(lu_g[b].ravel(), lu_o, equal_nan=True), b
        x_o = rhs[b].copy()
        rc = oracle.lib().orc_lu_solve(dp(Af), n, dp(x_o))
        assert (rc != 0) == (info2[b] != 0), b
        if rc == 0:
            assert np.array_equal(x_g[b], x_o), b


@pytest.mark.parametrize("method", ["bdf", "tr_bdf2", "esdirk34"])
def test_backward_integration(dsb, oracle, method):
    """negative_exponential_decay_problem (h0 = -1, points 0, -1, .., -9) through step()/interpolate(), as the
    reference's bdf.rs:1729-1733 / sdirk.rs:669-673 do; solve_dense itself is forward-only in the reference."""
    k, y0 = 0.1, np.exp(-1.0)
    pts = -np.arange(0.0, 10.0)
    p = np.tile([[k, y0]], (40, 1))
    prob = dsb.OdeBuilder().rhs_implicit("exp_decay").p(p).h0(-1.0).build()
    solver = getattr(prob, method)()
    ys = solver.step_and_interpolate(pts)
    desc = oracle.make_desc("exp_decay", method=method, powmode=1, h0=-1.0)
    rc, ys_o, stats_o, fin = oracle.harness(desc, [k, y0], pts)
    assert rc == 0 and (solver.status() == 0).all()
    for b in (0, 17, 39):
        assert np.array_equal(ys[b], ys_o) and solver.get_statistics(b) == stats_o
    tf, hf, _ = solver.final_state()
    assert tf[0] == fin["t"] and hf[0] == fin["h"] and hf[0] < 0
    with pytest.raises(dsb.DiffsolB200Error):
        solver.solve_dense(pts[1:])                  # decreasing t_eval: rejected


@pytest.mark.parametrize("method", ["bdf", "tr_bdf2"])
def test_warp_scheduling_knobs_never_change_results(dsb, method, monkeypatch):
    """The lane kernels' warp scheduler (quorum of the slow pool, Newton iterations per trip: dsb_bdf_kernel.cuh) decides
    WHEN a lane runs a block, never what it computes: counters, states and status are bitwise the same for every setting."""
    from diffsol_b200 import sweeps
    nb = 3000
    p = sweeps.robertson_sweep(np.arange(nb))
    tol = sweeps.ROBERTSON_ODE_TOL

    def run():
        solver = getattr(dsb.OdeBuilder().rhs_implicit("robertson_ode").p(p).rtol(tol["rtol"]).atol(tol["atol"]).build(), method)()
        ys = solver.solve_dense(sweeps.ROBERTSON_T_EVAL)
        return ys.copy(), solver.statistics_array().copy(), solver.status().copy()

    ys0, st0, status0 = run()
    assert (status0 == 0).all()
    for passes, quorum in ((1, 16), (2, 8), (5, 24), (16, 33)):
        monkeypatch.setenv("DSB_NEWTON_PASSES", str(passes))
        monkeypatch.setenv("DSB_QUORUM", str(quorum))
        ys, st, status = run()
        assert np.array_equal(ys.view(np.uint64), ys0.view(np.uint64)), (passes, quorum)
        assert np.array_equal(st, st0) and np.array_equal(status, status0), (passes, quorum)


def test_sensitivity_hold_is_scheduling_only(dsb, monkeypatch):
    """The end-of-step hold of the sensitivity instantiation (lanes wait for a quorum in front of POST) likewise."""
    from diffsol_b200 import sweeps
    nb = 1500
    p = sweeps.robertson_sweep(np.arange(nb))
    tol = sweeps.ROBERTSON_ODE_TOL

    def run():
        solver = (dsb.OdeBuilder().rhs_implicit("robertson_ode").p(p).rtol(tol["rtol"]).atol(tol["atol"])
                  .sens_rtol(tol["rtol"]).sens_atol([1e-6] * 3).build().bdf_sens())
        ys, ss = solver.solve_dense_sensitivities(sweeps.ROBERTSON_T_EVAL)
        return ys.copy(), ss.copy(), solver.statistics_array().copy()

    ys0, ss0, st0 = run()
    for passes, quorum in ((1, 1), (3, 33), (4, 6)):
        monkeypatch.setenv("DSB_NEWTON_PASSES", str(passes))
        monkeypatch.setenv("DSB_QUORUM", str(quorum))
        ys, ss, st = run()
        assert np.array_equal(ys.view(np.uint64), ys0.view(np.uint64)), (passes, quorum)
        assert np.array_equal(ss.view(np.uint64), ss0.view(np.uint64)), (passes, quorum)
        assert np.array_equal(st, st0), (passes, quorum)
