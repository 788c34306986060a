"""`OdeSolverMethod::solve(final_time)` (/root/reference/crates/diffsol/src/ode_solver/method.rs:227-258,
fn solve :881-961, write_out :965-1000): one result column per INTERNAL step.  Every instance of a batch takes its own number
of steps, so the device API is two passes -- dsb_batch_solve_count (integrate, count), then dsb_batch_solve_write (integrate
again, write at the prefix-sum offsets) -- on the DsbRagged<M> instantiation of the on-chip lane kernels.

CPU: the reference's own tests of solve() restated on the oracle (test_solve, test_solve_stops_on_root, method.rs:1068-1101),
and the kernel SOURCE (host emulation) bit for bit against the oracle.  GPU: the CUDA path through the C ABI against the oracle."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)


def sweep(B):
    from diffsol_b200 import sweeps
    i = np.arange(B)
    return np.stack([0.02 * 50.0 ** sweeps.uniform(i, 0), 0.4 + 1.6 * sweeps.uniform(i, 1)], axis=1)


def weighted_norm(y, ystar, rtol=1e-6, atol=1e-6):
    return float(np.sqrt(np.mean(((y - ystar) / (np.abs(ystar) * rtol + atol)) ** 2)))


def test_reference_test_solve(oracle):
    """method.rs:1068-1083: t[0] = 0, t[-1] = 10, every column within 15 tolerances of y0 e^{-k t}."""
    ts, ys, nc, st, status, roots = oracle.batch_solve_ragged(oracle.make_desc("exp_decay"), np.array([[0.1, 1.0]]), 10.0)
    assert status[0] == 0 and nc[0] == st[0, 6] + 1              # the initial column, then one per step
    t = ts[0, :nc[0]]
    assert abs(t[0]) < 1e-10 and abs(t[-1] - 10.0) < 1e-10 and (np.diff(t) > 0).all()
    for k in range(nc[0]):
        assert weighted_norm(ys[0, k], np.exp(-0.1 * t[k]) * np.ones(2)) < 15.0


def test_reference_test_solve_stops_on_root(oracle):
    """method.rs:1085-1101: RootFound, the last time within 1e-3 of -ln(0.6) / 0.1, the last column within 15 tolerances of 0.6."""
    ts, ys, nc, st, status, roots = oracle.batch_solve_ragged(oracle.make_desc("exp_decay_root"), np.array([[0.1, 1.0]]), 10.0)
    assert status[0] == 0 and int(roots[0, 1]) == 0
    t_root = -np.log(0.6) / 0.1
    assert abs(ts[0, nc[0] - 1] - t_root) < 1e-3 and abs(roots[0, 0] - t_root) < 1e-3
    assert weighted_norm(ys[0, nc[0] - 1], np.array([0.6, 0.6])) < 15.0


@pytest.mark.parametrize("method", ["bdf", "tr_bdf2", "esdirk34"])
@pytest.mark.parametrize("model,model_id", [("exp_decay", 0), ("exp_decay_root", 13)])
def test_kernel_source_equals_oracle(oracle, model, model_id, method):
    from host_emu import emu
    p = sweep(24)
    ts, ys, nc, st, status, roots = oracle.batch_solve_ragged(oracle.make_desc(model, method=method, powmode=1), p, 10.0)
    r = emu.solve_ragged(model_id, 2, 2, p, 10.0, method=method)
    assert np.array_equal(r["status"], status) and (status == 0).all()
    assert np.array_equal(r["ncols"], nc) and np.array_equal(r["stats"][:, :13], st[:, :13])
    assert np.array_equal(r["root_idx"], roots[:, 1].astype(np.int32))
    for b in range(len(p)):
        assert np.array_equal(r["ts"][b, :nc[b]], ts[b, :nc[b]]) and np.array_equal(r["ys"][b, :nc[b]], ys[b, :nc[b]])


def test_kernel_source_equals_oracle_robertson(oracle):
    from host_emu import emu
    from diffsol_b200 import sweeps
    p = sweeps.robertson_sweep(np.arange(8))
    tol = sweeps.ROBERTSON_ODE_TOL
    ts, ys, nc, st, status, roots = oracle.batch_solve_ragged(oracle.make_desc("robertson_ode", powmode=1, **tol), p, 1e4)
    r = emu.solve_ragged(3, 3, 3, p, 1e4, **tol)
    assert np.array_equal(r["ncols"], nc) and np.array_equal(r["stats"][:, :13], st[:, :13]) and (nc > 100).all()
    for b in range(len(p)):
        assert np.array_equal(r["ts"][b, :nc[b]], ts[b, :nc[b]]) and np.array_equal(r["ys"][b, :nc[b]], ys[b, :nc[b]])


# ---- the CUDA path -------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def dsb():
    import diffsol_b200
    from diffsol_b200 import capi
    capi.require_device()
    return diffsol_b200


def check_against_oracle(solver, oracle, desc, p, final_time):
    ys, ts, off = solver.solve(final_time)
    ts_o, ys_o, nc_o, st_o, status_o, roots_o = oracle.batch_solve_ragged(desc, p, final_time)
    assert np.array_equal(np.diff(off), nc_o) and off[0] == 0 and off[-1] == len(ts) == len(ys)
    assert np.array_equal(solver.status(), status_o) and (status_o == 0).all()
    assert np.array_equal(solver.statistics_array()[:, :13], st_o[:, :13])
    keep = np.arange(ts_o.shape[1])[None, :] < nc_o[:, None]                 # instance-major, each instance's first ncols columns
    assert np.array_equal(ts, ts_o[keep]) and np.array_equal(ys, ys_o[keep])
    return ys, ts, off, roots_o


@pytest.mark.gpu
@pytest.mark.parametrize("method", ["bdf", "tr_bdf2", "esdirk34"])
def test_gpu_solve_bit_exact(dsb, oracle, method):
    B = 3000
    p = sweep(B)
    solver = getattr(dsb.OdeBuilder().rhs_implicit("exp_decay").p(p).build(), method)()
    ys, ts, off, _ = check_against_oracle(solver, oracle, oracle.make_desc("exp_decay", method=method, powmode=1), p, 10.0)
    assert (ts[off[:-1]] == 0.0).all() and (ts[off[1:] - 1] == 10.0).all()     # every run starts at t0 and ends at the final time
    assert len(set(np.diff(off).tolist())) > 20                                   # ragged indeed


@pytest.mark.gpu
@pytest.mark.parametrize("method", ["bdf", "esdirk34"])
def test_gpu_solve_stops_on_root(dsb, oracle, method):
    B = 2000
    p = sweep(B)
    solver = getattr(dsb.OdeBuilder().rhs_implicit("exp_decay_root").p(p).build(), method)()
    ys, ts, off, roots = check_against_oracle(solver, oracle, oracle.make_desc("exp_decay_root", method=method, powmode=1), p, 10.0)
    root_idx, _ = solver.root_info()
    assert np.array_equal(root_idx, roots[:, 1].astype(np.int32))
    stopped = root_idx == 0
    assert 0 < stopped.sum() < B
    last = off[1:] - 1
    assert np.abs(ys[last[stopped], 0] - 0.6).max() < 1e-5 and np.array_equal(ts[last[stopped]], roots[stopped, 0])
    assert (ts[last[~stopped]] == 10.0).all()


@pytest.mark.gpu
def test_gpu_solve_robertson(dsb, oracle):
    from diffsol_b200 import sweeps
    p = sweeps.robertson_sweep(np.arange(1500))
    tol = sweeps.ROBERTSON_ODE_TOL
    solver = dsb.OdeBuilder().rhs_implicit("robertson_ode").p(p).rtol(tol["rtol"]).atol(tol["atol"]).build().bdf()
    check_against_oracle(solver, oracle, oracle.make_desc("robertson_ode", powmode=1, **tol), p, 1e4)


@pytest.mark.gpu
def test_gpu_solve_errors(dsb):
    import ctypes
    from diffsol_b200 import capi
    # equations with a reset function, and systems beyond the on-chip kernels, are not built in this form
    with pytest.raises(capi.DiffsolB200Error, match="solve\\(final_time\\)"):
        dsb.OdeBuilder().rhs_implicit("exp_decay_reset").p(sweep(4)).build().bdf().solve(10.0)
    with pytest.raises(capi.DiffsolB200Error, match="solve\\(final_time\\)"):
        dsb.OdeBuilder().rhs_implicit("heat1d_dae_32").p(np.ones((2, 3))).build().bdf().solve(0.5)
    # the writing pass needs its counting pass
    solver = dsb.OdeBuilder().rhs_implicit("exp_decay").p(sweep(4)).build().bdf()
    buf = np.zeros(16)
    L = capi.lib()
    assert L.dsb_batch_solve_write_host(solver._b, 0, 10.0, ctypes.c_void_p(buf.ctypes.data), ctypes.c_void_p(buf.ctypes.data)) != 0
    assert b"dsb_batch_solve_count" in L.dsb_last_error()
