"""GPU parity tests of the banded thread-per-instance BDF path (dsb_band_bdf_kernel.cuh: state in global memory,
band LU per lane): the single-particle battery model of BASELINE config 5 against the oracle's DENSE LU, with and
without Jacobian colouring, and against the block-per-instance path; the heat-equation DAE of BASELINE config 4
(singular mass: consistent initialisation by dsb_band_init_kernel.cuh, M - cJ in band storage)."""
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


@pytest.mark.parametrize("block", ["768", "128"])
@pytest.mark.parametrize("coloring", [False, True])
def test_spm_band_bit_exact(dsb, oracle, coloring, block, monkeypatch):
    monkeypatch.setenv("DSB_BAND_BLOCK", block)           # both block shapes of the kernel (big batches / small batches)
    """Counters, status and states bit-identical to the oracle (dense partial-pivoting LU restated from nalgebra):
    the band LU only skips operations whose multiplier or pivot-row entry is exactly zero."""
    B = 700                      # more than one warp per block, ragged tail
    current = spm_currents(B)
    t_eval = np.arange(1, 13) * 300.0
    solver = (dsb.OdeBuilder().rhs_implicit("spm").p(current).use_coloring(coloring).build().bdf()
              .set_execution("band"))
    ys = solver.solve_dense(t_eval)
    desc = oracle.make_desc("spm", powmode=1, use_coloring=coloring)
    ys_o, stats_o, status_o = oracle.batch_solve_dense(desc, current, t_eval)
    assert np.array_equal(solver.status(), status_o) and (status_o == 0).all()
    assert np.array_equal(solver.statistics_array()[:, :13], stats_o[:, :13])
    assert np.array_equal(ys, ys_o)
    if coloring:
        me = solver.statistics_array()[:, 12]
        assert (solver.statistics_array()[:, 11] == 3 * me + 42).all()      # 3 colours + n probes


def test_spm_band_tight_tolerances_and_free_running(dsb, oracle):
    """Tighter tolerances (higher orders, more rescales) and the step()/interpolate() loop."""
    B = 96
    current = spm_currents(B)
    t_eval = np.arange(1, 7) * 500.0
    solver = dsb.OdeBuilder().rhs_implicit("spm").p(current).rtol(1e-9).atol(1e-10).build().bdf().set_execution("band")
    ys = solver.solve_dense(t_eval)
    desc = oracle.make_desc("spm", powmode=1, rtol=1e-9, atol=1e-10)
    ys_o, stats_o, status_o = oracle.batch_solve_dense(desc, current, t_eval)
    assert np.array_equal(solver.status(), status_o) and (status_o == 0).all()
    assert np.array_equal(solver.statistics_array()[:, :13], stats_o[:, :13])
    assert np.array_equal(ys, ys_o)


def test_spm_band_equals_block_per_instance(dsb):
    B = 200
    current = spm_currents(B)
    t_eval = np.arange(1, 13) * 300.0
    prob = dsb.OdeBuilder().rhs_implicit("spm").p(current).build()
    a = prob.bdf().set_execution("band")
    b = prob.bdf().set_execution("block")
    ya, yb = a.solve_dense(t_eval), b.solve_dense(t_eval)
    assert np.array_equal(ya, yb)
    assert np.array_equal(a.statistics_array(), b.statistics_array())
    # automatic selection picks the banded path for this model
    c = prob.bdf()
    assert np.array_equal(c.solve_dense(t_eval), ya)


@pytest.mark.parametrize("coloring", [False, True])
def test_spm99_band_bit_exact(dsb, oracle, coloring):
    """The same model on 99 radial cells per particle (n = 200, BASELINE config 5's size): colours and pattern
    travel in a device array, tolerances too."""
    B = 100
    current = spm_currents(B)
    t_eval = np.arange(1, 7) * 600.0
    solver = (dsb.OdeBuilder().rhs_implicit("spm99").p(current).use_coloring(coloring).build().bdf()
              .set_execution("band"))
    ys = solver.solve_dense(t_eval)
    desc = oracle.make_desc("spm99", powmode=1, use_coloring=coloring)
    ys_o, stats_o, status_o = oracle.batch_solve_dense(desc, current, t_eval)
    assert np.array_equal(solver.status(), status_o) and (status_o == 0).all()
    assert np.array_equal(solver.statistics_array()[:, :13], stats_o[:, :13])
    assert np.array_equal(ys, ys_o)
    assert ys.shape == (B, 6, 200)


def test_spm99_band_equals_block_per_instance(dsb):
    B = 24
    current = spm_currents(B)
    t_eval = np.arange(1, 4) * 600.0
    prob = dsb.OdeBuilder().rhs_implicit("spm99").p(current).build()
    a = prob.bdf().set_execution("band")
    b = prob.bdf().set_execution("block")
    assert np.array_equal(a.solve_dense(t_eval), b.solve_dense(t_eval))
    assert np.array_equal(a.statistics_array(), b.statistics_array())


def test_band_execution_rejected_where_it_does_not_apply(dsb):
    p = np.tile(np.array([[0.04, 1.0e4, 3.0e7]]), (4, 1))
    prob = dsb.OdeBuilder().rhs_implicit("robertson_ode").p(p).build()
    with pytest.raises(dsb.DiffsolB200Error):
        prob.bdf().set_execution("band").solve_dense([1.0])


@pytest.mark.parametrize("block", ["768", "128"])
@pytest.mark.parametrize("model,B,coloring", [("heat1d_dae_32", 200, False), ("heat1d_dae_32", 200, True),
                                               ("heat1d_dae_256", 40, True)])
def test_heat_dae_band_bit_exact(dsb, oracle, model, B, coloring, block, monkeypatch):
    """BASELINE config 4 on the banded lane path: counters (including those of the consistent initialisation), status
    and all 100 dense-output columns bit-identical to the oracle's dense LU."""
    monkeypatch.setenv("DSB_BAND_BLOCK", block)
    p = heat_params(B)
    solver = (dsb.OdeBuilder().rhs_implicit(model).p(p).rtol(1e-6).atol(1e-6).use_coloring(coloring).build().bdf()
              .set_execution("band"))
    ys = solver.solve_dense(HEAT_T_EVAL)
    desc = oracle.make_desc(model, powmode=1, rtol=1e-6, atol=1e-6, use_coloring=coloring)
    ys_o, stats_o, status_o = oracle.batch_solve_dense(desc, p, HEAT_T_EVAL)
    assert np.array_equal(solver.status(), status_o) and (status_o == 0).all()
    assert np.array_equal(solver.statistics_array()[:, :13], stats_o[:, :13])
    assert np.array_equal(ys, ys_o)
    assert (ys[:, -1].max(axis=1) < p[:, 0]).all()


def test_heat_dae_band_equals_block_per_instance_and_is_default(dsb):
    B = 64
    p = heat_params(B)
    prob = dsb.OdeBuilder().rhs_implicit("heat1d_dae_256").p(p).rtol(1e-6).atol(1e-6).build()
    a = prob.bdf().set_execution("band")
    b = prob.bdf().set_execution("block")
    ya, yb = a.solve_dense(HEAT_T_EVAL), b.solve_dense(HEAT_T_EVAL)
    assert np.array_equal(ya, yb)
    assert np.array_equal(a.statistics_array(), b.statistics_array())
    # automatic selection: the banded warp-per-instance kernel for BDF (tests/test_gpu_warp_band_parity.py), same bits
    c = prob.bdf()
    w = prob.bdf().set_execution("warp")
    assert np.array_equal(c.solve_dense(HEAT_T_EVAL), ya) and np.array_equal(w.solve_dense(HEAT_T_EVAL), ya)
    assert c.last_launch_count() == w.last_launch_count()


@pytest.mark.parametrize("method", ["tr_bdf2", "esdirk34"])
@pytest.mark.parametrize("model,B,coloring,block", [("spm", 300, False, "768"), ("spm", 300, True, "128"),
                                                     ("spm99", 40, True, "768"), ("heat1d_dae_32", 200, False, "128"),
                                                     ("heat1d_dae_256", 24, True, "768")])
def test_band_sdirk_bit_exact(dsb, oracle, method, model, B, coloring, block, monkeypatch):
    """(E)SDIRK on the banded lane path (dsb_band_sdirk_kernel.cuh), ODE and singular-mass DAE: counters, status and
    states bit-identical to the oracle's dense LU."""
    monkeypatch.setenv("DSB_BAND_BLOCK", block)
    if model.startswith("spm"):
        p, t_eval, tol = spm_currents(B), np.arange(1, 13) * 300.0, dict(rtol=1e-6, atol=1e-6)
    else:
        p, t_eval, tol = heat_params(B), HEAT_T_EVAL, dict(rtol=1e-6, atol=1e-6)
    prob = dsb.OdeBuilder().rhs_implicit(model).p(p).rtol(tol["rtol"]).atol(tol["atol"]).use_coloring(coloring).build()
    solver = getattr(prob, method)().set_execution("band")
    ys = solver.solve_dense(t_eval)
    desc = oracle.make_desc(model, method=method, powmode=1, use_coloring=coloring, **tol)
    ys_o, stats_o, status_o = oracle.batch_solve_dense(desc, p, t_eval)
    assert np.array_equal(solver.status(), status_o) and (status_o == 0).all()
    assert np.array_equal(solver.statistics_array()[:, :13], stats_o[:, :13])
    assert np.array_equal(ys, ys_o)


def test_band_sdirk_free_running_and_default_path(dsb, oracle):
    """step()/interpolate() loop without a stop time; automatic selection picks the banded kernel for SDIRK too."""
    B = 64
    p = spm_currents(B)
    t_pts = np.arange(1, 7) * 500.0
    solver = dsb.OdeBuilder().rhs_implicit("spm").p(p).build().tr_bdf2()
    ys = solver.step_and_interpolate(t_pts)
    n, np_, _ = oracle.model_dims("spm")
    desc = oracle.make_desc("spm", method="tr_bdf2", powmode=1)
    for b in (0, 17, 63):
        rc, ys_o, stats_o, _ = oracle.harness(desc, p[b], t_pts)
        assert rc == 0
        assert np.array_equal(ys[b], ys_o)
        st = solver.get_statistics(b)
        assert all(st[k] == v for k, v in stats_o.items())


@pytest.mark.parametrize("method,execution", [("bdf", "band"), ("tr_bdf2", "band"), ("bdf", "block")])
def test_dae_inconsistent_initial_values_bit_exact(dsb, oracle, method, execution):
    """Boundary rows 0 = u - height / 4 with u = 0 initially: `new_and_consistent` (state.rs:84-162) has to move the
    algebraic components -- dsb_band_init_kernel.cuh on the banded path, the cooperative kernel's own initialisation on
    the block-per-instance path -- before the first step; counters include the initialisation's rhs calls."""
    B = 150
    p = heat_params(B)
    t_eval = np.arange(1, 11) / 10.0 * 0.99
    prob = dsb.OdeBuilder().rhs_implicit("heat1d_dae_32_bc").p(p).rtol(1e-6).atol(1e-6).build()
    solver = getattr(prob, method)().set_execution(execution)
    ys = solver.solve_dense(t_eval)
    desc = oracle.make_desc("heat1d_dae_32_bc", method=method, powmode=1, rtol=1e-6, atol=1e-6)
    ys_o, stats_o, status_o = oracle.batch_solve_dense(desc, p, t_eval)
    assert np.array_equal(solver.status(), status_o) and (status_o == 0).all()
    assert np.array_equal(solver.statistics_array()[:, :13], stats_o[:, :13])
    assert np.array_equal(ys, ys_o)
    assert np.abs(ys[:, :, 0] - 0.25 * p[:, :1]).max() < 1e-12
