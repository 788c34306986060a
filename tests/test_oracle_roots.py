"""CPU tests of the oracle's event handling (RootFinder, nonlinear_solver/root.rs; Bdf::step root check,
bdf.rs:1566-1579; fn solve_dense RootFound branch, method.rs:774-805, 493-503) against the reference's own tests of
the same problem: exponential_decay_problem_with_root (test_models/exponential_decay.rs:370-390, root y[0] - 0.6)."""
import numpy as np
import pytest


def weighted_norm(y, ystar, atol, rtol):
    return float(np.sqrt(np.mean(((y - ystar) / (np.abs(ystar) * rtol + atol)) ** 2)))


def test_dense_solve_stops_on_root(oracle):
    """ode_solver/method.rs:1141-1163 test_dense_solve_stops_on_root, same assertions."""
    k, y0 = 0.1, 1.0
    desc = oracle.make_desc("exp_decay_root")           # builder defaults rtol = atol = 1e-6, libm pow
    t_eval = np.arange(0.0, 11.0)
    ys, stats, status, t_root_found, root_idx, ncols = oracle.batch_solve_dense_roots(desc, [[k, y0]], t_eval)
    assert status[0] == 0 and root_idx[0] == 0           # OdeSolverStopReason::RootFound(_, 0)
    t_root = -np.log(0.6) / k
    assert abs(t_root_found[0] - t_root) < 1e-3
    assert ncols[0] < len(t_eval)
    minus_one = int(np.argmax(t_eval >= t_root)) - 1
    expected = np.full(2, np.exp(-k * t_eval[minus_one]))
    assert weighted_norm(ys[0, minus_one], expected, 1e-6, 1e-6) < 15.0
    assert weighted_norm(ys[0, ncols[0] - 1], np.array([0.6, 0.6]), 1e-6, 1e-6) < 15.0
    assert np.isnan(ys[0, ncols[0]:]).all()


def test_root_finder_known_answer(oracle):
    """nonlinear_solver/root.rs:177-221 test_root through the integrator: the located root is within 1e-10 relative
    of where the interpolant crosses, i.e. the returned state satisfies g = 0 to the interpolation accuracy."""
    desc = oracle.make_desc("exp_decay_root", powmode=1)
    ks = np.linspace(0.05, 0.9, 18)
    p = np.stack([ks, np.full_like(ks, 1.0)], axis=1)
    ys, stats, status, t_root, root_idx, ncols = oracle.batch_solve_dense_roots(desc, p, np.arange(1.0, 41.0))
    assert (status == 0).all() and (root_idx == 0).all()
    assert np.allclose(t_root, -np.log(0.6) / ks, rtol=1e-4)
    at_root = ys[np.arange(len(ks)), ncols - 1, 0]
    assert np.abs(at_root - 0.6).max() < 1e-9            # g(y(t_root)) = 0 on the interpolant


def test_no_root_runs_to_the_end(oracle):
    desc = oracle.make_desc("exp_decay_root", powmode=1)
    t_eval = np.arange(1.0, 6.0)
    ys, stats, status, t_root, root_idx, ncols = oracle.batch_solve_dense_roots(desc, [[0.1, 0.5], [0.01, 1.0]], t_eval)
    assert (status == 0).all() and (root_idx == -1).all() and (ncols == len(t_eval)).all()
    desc2 = oracle.make_desc("exp_decay", powmode=1)
    ys2, stats2, status2 = oracle.batch_solve_dense(desc2, [[0.1, 0.5], [0.01, 1.0]], t_eval)
    assert np.array_equal(ys, ys2) and np.array_equal(stats, stats2)      # the root function does not steer the steps


def test_root_finder_sdirk(oracle):
    """ode_solver/sdirk.rs:1027-1047 test_root_finder_tr_bdf2 (and the same for esdirk34): the solve stops at the root
    of y[0] - 0.6 and the state there is y0 exp(-k t_root) within the reference's norm bound of 15."""
    k, y0 = 0.1, 1.0
    t_root = -np.log(0.6 / y0) / k
    for method in ("tr_bdf2", "esdirk34"):
        desc = oracle.make_desc("exp_decay_root", method=method)      # builder defaults, libm pow
        t_eval = np.arange(0.0, 11.0)
        ys, stats, status, t_found, root_idx, ncols = oracle.batch_solve_dense_roots(desc, [[k, y0]], t_eval)
        assert status[0] == 0 and root_idx[0] == 0
        assert abs(t_found[0] - t_root) < 1e-3
        expected = np.full(2, y0 * np.exp(-k * t_root))
        assert weighted_norm(ys[0, ncols[0] - 1], expected, 1e-6, 1e-6) < 15.0
        assert np.isnan(ys[0, ncols[0]:]).all()


def spm_voltage_numpy(cn18, cn19, cp18, cp19, cur):
    """The model text's out_i (book/src/primer/src/spm.ds) with numpy's libm functions."""
    v2 = -25608.96286546366 * cp18 + 76826.88859639116 * cp19
    v3 = -0.4999999999999983 * cp18 + 1.4999999999999982 * cp19
    v4 = -12491.630996921805 * cn18 + 37474.892990765504 * cn19
    v5 = -0.4999999999999983 * cn18 + 1.4999999999999984 * cn19
    cps = max(min(v2, 51217.92521874824), 0.000512179257309275)
    xp = max(min(v3, 0.9999999999), 1e-10)
    cns = max(min(v4, 24983.261744011077), 0.000249832619938437)
    xn = max(min(v5, 0.9999999999), 1e-10)
    th = np.tanh
    eta_p = 0.05138515824298745 * np.arcsinh((-2.3508116177110145 * cur) / (2.0 * ((1.8973665961010275e-05 * cps ** 0.5) * (51217.9257309275 - cps) ** 0.5)))
    up = (2.16216 + 0.07645 * th(30.834 - 57.858397200000006 * xp) + 2.1581 * th(52.294 - 53.412228 * xp)
          - 0.14169 * th(11.0923 - 21.0852666 * xp) + 0.2051 * th(1.4684 - 5.829105600000001 * xp)
          + 0.2531 * th(4.291641337386018 - 8.069908814589667 * xp) - 0.02167 * th(-87.5 + 177.0 * xp)
          + 1e-06 * ((1.0 / xp) + (1.0 / (-1.0 + xp))))
    eta_n = 0.05138515824298745 * np.arcsinh((1.9590096814258458 * cur) / (2.0 * ((0.0006324555320336759 * cns ** 0.5) * (24983.2619938437 - cns) ** 0.5)))
    un = (0.194 + 1.5 * np.exp(-120.0 * xn) + 0.0351 * th(-3.44578313253012 + 12.048192771084336 * xn)
          - 0.0045 * th(-7.1344537815126055 + 8.403361344537815 * xn) - 0.035 * th(-18.466 + 20.0 * xn)
          - 0.0147 * th(-14.705882352941176 + 29.41176470588235 * xn) - 0.102 * th(-1.3661971830985917 + 7.042253521126761 * xn)
          - 0.022 * th(-54.8780487804878 + 60.975609756097555 * xn) - 0.011 * th(-5.486725663716814 + 44.24778761061947 * xn)
          + 0.0155 * th(-3.6206896551724133 + 34.48275862068965 * xn) + 1e-06 * ((1.0 / xn) + (1.0 / (-1.0 + xn))))
    return (eta_p + up) - (eta_n + un)


def test_shared_elementary_functions(oracle):
    """dsb_exp / dsb_log / dsb_tanh / dsb_asinh (csrc/dsb_math.h) against libm over the ranges the battery model uses."""
    x = np.concatenate([np.linspace(-120.0, 60.0, 2001), np.array([-700.0, 700.0, 1e-9, -1e-9, 0.0])])
    assert np.allclose(oracle.math_fn("exp", x), np.exp(x), rtol=4e-16, atol=0.0)
    xl = np.concatenate([np.logspace(-300, 300, 1201), np.linspace(0.5, 2.0, 997)])
    assert np.allclose(oracle.math_fn("log", xl), np.log(xl), rtol=4e-16, atol=2e-16)
    xt = np.linspace(-90.0, 90.0, 4001)
    assert np.abs(oracle.math_fn("tanh", xt) - np.tanh(xt)).max() < 4e-16
    xa = np.concatenate([np.linspace(-50.0, 50.0, 2001), np.logspace(-6, 8, 300)])
    assert np.allclose(oracle.math_fn("asinh", xa), np.arcsinh(xa), rtol=1e-15, atol=4e-16)


def test_spm_stop_function_matches_the_model_text(oracle):
    """`stop_i` of spm.ds restated in csrc/dsb_models.h (ModelSpmStopT::root) against the same formula evaluated with
    numpy: the two root functions are V - 3.105 and 4.1 - V."""
    rng = np.random.default_rng(7)
    for _ in range(200):
        y = np.zeros(42)
        y[2:22] = rng.uniform(0.05, 0.95)
        y[22:42] = rng.uniform(0.05, 0.95)
        y[20:22] += rng.uniform(-0.02, 0.02, 2)
        y[40:42] += rng.uniform(-0.02, 0.02, 2)
        cur = rng.uniform(0.6, 1.4)
        g = oracle.model_root("spm_stop", y, [cur])
        v = spm_voltage_numpy(y[20], y[21], y[40], y[41], cur)
        assert abs(g[0] - (-3.105 + v)) < 1e-13 and abs(g[1] - (4.1 - v)) < 1e-13
    # the fresh cell of the model text (stoichiometries 0.8 / 0.6) sits inside the voltage window
    y0 = np.concatenate([[0.0, 0.0], np.full(20, 0.8000000000000016), np.full(20, 0.6000000000000001)])
    g0 = oracle.model_root("spm_stop", y0, [1.0])
    assert g0[0] > 0 and g0[1] > 0 and 3.5 < 3.105 + g0[0] < 4.1


def test_spm_discharge_ends_at_the_lower_cut_off(oracle):
    """The reference's battery example (examples/physics-based-battery-simulation/src/main.rs: currents 0.6 .. 1.4 A,
    3600 s): all but the smallest current end on the 3.105 V cut-off, at a discharged capacity of about 0.68 A h."""
    cur = np.array([[0.6], [0.8], [1.0], [1.2], [1.4]])
    t_eval = np.arange(1, 1201) * 3.0
    desc = oracle.make_desc("spm_stop")
    ys, stats, status, t_root, root_idx, ncols = oracle.batch_solve_dense_roots(desc, cur, t_eval)
    assert (status == 0).all()
    assert root_idx.tolist() == [-1, 0, 0, 0, 0] and ncols[0] == len(t_eval)
    capacity = cur[1:, 0] * t_root[1:] / 3600.0
    assert np.all((capacity > 0.66) & (capacity < 0.69))
    # solve_dense returns the model's output function (terminal voltage, one row): a discharge curve that ends at 3.105 V
    assert ys.shape == (5, len(t_eval), 1)
    v = ys[:, :, 0]
    assert np.all(np.diff(v[0]) < 0) and 3.7 < v[0, 0] < 3.8 and v[0, -1] > 3.105
    at_root = v[np.arange(1, 5), ncols[1:] - 1]
    assert np.abs(at_root - 3.105).max() < 1e-6                  # V = 3.105 at the root, on the interpolant
    assert np.isnan(v[1, ncols[1]:]).all()


@pytest.mark.parametrize("method", ["bdf", "tr_bdf2", "esdirk34"])
def test_solve_dense_with_reset(oracle, method):
    """ode_solver/mod.rs:1302-1372 test_solve_dense_with_reset (bdf.rs test_solve_dense_with_reset_bdf,
    sdirk.rs:1114-1122 test_solve_dense_with_reset_tr_bdf2) on exponential_decay_with_reset_problem
    (test_models/exponential_decay.rs:818-880: roots y[0] - 0.6 and y[0] - 0.3, reset y -> 0.4): solve_dense applies the
    resets and runs on to the last evaluation time; the state just before the second event is 0.3, just after 0.4."""
    k = 0.1
    t_root0 = -np.log(0.6) / k
    t_stop = t_root0 + np.log(4.0 / 3.0) / k              # second event: 0.4 decays to 0.3
    final_time = 2.0 * t_stop
    dt = 1e-3                                             # the reference probes the exact event time with solve(); here: just around it
    t_eval = np.array([1e-12, t_stop - dt, t_stop + dt, final_time])
    desc = oracle.make_desc("exp_decay_reset", method=method)            # builder defaults, libm pow
    ys, stats, status, t_root, root_idx, ncols = oracle.batch_solve_dense_roots(desc, [[k, 1.0]], t_eval)
    assert status[0] == 0 and root_idx[0] == -1 and ncols[0] == len(t_eval)       # TstopReached, every column filled
    pre = np.full(2, 0.4 * np.exp(-k * (t_eval[1] - t_root0)))
    post = np.full(2, 0.4 * np.exp(-k * (t_eval[2] - t_stop)))
    assert weighted_norm(ys[0, 1], pre, 1e-6, 1e-6) < 20.0
    assert weighted_norm(ys[0, 2], post, 1e-6, 1e-6) < 20.0
    assert abs(ys[0, 1, 0] - 0.3) < 1e-4 and abs(ys[0, 2, 0] - 0.4) < 1e-4
    # the state keeps cycling between 0.4 and 0.3 until the final time
    assert 0.3 <= ys[0, 3, 0] <= 0.4


BALL_BOUNCE_BDF = {            # bdf.rs:2691-2697: the three steps after the first bounce
    "x": [0.003978879413779122, 0.007955671343150521, 0.015904102550507716],
    "v": [11.202229406994425, 11.198746111635534, 11.191779520917754],
    "t": [1.4281779078441663, 1.4285126937676944, 1.4292157442071036],
}
BALL_BOUNCE_TR_BDF2 = {"x": [6.375884661615263], "v": [0.6878538646461059], "t": [2.5]}     # sdirk.rs:1084-1086


@pytest.mark.parametrize("method,expected", [("bdf", BALL_BOUNCE_BDF), ("tr_bdf2", BALL_BOUNCE_TR_BDF2)])
def test_ball_bounce_known_answers(oracle, method, expected):
    """test_ball_bounce (ode_solver/mod.rs:1024-1080) with the known answers of bdf.rs:2683-2712 and sdirk.rs:1076-1092:
    g = 9.81, drop height 10, restitution 0.8, stop time 2.5.  The test's state update at the root (v <- -e v,
    x <- max(x, eps), dy[0] <- v) is the model's reset function here; the reference then takes up to three internal
    steps and compares (x, v, t) after each with 1e-4.  BDF restarts at first order with small steps; TR-BDF2 keeps its
    step size and reaches the stop time with the first step after the bounce."""
    desc = oracle.make_desc("ball_bounce", method=method)           # builder defaults: rtol = atol = 1e-6, h0 = 1
    rc, t, y = oracle.steps_after_first_root(desc, [9.81, 10.0, 0.8], 2.5, 3)
    assert rc == 0 and len(t) == len(expected["t"])
    assert np.abs(t - expected["t"]).max() < 1e-4                   # the reference's tolerance
    # x and v after each step reproduce the reference's recorded values to the last digit
    assert np.abs(y[:, 0] - expected["x"]).max() < 1e-14
    assert np.abs(y[:, 1] - expected["v"]).max() < 1e-14


@pytest.mark.parametrize("method", ["bdf", "tr_bdf2", "esdirk34"])
def test_ball_bounce_solve_dense(oracle, method):
    """The same problem through solve_dense with the reset function: the ball keeps bouncing (energy loss 1 - e^2 per
    bounce) and the trajectory follows the closed-form piecewise parabola."""
    g, h0, e = 9.81, 10.0, 0.8
    t_eval = np.linspace(0.05, 6.0, 120)
    desc = oracle.make_desc("ball_bounce", method=method)
    ys, stats, status, t_root, root_idx, ncols = oracle.batch_solve_dense_roots(desc, [[g, h0, e]], t_eval)
    assert status[0] == 0 and root_idx[0] == -1 and ncols[0] == len(t_eval)
    tb, v0, exact = np.sqrt(2 * h0 / g), 0.0, []
    t_start, x_start = 0.0, h0
    for t in t_eval:
        while True:
            t_land = t_start + (v0 + np.sqrt(v0 * v0 + 2 * g * x_start)) / g
            if t <= t_land:
                break
            v0 = e * (g * (t_land - t_start) - v0)
            t_start, x_start = t_land, 0.0
        exact.append(x_start + v0 * (t - t_start) - 0.5 * g * (t - t_start) ** 2)
    assert np.abs(ys[0, :, 0] - np.array(exact)).max() < 2e-3
    assert t_eval[-1] > tb + 2 * e * np.sqrt(2 * h0 / g)          # more than one bounce inside the window


@pytest.mark.parametrize("method", ["bdf", "tr_bdf2"])
def test_root_found_index(oracle, method):
    """test_root_found_index (ode_solver/mod.rs:1187-1220; bdf.rs:2725-2731, sdirk.rs:1096-1103) on
    exponential_decay_with_two_roots_problem: root 0 (y[0] = 0.6) fires at t = -ln(0.6) / 0.1 within 1e-4, and the solve
    reports index 0.  Started below 0.6 the second root function (y[0] = 0.3) fires instead, with index 1."""
    desc = oracle.make_desc("exp_decay_two_roots", method=method)
    t_eval = np.array([50.0, 100.0])                      # set_stop_time(100)
    ys, stats, status, t_root, root_idx, ncols = oracle.batch_solve_dense_roots(desc, [[0.1, 1.0], [0.1, 0.5]], t_eval)
    assert (status == 0).all() and root_idx.tolist() == [0, 1] and ncols.tolist() == [1, 1]
    assert abs(t_root[0] - (-np.log(0.6) / 0.1)) < 1e-4
    assert abs(t_root[1] - (-np.log(0.3 / 0.5) / 0.1)) < 1e-4
    assert abs(ys[0, 0, 0] - 0.6) < 1e-5 and abs(ys[1, 0, 0] - 0.3) < 1e-5


@pytest.mark.parametrize("method", ["bdf", "tr_bdf2"])
def test_tstop(oracle, method):
    """test_tstop_bdf (bdf.rs:2490-2494) / test_tstop_tr_bdf2: test_ode_solver(.., use_tstop = true) on the exponential
    decay problem -- set_stop_time(point), step until TstopReached, state().y against the analytic solution with the
    harness's criterion (ode_solver/mod.rs:164-173: weighted error norm < 15)."""
    k, y0 = 0.1, 1.0
    pts = np.arange(0.0, 10.0)[1:]                        # a stop time at the current time is an error (t = 0)
    desc = oracle.make_desc("exp_decay", method=method)
    rc, ys, stats, fin = oracle.harness(desc, [k, y0], pts, use_tstop=True)
    assert rc == 0 and abs(fin["t"] - pts[-1]) < 1e-12
    for t, y in zip(pts, ys):
        exact = np.full(2, y0 * np.exp(-k * t))
        assert weighted_norm(y, exact, 1e-6, 1e-6) < 15.0
