"""CPU suite, part 1: the oracle itself — tripwires of SURVEY.md §8(c)/Appendix A, the reference's own
known-answer tests (re-expressed headless), gradient cross-checks and the golden fixtures."""
import os

import numpy as np
import pytest

from oracle import pyoracle as po

P = po.nn_default_params()
FLAT = po.terrain(po.FLAT)


def test_tire_network_spot_values():
    # SURVEY.md §8(c) spot values (independent numpy scratch restatement made at survey time)
    assert po.tire_nn_forward(P, [0, 0, 0, 1e-3, 0, 0, 0, 0])[2] == pytest.approx(10.70978336, rel=1e-9)
    assert po.tire_nn_forward(P, [0, 0, 0, 3e-3, 0, 0, 0, 0])[2] == pytest.approx(42.69948819, rel=1e-9)
    assert np.all(po.tire_nn_forward(P, [0, 0, 0, -1e-3, 0, 0, 0, 0]) == 0.0)  # relu branch, xy(0)=0 without biases


def test_ode_spot_value():
    X = np.array([0, 0, 0, 1, 0, 0, 0.0605, 0, 0, 0, 0, 0, 0, 0.1, 0.3, 0, 0, 4, 3, 4, 3.0])
    xd = po.ode(po.TIRE_NN, P, FLAT, X, [4, 3])
    np.testing.assert_allclose(xd[11:17], [-0.65026408, 6.45993003, -1.15761040, -2.13477766, -0.58326425, -0.47412467], rtol=2e-8)
    np.testing.assert_allclose(xd[7:11], [4, 3, 4, 3])
    assert np.all(xd[17:21] == 0)


def test_whole_body_com_and_mass():
    np.testing.assert_allclose(po.whole_body_com(), [0.010757161375, 0.0, 0.063632110082], rtol=1e-9, atol=1e-15)


def test_settle_tripwire():
    # HybridDynamics::settle (HybridDynamics.cpp:142-190): SURVEY.md §8(c)
    s = po.settle(po.TIRE_NN, P, FLAT)
    np.testing.assert_allclose(s[:4], [8.6383e-05, 9.57191e-03, -5.7652e-04, 0.99995402], rtol=2e-5)
    assert s[6] == pytest.approx(0.0584427891, rel=1e-8)


def test_param_roundtrip():
    # TireNetwork.check_params (test/test_TireNetwork.cpp:54-71): 416 params, 4 identical nets of 104
    assert P.shape == (416,)
    for k in range(1, 4):
        assert np.array_equal(P[:104], P[104 * k:104 * (k + 1)])


def _yaw(x):
    qx, qy, qz, qw = x[:4]
    return np.arctan2(2 * (qw * qz + qx * qy), 1 - 2 * (qy * qy + qz * qz))


def test_kat_test_wz():
    # VehicleFixture.test_wz (test/test_Vehicle.cpp:275-374): zero NN params, free fall, wz reset to 10 every step
    x = np.zeros(23)
    x[3], x[6] = 1, 100
    for _ in range(100):
        x[11:17] = [0, 0, 10, 0, 0, 0]
        x[21] = x[22] = 0
        x = po.integrate(po.TIRE_NN, np.zeros(416), FLAT, x)
    final = np.arctan2(np.sin(10.0), np.cos(10.0))
    assert abs(_yaw(x) - final) < 1e-3


def test_kat_no_lateral_drift_and_straight():
    # VehicleFixture.no_lateral_drift (test_Vehicle.cpp:110-154) and .straight (:157-200) with z0 = .0605
    x0 = np.zeros(23)
    x0[:4] = po.settle(po.TIRE_NN, P, FLAT)[:4]
    x0[6] = .0605
    x = x0.copy()
    x[15] = .1
    for _ in range(1000):
        x[21] = x[22] = 0
        x = po.integrate(po.TIRE_NN, P, FLAT, x)
    assert 0.0 < x[5] < 0.1
    x = x0.copy()
    for _ in range(1000):
        x[21] = x[22] = 4.0
        x = po.integrate(po.TIRE_NN, P, FLAT, x)
    # the reference asserts 0.98 +- 0.05 but only with the author's private tire.net; with load_model()
    # weights the vehicle travels ~1.76 m — pin the sign/magnitude only
    assert x[4] == pytest.approx(1.76, abs=0.02)  # survey-time scratch restatement: 1.76 m


def test_kat_bekker_drives():
    # BekkerFixture.no_lateral_drift / straight (test/test_Bekker.cpp:399-495), z0 = .0605
    pb = po.bekker_default_params()
    x0 = np.zeros(23)
    x0[3], x0[6] = 1.0, .0605
    x = x0.copy()
    x[15] = .1
    for _ in range(300):
        x[21] = x[22] = 0
        x = po.integrate(po.TIRE_BEKKER, pb, FLAT, x)
    assert 0.0 < x[5] < 0.1
    x = x0.copy()
    for _ in range(1000):
        x[21] = x[22] = 4.0
        x = po.integrate(po.TIRE_BEKKER, pb, FLAT, x)
    assert x[4] == pytest.approx(3.92, abs=0.1)  # test_Bekker.cpp:492


def test_bumpy_terrain_is_a_window():
    t = po.terrain(po.BUMPY)
    assert po.altitude(t, 1.9, 0) == 0.0
    assert po.altitude(t, 2.8, 0) == pytest.approx(0.2 * np.sin(0.8))
    assert po.altitude(t, 3.7, 5) == 0.0  # drops back to 0 for x >= 3.6 (TestTerrainMaps.cpp:22-23)
    assert po.altitude(po.terrain(po.SLOPE), 2.0, 1.0) == pytest.approx(0.2)


def test_loss_and_evaluate():
    gt = np.zeros(21)
    gt[3], gt[4], gt[5] = 0.3, 1.0, 2.0
    x = np.zeros(23)
    x[2], x[3] = np.sin(0.1), np.cos(0.1)  # yaw = 0.2
    x[4], x[5] = 1.3, 2.4
    assert po.loss(gt, x) == pytest.approx(0.1 + 0.5, rel=1e-12)
    ang, lin = po.evaluate(gt, x)
    assert lin == pytest.approx(0.25, rel=1e-12)


def test_rmsprop_float_truncation():
    # Trainer::updateParams (Trainer.cpp:428-453): gradient cast to float before use
    p, s = po.rmsprop(np.array([1.0]), np.array([1.0]), np.array([0.1 + 1e-12]))
    g = np.float32(0.1 + 1e-12)
    s_ref = 0.95 + 0.05 * float(g * g)
    assert s[0] == s_ref
    assert p[0] == 1.0 - (1e-3 / (np.sqrt(s_ref) + 1e-6)) * float(g)


def _rows(T, seed):
    rng = np.random.default_rng(seed)
    rows = np.zeros((T, 10))
    rows[:, 0] = 0.01 * np.arange(T)
    rows[:, 1] = rng.uniform(1, 6, T)
    rows[:, 2] = rng.uniform(1, 6, T)
    rows[:, 3] = np.cumsum(rng.normal(0.004, 0.001, T))
    rows[:, 4] = np.cumsum(rng.normal(0.0, 0.001, T))
    rows[:, 5] = 0.0584427891
    rows[:, 6] = 0.4 + np.cumsum(rng.normal(0, 0.003, T))
    rows[:, 7], rows[:, 8], rows[:, 9] = 0.1, 0.4, 0.01
    return rows


def test_gradient_three_ways():
    # (1) per-step tape == (2) monolithic tape (the reference's structure) == (3) forward-mode duals
    rows = _rows(5, 0)
    L1, g1, _ = po.train_trajectory(P, FLAT, rows)
    L2, g2 = po.train_trajectory_monolithic(P, FLAT, rows)
    assert L1 == pytest.approx(L2, rel=1e-13)
    np.testing.assert_allclose(g1, g2, rtol=1e-9, atol=1e-13 * np.abs(g2).max())
    assert np.all(g1[104:] == 0.0)  # networks 1..3 are never evaluated (HybridDynamics.cpp:387)
    idx = [0, 5, 23, 24, 60, 87, 88, 103]
    L3, d = po.train_trajectory_dual8(P, FLAT, rows, idx)
    assert L3 == pytest.approx(L1, rel=1e-13)
    np.testing.assert_allclose(d, g1[idx], rtol=1e-8, atol=1e-12 * np.abs(g1).max())


def test_gradient_vs_central_differences():
    rows = _rows(4, 1)
    L, g, _ = po.train_trajectory(P, FLAT, rows)
    for k in (1, 30, 90):
        h = 1e-6
        pp, pm = P.copy(), P.copy()
        pp[k] += h
        pm[k] -= h
        fd = (po.train_trajectory(pp, FLAT, rows)[0] - po.train_trajectory(pm, FLAT, rows)[0]) / (2 * h)
        assert fd == pytest.approx(g[k], rel=2e-5, abs=1e-9)


def test_golden_fixtures_reproduce(golden_dir):
    G = np.load(os.path.join(golden_dir, "golden_v1.npz"))
    for terr in (po.FLAT, po.SLOPE, po.BUMPY):
        t = po.terrain(terr)
        for i in (0, 7, 31):
            np.testing.assert_array_equal(po.ode(po.TIRE_NN, P, t, G[f"ode_nn_X_{terr}"][i], G[f"ode_nn_U_{terr}"][i]), G[f"ode_nn_Xd_{terr}"][i])
    xl, _ = po.rollout_batch(po.TIRE_NN, P, FLAT, G["roll60_x0"], G["roll60_ctrl"], 60)
    np.testing.assert_array_equal(xl, G["roll60_xlist"])
    loss, gs, ls, nk, gps = po.rollout_grad_batch(P, FLAT, G["grad20_x0"], G["grad20_ctrl"], G["grad20_gt"], 20, nthreads=2)
    np.testing.assert_array_equal(loss, G["grad20_loss"])
    np.testing.assert_array_equal(gps, G["grad20_gps"])
    np.testing.assert_allclose(po.settle(po.TIRE_NN, P, FLAT), G["settle_nn"], rtol=0, atol=0)


def test_explosion_filter_modes():
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_v1.npz"))
    x0, ctrl, gt = G["grad20_x0"], G["grad20_ctrl"], G["grad20_gt"]
    gps = G["grad20_gps"]
    thresh = float(np.sort(np.abs(gps).max(axis=0))[2])  # makes exactly the two "largest" samples explode somewhere
    _, gs0, ls0, nk0, _ = po.rollout_grad_batch(P, FLAT, x0, ctrl, gt, 20, explode_thresh=thresh * 0.999, explode_mode=0, nthreads=2)
    _, gs1, ls1, nk1, _ = po.rollout_grad_batch(P, FLAT, x0, ctrl, gt, 20, explode_thresh=thresh * 0.999, explode_mode=1, nthreads=2)
    assert nk0 == 4 and nk1 == 2
    ref0 = np.where(np.abs(gps) > thresh * 0.999, 0.0, gps).sum(axis=1)
    np.testing.assert_allclose(gs0, ref0, rtol=1e-13, atol=1e-18)
    keep = np.abs(gps).max(axis=0) <= thresh * 0.999
    np.testing.assert_allclose(gs1, gps[:, keep].sum(axis=1), rtol=1e-13, atol=1e-18)


def test_bekker_reverse_sweep_is_not_finite():
    """SURVEY.md §8 row (f3), Bekker adjoint: with CppAD's conventions the reverse sweep through one
    BekkerSystem::integrate is NOT finite in the ordinary ze == zr regime — theta_c = acos(1 - (zr - ze)/R)
    (BekkerTireModel.cpp:232-238) is differentiated at acos(1).  This is what the reference's notes report
    ("You can't train the bekker model", todo.org:60; "the param -> nan", todo.org:1088) and why the
    CUDA path declines rollout_grad for the Bekker tire model instead of inventing a gradient."""
    p = po.bekker_default_params()
    st = po.settle(po.TIRE_BEKKER, p, FLAT)
    x = np.zeros(23)
    x[:21] = st
    x[17:21] = 0.0
    x[21], x[22] = 2.0, 3.0
    lam = np.zeros(21)
    lam[4], lam[5] = 1.0, 0.5
    x1, g, xb, bad = po.bekker_integrate_vjp(p, FLAT, x, lam)
    assert np.all(np.isfinite(x1))        # the forward values are fine
    assert bad > 0 and not np.all(np.isfinite(g))
    assert g[3] == 0.0                    # CondExpGt(p3, 0, p3, 0) at the default p3 = 0 selects the constant branch
