"""CPU suite, part 2: the device math (auvsl_neural_ode_b200/csrc/avs_model.cuh + avs_rollout.cuh, the
templates the sm_100a kernels are built from) compiled by g++ (tests/hostcheck) and checked against the
oracle: constant-folded ODE / RK4, the hand-written adjoint, rollouts, Bekker model, grid terrain."""
import os

import numpy as np
import pytest

import hostcheck_lib as hc
from oracle import pyoracle as po

P = po.nn_default_params()
LIVE = hc.LIVE


def _golden():
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_v1.npz"))


def test_default_weights_match_reference_data():
    assert np.array_equal(hc.default_params(), P)


def test_tanh_accuracy():
    rng = np.random.default_rng(0)
    xs = np.concatenate([rng.normal(0, 3, 4000), rng.normal(0, 0.01, 1000), [0.0, 1e-300, 19.9, 30.0, -40.0, 400.0, -1e4]])
    err = max(abs(hc.lib().hc_tanh(float(x)) - np.tanh(x)) for x in xs)
    assert err <= 4.5e-16
    # the 8-wide tanh of the kernels' hot path (tanh_vec: -2|x| folded into the polynomial, one-word reduction, 1 - 2e/(1+e))
    err_vec = max(abs(hc.lib().hc_tanh_vec(float(x)) - np.tanh(x)) for x in xs)
    print(f"tanh_vec: max abs error {err_vec:.2e}")
    assert err_vec <= 4.5e-16


def test_ztable_matches_fixed_z_network():
    """the constant-folded z-branch (piecewise degree-9 table, Estrin evaluation, stored derivative polynomial) against the
    network it replaces (TireNetwork.cpp:17-32, 60-71) evaluated in extended precision: value and derivative"""
    zw = hc.zweights().astype(np.longdouble)
    Z0, Z2, Z4 = zw[:8], zw[8:72].reshape(8, 8), zw[72:80]

    def znet(s):
        g0 = np.tanh(Z0 * s)
        g2 = np.tanh(Z2 @ g0)
        dz = Z4 @ ((1 - g2 * g2) * (Z2 @ ((1 - g0 * g0) * Z0)))
        return Z4 @ g2, dz

    rng = np.random.default_rng(5)
    ss = np.concatenate([rng.uniform(0, 2, 3000), rng.uniform(0, 52, 2000), np.arange(0, 833) / 16.0, np.arange(0, 833) / 16.0 + 1e-13,
                         [1e-300, 1e-9, 51.99999, 52.0, 52.5, 1e3]])
    scale = max(abs(float(znet(np.longdouble(s))[0])) for s in (0.5, 5.0, 52.0))
    ev = ed = 0.0
    for s in ss:
        v, d = hc.ztab(s)
        rv, rd = znet(np.longdouble(s))
        ev = max(ev, abs(v - float(rv)) / scale)
        if s > 0.0:  # at s0 <= 0 the relu gate is closed: value 0, derivative 0 (CondExpGt)
            ed = max(ed, abs(d - float(rd)) / scale)
    assert ev < 4e-16 and ed < 2e-13, (ev, ed)


def test_aiinv_is_inverse_of_assembled_inertia():
    # SURVEY.md Appendix A3 tripwires for AI_b
    A = np.linalg.inv(hc.aiinv())
    np.testing.assert_allclose(np.diag(A), [0.4625982159472, 0.5037652562592, 0.5561779557127, 18.43200016022, 18.43200016022, 18.43200016022], rtol=1e-9)
    assert A[0, 4] == pytest.approx(-1.172867063224, rel=1e-9)
    assert A[1, 3] == pytest.approx(1.172867063224, rel=1e-9)
    assert A[1, 5] == pytest.approx(-0.1982760001817, rel=1e-9)


@pytest.mark.parametrize("terr", [po.FLAT, po.SLOPE, po.BUMPY])
def test_ode_and_rk4_vs_golden(terr):
    G = _golden()
    m = hc.model(0, terr, P)
    X, U, Xd, X1 = G[f"ode_nn_X_{terr}"], G[f"ode_nn_U_{terr}"], G[f"ode_nn_Xd_{terr}"], G[f"rk4_nn_X1_{terr}"]
    for i in range(X.shape[0]):
        a = hc.ode(m, X[i][LIVE], U[i])
        np.testing.assert_allclose(a, Xd[i][LIVE], rtol=1e-10, atol=1e-11)
        b, _ = hc.rk4(m, X[i][LIVE], U[i])
        np.testing.assert_allclose(b, X1[i][LIVE], rtol=1e-13, atol=1e-14)


def test_bekker_ode_vs_golden():
    G = _golden()
    m = hc.model(1, po.FLAT, po.bekker_default_params())
    X, U, Xd = G["ode_bk_X"], G["ode_bk_U"], G["ode_bk_Xd"]
    for i in range(X.shape[0]):
        a = hc.ode(m, X[i][LIVE], U[i], X[i][17:21])
        np.testing.assert_allclose(a, Xd[i][LIVE], rtol=1e-9, atol=1e-9)


def test_rollout_vs_golden():
    G = _golden()
    m = hc.model(0, po.FLAT, P)
    xl, xf = hc.rollout(m, G["roll60_x0"], G["roll60_ctrl"], 60)
    np.testing.assert_allclose(xl, G["roll60_xlist"], rtol=1e-11, atol=1e-13)
    np.testing.assert_allclose(xf, G["roll60_xfinal"], rtol=1e-11, atol=1e-13)
    mb = hc.model(0, po.BUMPY, P)
    xl, _ = hc.rollout(mb, G["bumpy_x0"], G["bumpy_ctrl"], 120)
    np.testing.assert_allclose(xl, G["bumpy_xlist"], rtol=1e-10, atol=1e-12)


def test_bekker_rollout_vs_golden():
    G = _golden()
    m = hc.model(1, po.FLAT, po.bekker_default_params())
    xl, _ = hc.rollout(m, G["bk_x0"], G["bk_ctrl"], 100)
    err = float(np.max(np.abs(xl - G["bk_xlist"]) / np.maximum(1.0, np.abs(G["bk_xlist"]))))
    print(f"bekker rollout (device math on CPU) vs golden: {err:.3e}")
    assert err < 1e-9  # north star: states within 1e-9


def test_bekker_branch_margins_and_long_rollout_parity():
    """per-trajectory branch margins are measured by the device math AND by the oracle; they agree, and every
    margin-safe trajectory holds 1e-9 over a 600-row rollout (5990 RK4 steps) on the randomised height / soil grid."""
    from auvsl_neural_ode_b200 import synth
    g = synth.make_grid_terrain(n=128)
    N, T = 6, 600
    x0, ctrl, _ = synth.make_batch(N, T, z_stable=0.065, seed=11)
    ctrl = np.clip(ctrl, 0.5, 8.0)
    pb = po.bekker_default_params()
    t = po.terrain(po.GRID, height=g["height"], soil=g["soil"], x0=g["x0"], y0=g["y0"], dx=g["dx"], dy=g["dy"])
    xl_o, _, mg_o = po.rollout_batch_margins(po.TIRE_BEKKER, pb, t, x0, ctrl, T, nthreads=po.hardware_threads())
    m = hc.model(1, po.GRID, pb, height=g["height"], soil=g["soil"], x0=g["x0"], y0=g["y0"], dx=g["dx"], dy=g["dy"])
    xl, _, mg = hc.rollout_margins(m, x0, ctrl, T)
    assert mg.shape == (6, N) and np.all(mg > 0) and np.all(mg < 1e300)
    big = np.minimum(mg, mg_o) > 1e-6
    np.testing.assert_allclose(mg[big], mg_o[big], rtol=1e-6)
    per = (np.abs(xl - xl_o) / np.maximum(1.0, np.abs(xl_o))).reshape(-1, N).max(axis=0)
    safe = np.minimum(mg.min(axis=0), mg_o.min(axis=0)) > 1e-10
    print(f"bekker T=600 on grid: {int(safe.sum())}/{N} margin-safe, worst {per[safe].max():.3e}; min margin {np.minimum(mg, mg_o).min():.3e}")
    assert safe.sum() >= N // 2
    assert per[safe].max() < 1e-9


def test_hand_adjoint_vs_tape():
    G = _golden()
    m = hc.model(0, po.FLAT, P)
    loss, gps = hc.rollout_grad(m, G["grad20_x0"], G["grad20_ctrl"], G["grad20_gt"], 20)
    np.testing.assert_allclose(loss, G["grad20_loss"], rtol=1e-11)
    ref = G["grad20_gps"][:104]
    # gradient tolerance of the north star: 1e-7 relative; here ~1e-13 of the gradient scale
    np.testing.assert_allclose(gps, ref, rtol=1e-7, atol=1e-12 * np.abs(ref).max())
    assert np.all(G["grad20_gps"][104:] == 0)


def test_hand_adjoint_T200():
    G = _golden()
    m = hc.model(0, po.FLAT, P)
    loss, gps = hc.rollout_grad(m, G["grad200_x0"], G["grad200_ctrl"], G["grad200_gt"], 200)
    np.testing.assert_allclose(loss, G["grad200_loss"], rtol=1e-10)
    ref = G["grad200_gps"][:104]
    np.testing.assert_allclose(gps, ref, rtol=1e-7, atol=1e-10 * np.abs(ref).max())


def test_ode_vjp_matches_finite_differences():
    rng = np.random.default_rng(3)
    G = _golden()
    m = hc.model(0, po.SLOPE, P)
    X, u = G["ode_nn_X_1"][3][LIVE], G["ode_nn_U_1"][3]
    lam = rng.normal(size=13)
    xb, gW = hc.ode_vjp(m, X, u, lam)
    for c in range(13):
        h = 1e-7
        xp, xm = X.copy(), X.copy()
        xp[c] += h
        xm[c] -= h
        fd = lam @ (hc.ode(m, xp, u) - hc.ode(m, xm, u)) / (2 * h)
        assert fd == pytest.approx(xb[c], rel=2e-5, abs=2e-5 * np.abs(xb).max())


def test_grid_terrain_and_soil_vs_oracle():
    rng = np.random.default_rng(5)
    n, cell = 40, 0.05
    x0g = y0g = -0.5 * (n - 1) * cell
    xs = x0g + cell * np.arange(n)
    XX, YY = np.meshgrid(xs, xs)
    h = 0.01 * np.sin(3 * XX + 1.0) * np.cos(2 * YY)
    d = po.bekker_default_params()
    soil = np.stack([d[k] * rng.uniform(0.8, 1.2, (n, n)) if k != 3 else rng.uniform(0, 0.1, (n, n)) for k in range(5)])
    t = po.terrain(po.GRID, height=h, soil=soil, x0=x0g, y0=y0g, dx=cell, dy=cell)
    G = _golden()
    # NN on the grid
    m = hc.model(0, po.GRID, P, height=h, x0=x0g, y0=y0g, dx=cell, dy=cell)
    for i in range(8):
        X = G["ode_nn_X_0"][i].copy()
        X[4:6] = rng.uniform(-0.6, 0.6, 2)
        a = hc.ode(m, X[LIVE], G["ode_nn_U_0"][i])
        b = po.ode(po.TIRE_NN, P, t, X, G["ode_nn_U_0"][i])[LIVE]
        np.testing.assert_allclose(a, b, rtol=1e-10, atol=1e-10)
    # Bekker with per-cell soil
    mb = hc.model(1, po.GRID, d, height=h, soil=soil, x0=x0g, y0=y0g, dx=cell, dy=cell)
    for i in range(8):
        X = G["ode_bk_X"][i].copy()
        X[4:6] = rng.uniform(-0.6, 0.6, 2)
        a = hc.ode(mb, X[LIVE], G["ode_bk_U"][i], X[17:21])
        b = po.ode(po.TIRE_BEKKER, d, t, X, G["ode_bk_U"][i])[LIVE]
        np.testing.assert_allclose(a, b, rtol=1e-9, atol=1e-9)
    # gradient through the height field (adjoint of the bilinear lookup)
    x0, ctrl, gt = G["grad20_x0"], G["grad20_ctrl"], G["grad20_gt"]
    lo, _, _, _, gps_o = po.rollout_grad_batch(P, t, x0, ctrl, gt, 20, nthreads=2)
    lh, gps_h = hc.rollout_grad(m, x0, ctrl, gt, 20)
    np.testing.assert_allclose(lh, lo, rtol=1e-10)
    np.testing.assert_allclose(gps_h, gps_o[:104], rtol=1e-7, atol=1e-11 * np.abs(gps_o).max())


def test_bekker_vectorised_quadrature_vs_oracle_random_inputs():
    """avs_bekker.cuh evaluates the front / rear arcs with its own branch-free log / exp / rsqrt and rotated sines;
    against the oracle's scalar walk (BekkerTireModel.cpp:163-278) over random sinkage, slip and soil — including
    soft soil, where the centre arc is active."""
    import ctypes as C
    rng = np.random.default_rng(1)
    p = po.bekker_default_params()
    worst = 0.0
    for it in range(1500):
        zr = 10 ** rng.uniform(-5, -1.0)
        vx, vy, w = rng.uniform(-1, 2), rng.uniform(-0.5, 0.5), rng.uniform(-5, 12)
        soil = p.copy()
        if it % 3 == 1:
            soil = p * rng.uniform(0.5, 1.5, 5)
            soil[3] = rng.uniform(0, 0.3)
        if it % 3 == 2:
            soil = p * np.array([0.02, 0.02, 1, 1, 1])
        m = hc.model(1, po.FLAT, soil)
        F = np.zeros(3)
        hc.lib().hc_bekker_tire(C.byref(m), C.c_double(vx), C.c_double(vy), C.c_double(w), C.c_double(zr), F.ctypes.data_as(C.POINTER(C.c_double)))
        feat_soil = np.array([100 * soil[0], 1000 * soil[1], soil[2], max(soil[3], 0), soil[4]])  # BekkerDynamics.cpp:59-63
        _, f = po.bekker_forces([zr, vx, vy, 0.098 * w], feat_soil)
        fx = abs(f[0]) if (0.098 * w - vx > 0) else -abs(f[0])                                     # BekkerDynamics.cpp:93
        ref = np.array([fx, f[1], f[2]])
        worst = max(worst, float(np.max(np.abs(F - ref) / np.maximum(1e-6, np.abs(ref).max()))))
    assert worst < 1e-10


def test_producer_adjoint_is_the_one_piece_adjoint():
    """The reverse kernel's producer warp runs ode_bwd_nn_p: the layered back-propagation of the tire network on the state chain,
    the operands of the three weight-gradient outer products handed to the consumer warp.  Against the one-piece adjoint
    (ode_bwd_nn / tire_nn_bwd, applied at once) on all terrains: state adjoint and weight gradient bit-identical."""
    rng = np.random.default_rng(11)
    G = _golden()
    for terr in (po.FLAT, po.SLOPE, po.BUMPY):
        m = hc.model(0, terr, P)
        for i in range(8):
            X, u = G[f"ode_nn_X_{terr}"][i][LIVE], G[f"ode_nn_U_{terr}"][i]
            lam = rng.normal(size=13)
            xb, gW = hc.ode_vjp(m, X, u, lam)
            xb3, gW3 = hc.ode_vjp_layered(m, X, u, lam)
            np.testing.assert_array_equal(xb3, xb)
            np.testing.assert_array_equal(gW3, gW)


def test_contact_search_matches_oracle():
    """HybridDynamics::get_tire_cpts_sinkages_fancy (10 candidate contact angles per wheel) as a forward
    option, device math vs oracle (whose function is itself pinned to the reference's in tests/test_ref_pinning.py): ODE
    values and a rollout over the 96^2 height / soil grid for both tire models, and Bumpy terrain for the NN model."""
    from auvsl_neural_ode_b200 import synth
    g = synth.make_grid_terrain(n=96)
    kw = dict(height=g["height"], soil=g["soil"], x0=g["x0"], y0=g["y0"], dx=g["dx"], dy=g["dy"])
    N, T = 5, 30
    x0, ctrl, _ = synth.make_batch(N, T, seed=41, z_stable=0.06)
    x0[4], x0[5] = np.linspace(-1.5, 1.5, N), np.linspace(1.0, -1.0, N)
    ctrl = np.clip(ctrl, 1, 8)
    hc.set_contact_search(True)
    try:
        for tire, params in ((0, P), (1, po.bekker_default_params())):
            t = po.terrain(po.GRID, contact_search=1, **kw)
            m = hc.model(tire, po.GRID, params, **kw)
            xl, _ = hc.rollout(m, x0, ctrl, T)
            xl_o, _ = po.rollout_batch(tire, params, t, x0, ctrl, T, nthreads=4)
            xl_plain, _ = po.rollout_batch(tire, params, po.terrain(po.GRID, **kw), x0, ctrl, T, nthreads=4)
            err = float(np.max(np.abs(xl - xl_o) / np.maximum(1.0, np.abs(xl_o))))
            print(f"contact search, tire {tire}: device math vs oracle {err:.2e}; effect of the option on the states {np.abs(xl_o - xl_plain).max():.2e}")
            assert err < 1e-9
            assert np.abs(xl_o - xl_plain).max() > 1e-6  # the option does change the result
        xb = x0.copy()
        xb[4] = np.linspace(1.9, 3.5, N)
        xl, _ = hc.rollout(hc.model(0, po.BUMPY, P), xb, ctrl, T)
        xl_o, _ = po.rollout_batch(0, P, po.terrain(po.BUMPY, contact_search=1), xb, ctrl, T, nthreads=4)
        assert float(np.max(np.abs(xl - xl_o) / np.maximum(1.0, np.abs(xl_o)))) < 1e-9
    finally:
        hc.set_contact_search(False)
