"""GPU suite (-m gpu): the CUDA path through the C ABI (libavsim.so) against the oracle on seeded inputs,
against the committed golden fixtures, and — at BASELINE.json's full sizes — through size-independent
properties.  Tolerances are the north star's: states 1e-9 relative, gradients 1e-7 relative.
Nothing here reads /root/reference."""
import os

import numpy as np
import pytest

from auvsl_neural_ode_b200 import capi, synth
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu

P = po.nn_default_params()
STATE_RTOL = 1e-9
GRAD_RTOL = 1e-7


def golden():
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_v1.npz"))


def golden_ref():
    """inputs / outputs of the reference's OWN code (tests/golden/make_golden_ref.py, oracle/ref_capi.cpp)"""
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_ref_v1.npz"))


def rel_state_err(a, b):
    """1e-9 'relative' for states: relative to max(1, |value|) (angles / positions near 0 are absolute)."""
    return float(np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b))))


MARGIN_GATE = 1e-10


def bekker_parity(tag, got, want, mg_gpu, mg_orc, tol=STATE_RTOL):
    """the Bekker path branches on values, so parity is asserted per trajectory where every branch stayed
    farther than MARGIN_GATE from flipping (margins measured on BOTH sides); the rest is counted and its worst error
    printed.  got / want: [..., N] arrays; margins [6][N].  Returns (n_safe, worst_safe, n_unsafe, worst_unsafe)."""
    N = got.shape[-1]
    err = np.abs(got - want) / np.maximum(1.0, np.abs(want))
    per = err.reshape(-1, N).max(axis=0)
    margin = np.minimum(mg_gpu.min(axis=0), mg_orc.min(axis=0))
    safe = margin > MARGIN_GATE
    worst_safe = float(per[safe].max()) if safe.any() else 0.0
    worst_unsafe = float(per[~safe].max()) if (~safe).any() else 0.0
    kinds = [capi.MARGIN_KINDS[int(k)] for k in np.argmin(np.minimum(mg_gpu, mg_orc), axis=0)[~safe]]
    print(f"[bekker parity] {tag}: {int(safe.sum())}/{N} trajectories margin-safe, worst rel_state_err {worst_safe:.3e} (bar {tol:g}); "
          f"{int((~safe).sum())} below the {MARGIN_GATE:g} gate, worst {worst_unsafe:.3e} {sorted(set(kinds))}; "
          f"smallest margin {float(margin.min()):.3e}")
    assert safe.sum() >= max(1, N // 2), "branch-margin gate rejected most trajectories: the test inputs are degenerate"
    assert worst_safe < tol
    # the two sides must agree on the margins themselves wherever they are not tiny (same predicates, same operands)
    big = np.minimum(mg_gpu, mg_orc) > 1e-6
    np.testing.assert_allclose(mg_gpu[big], mg_orc[big], rtol=1e-6)
    return int(safe.sum()), worst_safe, int((~safe).sum()), worst_unsafe


@pytest.fixture(scope="module")
def nn():
    c = capi.Context(0, capi.TIRE_NN)
    yield c
    c.close()


@pytest.fixture(scope="module")
def bk():
    c = capi.Context(0, capi.TIRE_BEKKER)
    yield c
    c.close()


def test_cuda_path_vs_reference_own_numbers(nn, bk):
    """The CUDA path against numbers produced by THE REFERENCE'S OWN CODE (first-party sources compiled in place behind stand-in
    third-party headers; fixture generated in the build container, where /root/reference exists): ODE values, every row of
    120-row rollouts on flat and bumpy terrain, final states after 600 rows, Trainer::trainTrajectory losses and gradients at
    T = 20 and T = 200, the RMSProp update, settle; Bekker ODE values and 100-row rollouts.  Bars: 1e-9 / 1e-7."""
    R = golden_ref()
    P_, PB_ = R["nn_params"], R["bekker_params"]
    assert np.array_equal(P_, capi.nn_default_params()) and np.array_equal(PB_, capi.bekker_default_params())
    nn.set_nn_params(P_)
    bk.set_bekker_params(PB_)
    report = {}
    for ctx, tire in ((nn, 0), (bk, 1)):
        worst = 0.0
        for terr in (capi.FLAT, capi.SLOPE, capi.BUMPY):
            ctx.set_terrain(terr)
            X, U, Xd = R[f"ode_{tire}_{terr}_X"], R[f"ode_{tire}_{terr}_U"], R[f"ode_{tire}_{terr}_Xd"]
            worst = max(worst, rel_state_err(ctx.ode(X.T.copy(), U.T.copy()).T, Xd))
        report[f"ode tire {tire}"] = worst
        assert worst < STATE_RTOL
    nn.set_terrain(capi.FLAT)
    report["settle nn"] = rel_state_err(nn.settle()[[0, 1, 2, 3, 6]], R["settle_nn"][[0, 1, 2, 3, 6]])
    bk.set_terrain(capi.FLAT)
    report["settle bekker"] = rel_state_err(bk.settle()[[0, 1, 2, 3, 6]], R["settle_bekker"][[0, 1, 2, 3, 6]])
    for tag, ctx, terr, T in (("roll_nn_flat", nn, capi.FLAT, 120), ("roll_nn_bumpy", nn, capi.BUMPY, 120), ("roll_bekker_flat", bk, capi.FLAT, 100)):
        ctx.set_terrain(terr)
        xl, _ = ctx.rollout(R[tag + "_x0"], R[tag + "_ctrl"], T)
        report[tag] = rel_state_err(xl, R[tag + "_xlist"])
    nn.set_terrain(capi.FLAT)
    _, xf = nn.rollout(R["roll_nn_flat600_x0"], R["roll_nn_flat600_ctrl"], 600, want_list=False)
    report["roll_nn_flat600 final"] = rel_state_err(xf, R["roll_nn_flat600_xfinal"])
    for k, v in report.items():
        assert v < STATE_RTOL, (k, v)
    for tag, terr, T in (("grad_flat20", capi.FLAT, 20), ("grad_slope20", capi.SLOPE, 20), ("grad_bumpy20", capi.BUMPY, 20), ("grad_flat200", capi.FLAT, 200)):
        nn.set_terrain(terr)
        nn.set_nn_params(R[tag + "_params"])
        loss, red, gps = nn.rollout_grad(R[tag + "_x0"], R[tag + "_ctrl"], R[tag + "_gt"], T)
        np.testing.assert_allclose(loss, R[tag + "_loss"], rtol=1e-9)
        g = R[tag + "_grad"]
        e = float(np.max(np.abs(gps - g) / np.abs(g).max(axis=0)))
        report[tag + " grad"] = e
        assert e < GRAD_RTOL, (tag, e)
        assert np.all(gps[104:] == 0.0)
    nn.set_terrain(capi.FLAT)
    nn.set_nn_params(P_)
    nn.reset_optimizer()
    red = np.zeros(418)
    red[:416], red[417] = R["rmsprop_grad"] * 4, 4
    nn.rmsprop_step(red)
    np.testing.assert_allclose(nn.get_nn_params(), R["rmsprop_params"], rtol=1e-14, atol=0)
    nn.set_nn_params(P)
    nn.reset_optimizer()
    print("[reference parity] CUDA path vs the reference's own numbers: " + ", ".join(f"{k} {v:.2e}" for k, v in report.items()))


def test_ode_vs_golden(nn):
    G = golden()
    for terr in (capi.FLAT, capi.SLOPE, capi.BUMPY):
        nn.set_terrain(terr)
        X, U, Xd = G[f"ode_nn_X_{terr}"], G[f"ode_nn_U_{terr}"], G[f"ode_nn_Xd_{terr}"]
        got = nn.ode(X.T.copy(), U.T.copy()).T
        np.testing.assert_allclose(got, Xd, rtol=1e-9, atol=1e-10)
    nn.set_terrain(capi.FLAT)


def test_settle_vs_golden(nn, bk):
    G = golden()
    nn.set_terrain(capi.FLAT)
    nn.set_nn_params(P)
    assert rel_state_err(nn.settle(), G["settle_nn"]) < STATE_RTOL
    bk.set_terrain(capi.FLAT)
    e = rel_state_err(bk.settle(), G["settle_bk"])
    print(f"[bekker parity] settle: rel_state_err {e:.3e}")
    assert e < STATE_RTOL


def test_rollout_vs_golden(nn):
    G = golden()
    nn.set_terrain(capi.FLAT)
    nn.set_nn_params(P)
    xl, xf = nn.rollout(G["roll60_x0"], G["roll60_ctrl"], 60)
    assert rel_state_err(xl, G["roll60_xlist"]) < STATE_RTOL
    assert rel_state_err(xf, G["roll60_xfinal"]) < STATE_RTOL
    # T = 600 (the reference's m_eval_steps, 5990 RK4 steps): final states
    _, xf = nn.rollout(G["roll600_x0"], G["roll600_ctrl"], 600, want_list=False)
    assert rel_state_err(xf, G["roll600_xfinal"]) < STATE_RTOL
    # Bumpy terrain
    nn.set_terrain(capi.BUMPY)
    xl, _ = nn.rollout(G["bumpy_x0"], G["bumpy_ctrl"], 120)
    assert rel_state_err(xl, G["bumpy_xlist"]) < STATE_RTOL
    nn.set_terrain(capi.FLAT)


def test_rollout_vs_oracle_seeded_ragged_batch(nn):
    # N not a multiple of the 8 vehicles a warp holds (tail lanes), T small
    for N, T in ((1, 5), (13, 9), (100, 4)):
        x0, ctrl, gt = synth.make_batch(N, T, seed=100 + N)
        xl, xf = nn.rollout(x0, ctrl, T)
        xl_o, xf_o = po.rollout_batch(po.TIRE_NN, P, po.terrain(po.FLAT), x0, ctrl, T, nthreads=4)
        assert rel_state_err(xl, xl_o) < STATE_RTOL
        assert rel_state_err(xf, xf_o) < STATE_RTOL


def test_substeps_and_minimal_windows(nn):
    """substeps other than the reference's 10 (VehicleSystem.cpp:129), the shortest window (T = 2) and N = 1, forward and adjoint."""
    nn.set_terrain(capi.FLAT)
    nn.set_nn_params(P)
    flat = po.terrain(po.FLAT)
    for N, T, sub in ((1, 2, 10), (3, 4, 3), (5, 3, 1)):
        x0, ctrl, gt = synth.make_batch(N, T, seed=200 + N)
        xl, xf = nn.rollout(x0, ctrl, T, substeps=sub)
        xl_o, xf_o = po.rollout_batch(po.TIRE_NN, P, flat, x0, ctrl, T, substeps=sub)
        assert rel_state_err(xl, xl_o) < STATE_RTOL and rel_state_err(xf, xf_o) < STATE_RTOL
    # the oracle's gradient entry point is fixed at 10 sub-steps, like the reference
    for N, T in ((1, 2), (1, 7), (2, 3)):
        x0, ctrl, gt = synth.make_batch(N, T, seed=300 + T)
        loss, red, gps = nn.rollout_grad(x0, ctrl, gt, T)
        lo, gs, ls, nk, gps_o = po.rollout_grad_batch(P, flat, x0, ctrl, gt, T)
        np.testing.assert_allclose(loss, lo, rtol=1e-9)
        np.testing.assert_allclose(gps, gps_o, rtol=GRAD_RTOL, atol=1e-11 * np.abs(gps_o).max())
        assert red[417] == N


def test_zero_trajectory_length_and_zero_params(nn):
    """traj_len == 0 -> divisor 1 (Trainer.cpp:643-647); all-zero network (the reference's test_wz setup) gives a zero-force vehicle."""
    flat = po.terrain(po.FLAT)
    N, T = 2, 4
    x0, ctrl, gt = synth.make_batch(N, T, seed=5)
    gt[:, 1:, :] = 0.25  # constant ground-truth position -> traj_len = 0
    nn.set_nn_params(P)
    loss, red, gps = nn.rollout_grad(x0, ctrl, gt, T)
    lo, gs, ls, nk, gps_o = po.rollout_grad_batch(P, flat, x0, ctrl, gt, T)
    np.testing.assert_allclose(loss, lo, rtol=1e-9)
    np.testing.assert_allclose(gps, gps_o, rtol=GRAD_RTOL, atol=1e-11 * np.abs(gps_o).max())
    pz = np.zeros(416)
    nn.set_nn_params(pz)
    x = np.zeros((21, 1)); x[3] = 1; x[6] = 100; x[13] = 10
    xl, xf = nn.rollout(x, np.zeros((9, 2, 1)), 10)
    xl_o, xf_o = po.rollout_batch(po.TIRE_NN, pz, flat, x, np.zeros((9, 2, 1)), 10)
    assert rel_state_err(xf, xf_o) < STATE_RTOL
    nn.set_nn_params(P)


def test_rollout_T1_returns_initial_state(nn):
    x0, ctrl, _ = synth.make_batch(4, 2, seed=1)
    xl, xf = nn.rollout(x0, np.zeros((0, 2, 4)), 1)
    np.testing.assert_array_equal(xf, x0)
    np.testing.assert_array_equal(xl[0, :21], x0)


def test_grad_vs_golden(nn):
    G = golden()
    nn.set_terrain(capi.FLAT)
    nn.set_nn_params(P)
    loss, red, gps = nn.rollout_grad(G["grad20_x0"], G["grad20_ctrl"], G["grad20_gt"], 20)
    np.testing.assert_allclose(loss, G["grad20_loss"], rtol=1e-9)
    scale = np.abs(G["grad20_gps"]).max()
    np.testing.assert_allclose(gps, G["grad20_gps"], rtol=GRAD_RTOL, atol=1e-11 * scale)
    np.testing.assert_allclose(red[:416], G["grad20_sum"], rtol=GRAD_RTOL, atol=1e-11 * scale)
    assert red[417] == 4 and red[416] == pytest.approx(G["grad20_loss"].sum(), rel=1e-9)
    assert np.all(gps[104:] == 0.0)
    # the reference's training window (T = 200)
    loss, red, gps = nn.rollout_grad(G["grad200_x0"], G["grad200_ctrl"], G["grad200_gt"], 200)
    np.testing.assert_allclose(loss, G["grad200_loss"], rtol=1e-9)
    scale = np.abs(G["grad200_gps"]).max()
    np.testing.assert_allclose(gps, G["grad200_gps"], rtol=GRAD_RTOL, atol=1e-9 * scale)


def test_grad_vs_oracle_other_terrains_and_params(nn):
    rng = np.random.default_rng(7)
    p2 = P.copy()
    p2[:104] += rng.normal(0, 0.05, 104)
    p2[104:] = rng.normal(0, 1, 312)  # networks 1..3 are dead parameters
    nn.set_nn_params(p2)
    assert np.array_equal(nn.get_nn_params(), p2)
    for terr in (capi.SLOPE, capi.BUMPY):
        nn.set_terrain(terr)
        N, T = 11, 8
        x0, ctrl, gt = synth.make_batch(N, T, seed=50 + terr)
        if terr == capi.BUMPY:
            x0[4] = np.linspace(1.8, 3.7, N)
        loss, red, gps = nn.rollout_grad(x0, ctrl, gt, T)
        lo, gs, ls, nk, gps_o = po.rollout_grad_batch(p2, po.terrain(terr), x0, ctrl, gt, T, nthreads=4)
        np.testing.assert_allclose(loss, lo, rtol=1e-9)
        np.testing.assert_allclose(gps, gps_o, rtol=GRAD_RTOL, atol=1e-11 * np.abs(gps_o).max())
    nn.set_terrain(capi.FLAT)
    nn.set_nn_params(P)


def test_explosion_filter_modes(nn):
    G = golden()
    x0, ctrl, gt, gps = G["grad20_x0"], G["grad20_ctrl"], G["grad20_gt"], G["grad20_gps"]
    thresh = float(np.sort(np.abs(gps).max(axis=0))[2]) * 0.999
    for mode in (0, 1):
        _, red, _ = nn.rollout_grad(x0, ctrl, gt, 20, explode_thresh=thresh, explode_mode=mode)
        _, gs, ls, nk, _ = po.rollout_grad_batch(P, po.terrain(po.FLAT), x0, ctrl, gt, 20, explode_thresh=thresh, explode_mode=mode, nthreads=2)
        assert red[417] == nk
        np.testing.assert_allclose(red[:416], gs, rtol=GRAD_RTOL, atol=1e-11 * np.abs(gps).max())
        assert red[416] == pytest.approx(ls, rel=1e-9)


def test_checkpoint_chunking_is_transparent(nn):
    N, T = 600, 6
    x0, ctrl, gt = synth.make_batch(N, T, seed=77)
    loss_a, red_a, gps_a = nn.rollout_grad(x0, ctrl, gt, T)
    per_vehicle = ((T - 1) * 10 * 4 + 1) * 14 * 8 + (T - 1) * 10 * 4 * 20 * 4 * 8  # state records + activation records
    nn.set_checkpoint_budget(per_vehicle * 256 + 64)  # forces 256-vehicle chunks: 256 + 256 + 88
    loss_b, red_b, gps_b = nn.rollout_grad(x0, ctrl, gt, T)
    nn.set_checkpoint_budget(0)
    np.testing.assert_array_equal(loss_a, loss_b)
    np.testing.assert_array_equal(gps_a, gps_b)
    np.testing.assert_array_equal(red_a, red_b)


def test_rmsprop_matches_reference_update(nn):
    G = golden()
    nn.set_nn_params(P)
    nn.reset_optimizer()
    _, red, _ = nn.rollout_grad(G["grad20_x0"], G["grad20_ctrl"], G["grad20_gt"], 20)
    nn.rmsprop_step(red)
    got = nn.get_nn_params()
    want, _ = po.rmsprop(P, np.ones(416), red[:416] / red[417])
    np.testing.assert_allclose(got, want, rtol=1e-14, atol=0)
    assert np.array_equal(got[104:], P[104:])  # zero gradient -> untouched
    nn.set_nn_params(P)
    nn.reset_optimizer()


def test_reset_optimizer_keeps_stepped_params(nn):
    """avs_reset_optimizer re-initialises the squared-gradient state only: after a device RMSProp step the parameters must
    stay the stepped ones (regression: it used to re-upload the stale host snapshot)."""
    G = golden()
    nn.set_terrain(capi.FLAT)
    nn.set_nn_params(P)
    nn.reset_optimizer()
    _, red, _ = nn.rollout_grad(G["grad20_x0"], G["grad20_ctrl"], G["grad20_gt"], 20)
    nn.rmsprop_step(red)
    nn.reset_optimizer()
    stepped = nn.get_nn_params()
    want, _ = po.rmsprop(P, np.ones(416), red[:416] / red[417])
    np.testing.assert_allclose(stepped, want, rtol=1e-14, atol=0)
    assert not np.array_equal(stepped[:104], P[:104])
    # the rollout that follows must run on the stepped weights, and a second step must start from sq = 1 again
    _, red2, _ = nn.rollout_grad(G["grad20_x0"], G["grad20_ctrl"], G["grad20_gt"], 20)
    _, gs, _, _, _ = po.rollout_grad_batch(stepped, po.terrain(po.FLAT), G["grad20_x0"], G["grad20_ctrl"], G["grad20_gt"], 20, nthreads=4)
    np.testing.assert_allclose(red2[:416], gs, rtol=GRAD_RTOL, atol=1e-11 * np.abs(gs).max())
    nn.rmsprop_step(red2)
    want2, _ = po.rmsprop(stepped, np.ones(416), red2[:416] / red2[417])
    np.testing.assert_allclose(nn.get_nn_params(), want2, rtol=1e-14, atol=0)
    nn.set_nn_params(P)
    nn.reset_optimizer()


def test_bekker_vs_golden(bk):
    G = golden()
    bk.set_terrain(capi.FLAT)
    bk.set_bekker_params(po.bekker_default_params())
    got = bk.ode(G["ode_bk_X"].T.copy(), G["ode_bk_U"].T.copy()).T
    e = float(np.max(np.abs(got - G["ode_bk_Xd"]) / np.maximum(1.0, np.abs(G["ode_bk_Xd"]))))
    print(f"[bekker parity] ODE spot values: worst error relative to max(1, |value|) {e:.3e}")
    np.testing.assert_allclose(got, G["ode_bk_Xd"], rtol=1e-9, atol=1e-9)
    xl, _, mg = bk.rollout_margins(G["bk_x0"], G["bk_ctrl"], 100)
    _, _, mg_o = po.rollout_batch_margins(po.TIRE_BEKKER, po.bekker_default_params(), po.terrain(po.FLAT), G["bk_x0"], G["bk_ctrl"], 100, nthreads=4)
    bekker_parity("golden flat rollout, 100 rows", xl, G["bk_xlist"], mg, mg_o)
    # the margin instantiation is a different kernel: the production kernel must give the same states bit for bit
    xl_p, _ = bk.rollout(G["bk_x0"], G["bk_ctrl"], 100)
    assert np.array_equal(xl_p, xl)


def test_grid_terrain_and_soil(nn, bk):
    g = synth.make_grid_terrain(n=96)
    t = po.terrain(po.GRID, height=g["height"], soil=g["soil"], x0=g["x0"], y0=g["y0"], dx=g["dx"], dy=g["dy"])
    N, T = 9, 12
    x0, ctrl, gt = synth.make_batch(N, T, seed=31)
    x0[4], x0[5] = np.linspace(-1.5, 1.5, N), np.linspace(1.0, -1.0, N)
    nn.set_terrain_grid(g["height"], None, g["x0"], g["y0"], g["dx"], g["dy"])
    xl, _ = nn.rollout(x0, ctrl, T)
    xl_o, _ = po.rollout_batch(po.TIRE_NN, P, t, x0, ctrl, T, nthreads=4)
    assert rel_state_err(xl, xl_o) < STATE_RTOL
    loss, red, gps = nn.rollout_grad(x0, ctrl, gt, T)
    lo, _, _, _, gps_o = po.rollout_grad_batch(P, t, x0, ctrl, gt, T, nthreads=4)
    np.testing.assert_allclose(loss, lo, rtol=1e-9)
    np.testing.assert_allclose(gps, gps_o, rtol=GRAD_RTOL, atol=1e-11 * np.abs(gps_o).max())
    nn.set_terrain(capi.FLAT)
    pb = po.bekker_default_params()
    bk.set_terrain_grid(g["height"], g["soil"], g["x0"], g["y0"], g["dx"], g["dy"])
    x0[6] = 0.06
    xl, _, mg = bk.rollout_margins(x0, np.clip(ctrl, 1, 8), T)
    xl_o, _, mg_o = po.rollout_batch_margins(po.TIRE_BEKKER, pb, t, x0, np.clip(ctrl, 1, 8), T, nthreads=4)
    bekker_parity("96^2 height + soil grid, 12 rows", xl, xl_o, mg, mg_o)
    bk.set_terrain(capi.FLAT)


def test_contact_search_option(nn, bk):
    """the 10-angle contact search of HybridDynamics::get_tire_cpts_sinkages_fancy as a forward option
    (avs_set_contact_search), on the 96^2 height + soil grid for both tire models and on Bumpy terrain, vs the oracle (whose
    function is pinned to the reference's own in tests/test_ref_pinning.py); the adjoint refuses to run with it."""
    g = synth.make_grid_terrain(n=96)
    kw = dict(height=g["height"], soil=g["soil"], x0=g["x0"], y0=g["y0"], dx=g["dx"], dy=g["dy"])
    N, T = 40, 30
    x0, ctrl, gt = synth.make_batch(N, T, seed=41, z_stable=0.06)
    x0[4], x0[5] = np.linspace(-1.5, 1.5, N), np.linspace(1.0, -1.0, N)
    ctrl = np.clip(ctrl, 1, 8)
    t = po.terrain(po.GRID, contact_search=1, **kw)
    try:
        for ctx, tire, params in ((nn, po.TIRE_NN, P), (bk, po.TIRE_BEKKER, po.bekker_default_params())):
            ctx.set_terrain_grid(g["height"], g["soil"], g["x0"], g["y0"], g["dx"], g["dy"])
            xl_plain, _ = ctx.rollout(x0, ctrl, T)
            ctx.set_contact_search(True)
            xl, _ = ctx.rollout(x0, ctrl, T)
            xl_o, _ = po.rollout_batch(tire, params, t, x0, ctrl, T, nthreads=8)
            e = rel_state_err(xl, xl_o)
            print(f"[contact search] tire {tire}: vs oracle {e:.2e}; the option moves the states by up to {np.abs(xl - xl_plain).max():.2e}")
            assert e < STATE_RTOL and np.abs(xl - xl_plain).max() > 1e-6
            if tire == po.TIRE_NN:
                with pytest.raises(capi.AvsError):
                    ctx.rollout_grad(x0, ctrl, gt, T)
                ctx.set_terrain(capi.BUMPY)
                xb = x0.copy()
                xb[4] = np.linspace(1.9, 3.5, N)
                xl, _ = ctx.rollout(xb, ctrl, T)
                xl_o, _ = po.rollout_batch(tire, params, po.terrain(po.BUMPY, contact_search=1), xb, ctrl, T, nthreads=8)
                assert rel_state_err(xl, xl_o) < STATE_RTOL
                ctx.set_terrain(capi.FLAT)  # flat terrain + contact search must leave the lean flat kernel
                xl, _ = ctx.rollout(x0, ctrl, T)
                xl_o, _ = po.rollout_batch(tire, params, po.terrain(po.FLAT, contact_search=1), x0, ctrl, T, nthreads=8)
                assert rel_state_err(xl, xl_o) < STATE_RTOL
            ctx.set_contact_search(False)
    finally:
        nn.set_contact_search(False)
        bk.set_contact_search(False)
        nn.set_terrain(capi.FLAT)
        bk.set_terrain(capi.FLAT)


def test_auxiliary_networks_on_device(nn):
    """BaseNetwork::forward and the data/nn_pretrained.csv network as batched device functions vs their
    restatements (BaseNetwork's is pinned to the reference's own class in tests/test_ref_pinning.py)."""
    rng = np.random.default_rng(8)
    N = 1000
    p = rng.normal(0, 0.7, 131)
    x = rng.normal(0, 2.0, (3, N))
    np.testing.assert_allclose(nn.mlp_forward(capi.MLP_BASE_NETWORK, p, x), po.base_network_forward(p, x), rtol=1e-12, atol=1e-13)
    csv = os.path.join(os.path.dirname(__file__), "golden", "legacy_net_like.csv")
    q = capi.load_legacy_net_csv(csv)
    assert q.size == 435 and np.array_equal(q, np.load(os.path.join(os.path.dirname(__file__), "golden", "legacy_net_like.npy")))
    x5 = q[425:430, None] + q[430:435, None] * rng.normal(0, 1.5, (5, N))
    want = po.legacy_net_forward(q, x5)
    got = nn.mlp_forward(capi.MLP_LEGACY_TIRE, q, x5)
    np.testing.assert_allclose(got, want, rtol=1e-11, atol=1e-11 * np.abs(want).max())


def test_full_size_properties(nn):
    """BASELINE configs[1]/[2] sizes (N = 4096): the oracle cannot finish these, so check size-independent
    properties: permutation equivariance over the batch, determinism, agreement of a strided sample with the
    oracle, unit quaternions, and linearity of the reduction."""
    nn.set_terrain(capi.FLAT)
    nn.set_nn_params(P)
    N, T = 4096, 200
    x0, ctrl, gt = synth.make_batch(N, T)
    loss, red, gps = nn.rollout_grad(x0, ctrl, gt, T)
    assert np.all(np.isfinite(loss)) and np.all(np.isfinite(gps))
    assert red[417] == N
    np.testing.assert_allclose(red[:104], gps[:104].sum(axis=1), rtol=1e-9, atol=1e-9 * np.abs(red[:104]).max())
    # determinism
    loss2, red2, gps2 = nn.rollout_grad(x0, ctrl, gt, T)
    assert np.array_equal(loss, loss2) and np.array_equal(gps, gps2) and np.array_equal(red, red2)
    # permutation equivariance (a vehicle's result does not depend on its lane / CTA)
    perm = np.random.default_rng(0).permutation(N)
    loss_p, _, gps_p = nn.rollout_grad(x0[:, perm].copy(), ctrl[:, :, perm].copy(), gt[:, :, perm].copy(), T)
    assert np.array_equal(loss_p, loss[perm]) and np.array_equal(gps_p, gps[:, perm])
    # strided sample against the oracle
    idx = np.arange(0, N, 64)  # 64 vehicles
    lo, _, _, _, gps_o = po.rollout_grad_batch(P, po.terrain(po.FLAT), x0[:, idx].copy(), ctrl[:, :, idx].copy(), gt[:, :, idx].copy(), T,
                                               nthreads=po.hardware_threads())
    np.testing.assert_allclose(loss[idx], lo, rtol=1e-9)
    np.testing.assert_allclose(gps[:, idx], gps_o, rtol=GRAD_RTOL, atol=1e-9 * np.abs(gps_o).max())
    # forward config: T = 600, final states of a strided sample + unit quaternions everywhere
    x0, ctrl, gt = synth.make_batch(N, 600)
    _, xf = nn.rollout(x0, ctrl, 600, want_list=False)
    np.testing.assert_allclose(np.sqrt((xf[:4] ** 2).sum(axis=0)), 1.0, rtol=0, atol=1e-14)
    _, xf_o = po.rollout_batch(po.TIRE_NN, P, po.terrain(po.FLAT), x0[:, idx].copy(), ctrl[:, :, idx].copy(), 600, want_list=False,
                               nthreads=po.hardware_threads())
    assert rel_state_err(xf[:, idx], xf_o) < STATE_RTOL


def test_external_stream_orders_with_torch(nn):
    """avs_set_stream(0) = the legacy default stream: torch events on it must bracket the context's kernels
    (regression: a NULL handle used to fall back to the context's private stream, so event timing and NCCL
    ordering on torch's stream silently missed the work)."""
    import torch
    N, T = 512, 40
    x0, ctrl, gt = synth.make_batch(N, T, seed=3)
    dev = torch.device("cuda:0")
    d = [torch.from_numpy(a).to(dev) for a in (x0, ctrl, gt)]
    loss = torch.zeros(N, dtype=torch.float64, device=dev)
    red = torch.zeros(418, dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream()
    nn.set_stream(stream.cuda_stream)
    try:
        for _ in range(2):
            nn.rollout_grad_dev(N, T, 10, d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), loss.data_ptr(), red.data_ptr())
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        nn.rollout_grad_dev(N, T, 10, d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), loss.data_ptr(), red.data_ptr())
        e1.record(stream)
        total = red.sum().item()  # implicit sync on torch's stream only: must already see the kernel's result
        f, b = nn.last_kernel_ms()
        assert e0.elapsed_time(e1) >= 0.98 * (f + b) > 0
        assert np.isfinite(total) and red[417].item() == N
    finally:
        nn.use_own_stream()


def test_argument_errors(nn):
    with pytest.raises(capi.AvsError):
        nn.set_terrain(9)
    x0, ctrl, gt = synth.make_batch(4, 5, seed=1)
    with pytest.raises(capi.AvsError):
        nn._ck(capi.lib().avs_rollout(nn._h, 0, 5, 10, None, None, None, None))
    b = capi.Context(0, capi.TIRE_BEKKER)
    with pytest.raises(capi.AvsError):
        b.rollout_grad(x0, ctrl, gt, 5)  # adjoint exists for the NN model only
    b.close()


def test_safety_sweep_matches_oracle(nn):
    """the reference's 10 x 10 (vl, vr) safety sweep over BumpyTerrainMap (test/test_Terrain.cpp:71-122,
    213-346) as one batched rollout; tilt angles and failure rows against the oracle (600 rows instead of 1000: the
    fastest rollouts are on the bump by then)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("safety_sweep", os.path.join(os.path.dirname(__file__), "..", "scripts", "safety_sweep.py"))
    ss = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ss)
    rows = 600
    nn.set_nn_params(P)
    nn.set_terrain(capi.BUMPY)
    x = nn.settle()
    x_o = po.settle(po.TIRE_NN, P, po.terrain(po.BUMPY))
    assert rel_state_err(x, x_o) < STATE_RTOL
    x[4] = x[5] = 0.0
    x[7:] = 0.0
    x0, ctrl = ss.sweep_inputs(x, rows)
    xl, _ = nn.rollout(x0, ctrl, rows + 1, want_final=False)
    xl_o, _ = po.rollout_batch(po.TIRE_NN, P, po.terrain(po.BUMPY), x0, ctrl, rows + 1, nthreads=16)
    nn.set_terrain(capi.FLAT)
    ang, length = ss.safety(xl)
    ang_o, length_o = ss.safety(xl_o)
    np.testing.assert_array_equal(length, length_o)
    np.testing.assert_allclose(ang, ang_o, rtol=0, atol=1e-9)
    ok = length == rows  # compare states only while the rollout counts (a tipped-over vehicle is chaotic afterwards)
    assert ok.sum() > 50
    assert rel_state_err(xl[:, :, ok], xl_o[:, :, ok]) < STATE_RTOL
    assert ang.max() > 0.1  # some rollouts climbed the bump


def test_reference_physical_kats_on_gpu(nn, bk):
    """The assertions the reference's own gtest suites pin, headless and through the C ABI:
    VehicleFixture.no_lateral_drift / straight / circle (test/test_Vehicle.cpp:110-266), BekkerFixture.no_lateral_drift /
    straight / circle (test/test_Bekker.cpp:399-563).  The scenarios of one tire model are ONE batched rollout (N = 3)."""
    def run(ctx, quat, rows):
        x0 = np.zeros((21, 3))
        x0[0:4, :] = np.asarray(quat)[:, None]
        x0[6, :] = .0605                      # the fixtures' "stable" z (test_Vehicle.cpp:64, test_Bekker.cpp:81)
        x0[15, 0] = .1                        # scenario 0: lateral velocity 0.1, no control
        ctrl = np.zeros((rows, 2, 3))
        ctrl[:, :, 1] = 4.0                   # scenario 1: straight, vl = vr = 4
        return x0, ctrl

    nn.set_terrain(capi.FLAT)
    nn.set_nn_params(P)
    x0, ctrl = run(nn, nn.settle()[:4], 1000)
    ctrl[:, 0, 2], ctrl[:, 1, 2] = 2.0, 1.0  # scenario 2: circle, vl = 2, vr = 1 (first 1000 of the reference's 10000 rows)
    xl, _ = nn.rollout(x0, ctrl, 1001, want_final=False)
    assert 0.0 < xl[-1, 5, 0] < 0.1
    assert xl[-1, 4, 1] == pytest.approx(1.76, abs=0.02)  # 0.98 needs the author's private tire.net
    xl_o, _ = po.rollout_batch(po.TIRE_NN, P, po.terrain(po.FLAT), x0, ctrl, 1001, nthreads=3)
    assert rel_state_err(xl, xl_o) < STATE_RTOL
    # the circle proper: 10000 rows, x extent == y extent within 1e-2
    x0c, ctrlc = x0[:, 2:3].copy(), np.zeros((10000, 2, 1))
    ctrlc[:, 0, 0], ctrlc[:, 1, 0] = 2.0, 1.0
    xc, _ = nn.rollout(x0c, ctrlc, 10001, want_final=False)
    ext_x, ext_y = np.ptp(xc[:, 4, 0]), np.ptp(xc[:, 5, 0])
    assert ext_x == pytest.approx(ext_y, abs=1e-2)

    bk.set_terrain(capi.FLAT)
    bk.set_bekker_params(po.bekker_default_params())
    x0, ctrl = run(bk, [0, 0, 0, 1], 1000)
    ctrl[:, 0, 2], ctrl[:, 1, 2] = 2.0, -2.0  # test_Bekker.cpp:497-563: vl = 2, vr = -2 (turn in place)
    xl, _ = bk.rollout(x0, ctrl, 1001, want_final=False)
    assert 0.0 < xl[300, 5, 0] < 0.1
    assert xl[-1, 4, 1] == pytest.approx(3.92, abs=0.1)   # test_Bekker.cpp:492
    assert np.ptp(xl[:, 4, 2]) == pytest.approx(np.ptp(xl[:, 5, 2]), abs=1e-2)
    # scenario 0 (no control: wR = 0 sits exactly on the slip clamp and the Fx sign switch) is reported, scenario 1 asserted
    xl_m, _, mg = bk.rollout_margins(x0[:, :2], ctrl[:300, :, :2], 301)
    assert np.array_equal(xl_m, xl[:301, :, :2])
    xl_o, _, mg_o = po.rollout_batch_margins(po.TIRE_BEKKER, po.bekker_default_params(), po.terrain(po.FLAT), x0[:, :2], ctrl[:300, :, :2], 301, nthreads=2)
    bekker_parity("reference KAT drives, 300 rows", xl_m, xl_o, mg, mg_o)


def test_config3_bekker_grid_full_size_properties(bk):
    """BASELINE configs[3]: 65 536 Bekker rollouts on randomised height / soil grids over 8 GPUs = 8192 per GPU, T = 600 rows
    (5990 RK4 steps).  At that size: determinism, permutation equivariance, finite states, unit quaternions, and a strided
    sample of 64 vehicles against the oracle at 1e-9 on every margin-safe trajectory."""
    g = synth.make_grid_terrain()
    bk.set_terrain_grid(g["height"], g["soil"], g["x0"], g["y0"], g["dx"], g["dy"])
    bk.set_bekker_params(po.bekker_default_params())
    N, T = 8192, 600
    x0, ctrl, _ = synth.make_batch(N, T, z_stable=0.065)
    ctrl = np.clip(ctrl, 0.5, 8.0)
    _, xf = bk.rollout(x0, ctrl, T, want_list=False)
    assert np.all(np.isfinite(xf))
    np.testing.assert_allclose(np.sqrt((xf[:4] ** 2).sum(axis=0)), 1.0, rtol=0, atol=1e-14)
    _, xf2 = bk.rollout(x0, ctrl, T, want_list=False)
    assert np.array_equal(xf, xf2)
    perm = np.random.default_rng(1).permutation(N)
    _, xf_p = bk.rollout(x0[:, perm].copy(), ctrl[:, :, perm].copy(), T, want_list=False)
    assert np.array_equal(xf_p, xf[:, perm])
    idx = np.arange(0, N, 128)  # 64 vehicles
    xs, cs = x0[:, idx].copy(), ctrl[:, :, idx].copy()
    xl_s, xf_s, mg = bk.rollout_margins(xs, cs, T)
    assert np.array_equal(xf_s, xf[:, idx])  # margin kernel == production kernel, and a vehicle's result does not depend on the batch
    t = po.terrain(po.GRID, height=g["height"], soil=g["soil"], x0=g["x0"], y0=g["y0"], dx=g["dx"], dy=g["dy"])
    xl_o, xf_o, mg_o = po.rollout_batch_margins(po.TIRE_BEKKER, po.bekker_default_params(), t, xs, cs, T, nthreads=po.hardware_threads())
    bekker_parity("configs[3] 8192 x 600 rows, 64-vehicle strided sample, all rows", xl_s, xl_o, mg, mg_o)
    bk.set_terrain(capi.FLAT)


def test_config4_65536_windows_per_step(nn):
    """BASELINE configs[4] on one GPU: 65 536 windows x 200 rows per training step do not fit the record budget at once
    (392 GB of records) and run as chunks; the first 4096 windows must come out bit-identical to a 4096-window call, the
    reduction must be the sum of the per-sample gradients, and nothing may be dropped."""
    nn.set_terrain(capi.FLAT)
    nn.set_nn_params(P)
    N, T = 65536, 200
    x0s, ctrls, gts = synth.make_batch(4096, T)
    reps = N // 4096
    x0, ctrl, gt = np.tile(x0s, (1, reps)), np.tile(ctrls, (1, 1, reps)), np.tile(gts, (1, 1, reps))
    loss, red, _ = nn.rollout_grad(x0, ctrl, gt, T, per_sample=False)
    loss_s, red_s, _ = nn.rollout_grad(x0s, ctrls, gts, T, per_sample=False)
    assert np.all(np.isfinite(loss)) and red[417] == N
    for r in range(reps):
        assert np.array_equal(loss[r * 4096:(r + 1) * 4096], loss_s)
    np.testing.assert_allclose(red[:416], reps * red_s[:416], rtol=1e-10, atol=1e-12 * np.abs(red_s[:416]).max() * reps)
    np.testing.assert_allclose(red[416], reps * red_s[416], rtol=1e-12)
