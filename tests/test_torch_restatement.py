"""A second, independent engine behind the oracle: tests/golden/golden_torch_v1.npz was
produced by tests/torch_restatement.py — dense 6x6 spatial algebra on batched torch fp64 tensors, differentiated by
torch.autograd — which shares neither code nor structure with the oracle, the compiled reference sources (both use this
repo's reverse tape where the reference uses CppAD) or the kernels.  CPU part: the oracle, the reference library (where it was
built) and the CPU build of the device math against those numbers (ODE 1e-12, rollout states 1e-12, gradients 1e-9: well
inside the north star's 1e-9 / 1e-7), plus a live run of the restatement so the fixture cannot go stale.  GPU part: the CUDA
path through the C ABI against the same numbers at the north star's bars."""
import os

import numpy as np
import pytest

import hostcheck_lib as hc
from oracle import pyoracle as po

T_ = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_torch_v1.npz"))
TERR = {"flat": 0, "slope": 1}


def rel(a, b):
    return float(np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b))))


def p416(live):
    p = po.nn_default_params().copy()
    p[:104] = live
    return p


def test_fixture_weights_are_the_reference_defaults():
    assert np.array_equal(T_["live_params_default"], po.nn_default_params()[:104])


@pytest.mark.parametrize("terr", ["flat", "slope"])
def test_ode_values(terr):
    X, U, Xd = (T_["ode_X"] if terr == "flat" else T_["ode_X_slope"]), T_["ode_U"], T_[f"ode_Xd_{terr}"]
    P = po.nn_default_params()
    m = hc.model(0, TERR[terr], P)
    worst_o = worst_h = 0.0
    for i in range(X.shape[0]):
        worst_o = max(worst_o, rel(po.ode(0, P, po.terrain(TERR[terr]), X[i], U[i]), Xd[i]))
        worst_h = max(worst_h, rel(hc.ode(m, X[i][hc.LIVE], U[i], X[i][17:21]), Xd[i][hc.LIVE]))
    print(f"ODE {terr}: oracle vs torch {worst_o:.2e}, device math (CPU build) vs torch {worst_h:.2e}")
    assert worst_o < 1e-12 and worst_h < 1e-12
    assert np.abs(Xd[:, 11:17]).max() > 1.0  # the wheels are in contact in the fixture: not a free-fall comparison


def test_ten_macro_steps_every_row():
    x0, ctrl, want = T_["roll_x0"], T_["roll_ctrl"], T_["roll_states"]
    P = po.nn_default_params()
    xl_o, _ = po.rollout_batch(0, P, po.terrain(0), x0, ctrl, 11)
    xl_h, _ = hc.rollout(hc.model(0, 0, P), x0, ctrl, 11)
    eo, eh = rel(xl_o[:, :21], want), rel(xl_h[:, :21], want)
    print(f"10 macro steps: oracle vs torch {eo:.2e}, device math (CPU build) vs torch {eh:.2e}")
    assert eo < 1e-12 and eh < 1e-12


@pytest.mark.parametrize("terr", ["flat", "slope"])
def test_training_objective_loss_and_gradient(terr):
    """torch.autograd against the oracle's reverse tape and the hand-written adjoint"""
    P = p416(T_["grad_params"])
    x0, ctrl, gt = T_["grad_x0"], T_["grad_ctrl"], T_["grad_gt"]
    L, G = T_[f"grad_loss_{terr}"], T_[f"grad_grad_{terr}"]
    lo, _, _, _, gps = po.rollout_grad_batch(P, po.terrain(TERR[terr]), x0, ctrl, gt, 20, explode_thresh=1e30, nthreads=4)
    lh, gh = hc.rollout_grad(hc.model(0, TERR[terr], P), x0, ctrl, gt, 20)
    scale = np.abs(G).max(axis=0)
    eo, eh = float(np.max(np.abs(gps[:104] - G) / scale)), float(np.max(np.abs(gh - G) / scale))
    print(f"T = 20 {terr}: gradient error / max|g| per sample: oracle tape vs autograd {eo:.2e}, hand adjoint (CPU build) vs autograd {eh:.2e}")
    np.testing.assert_allclose(lo, L, rtol=1e-12)
    np.testing.assert_allclose(lh, L, rtol=1e-12)
    assert eo < 1e-9 and eh < 1e-9
    assert np.all(gps[104:] == 0.0)


def test_reference_library_agrees_with_torch():
    """the reference's own sources (oracle/_ref, built only where /root/reference exists) against the torch numbers"""
    from oracle import pyref
    if not pyref.available():
        pytest.skip("oracle/_ref not built here")
    P = po.nn_default_params()
    worst = 0.0
    for terr in ("flat", "slope"):
        X, U, Xd = (T_["ode_X"] if terr == "flat" else T_["ode_X_slope"]), T_["ode_U"], T_[f"ode_Xd_{terr}"]
        for i in range(X.shape[0]):
            worst = max(worst, rel(pyref.ode(0, P, TERR[terr], X[i], U[i]), Xd[i]))
    print(f"reference library ODE vs torch: {worst:.2e}")
    assert worst < 1e-12


def test_live_restatement_matches_fixture_and_oracle():
    """runs the torch restatement here (a few ODE evaluations and one RK4 step with autograd) so that the committed fixture
    and the module that made it cannot drift apart"""
    import torch

    import torch_restatement as tr
    p = torch.tensor(T_["live_params_default"], requires_grad=True)
    X, U = torch.tensor(T_["ode_X"]), torch.tensor(T_["ode_U"])
    Xd = tr.ode(p, X, U, "flat")
    assert rel(Xd.detach().numpy(), T_["ode_Xd_flat"]) < 1e-14
    # d(sum of one RK4 step's base twist)/d(weights) by autograd vs central differences of the ORACLE's rk4
    Y = tr.rk4(p, X[:2], U[:2], "flat")
    g, = torch.autograd.grad(Y[:, 11:17].sum(), p)
    g = g.numpy()
    P, t = po.nn_default_params(), po.terrain(0)
    for k in (0, 5, 30, 60, 90, 103):
        h = 1e-6 * max(1.0, abs(P[k]))
        Pp, Pm = P.copy(), P.copy()
        Pp[k] += h
        Pm[k] -= h
        fd = sum((po.rk4(0, Pp, t, T_["ode_X"][i], T_["ode_U"][i])[11:17] - po.rk4(0, Pm, t, T_["ode_X"][i], T_["ode_U"][i])[11:17]).sum() for i in range(2)) / (2 * h)
        assert abs(fd - g[k]) <= 1e-6 * max(1e-3, abs(g).max()), (k, fd, g[k])


# ---- Bekker tire model: independent numpy restatement (tests/bekker_restatement.py: whole arcs as numpy expressions) ----------------
def _bekker_cases(n=300, seed=5):
    rng = np.random.default_rng(seed)
    for i in range(n):
        zr = rng.uniform(1e-4, 0.03) if i % 10 else rng.uniform(0.03, 0.12)  # every tenth: saturated ground pressure / zr > R
        soil = np.array([29.76 * rng.uniform(.8, 1.2), 2083 * rng.uniform(.8, 1.2), 0.8 * rng.uniform(.8, 1.2), rng.uniform(0, .1), .3927 * rng.uniform(.8, 1.2)])
        yield [zr, rng.uniform(-1, 1.5), rng.uniform(-0.5, 0.5), 0.098 * rng.uniform(-4, 12)], soil


def test_bekker_tire_forces_vs_numpy_restatement():
    import bekker_restatement as br
    from oracle import pyref
    have_ref = pyref.available()
    worst_o = worst_r = 0.0
    for inp, soil in _bekker_cases():
        want = np.array(br.get_forces(*br.get_features(*inp), *soil))
        feat, f = po.bekker_forces(inp, soil)
        assert np.allclose(feat[1:], br.get_features(*inp)[1:], rtol=0, atol=1e-15)
        worst_o = max(worst_o, rel(f[:3], want))
        if have_ref:
            _, fr = pyref.bekker_forces(inp, soil)
            worst_r = max(worst_r, rel(np.asarray(fr)[:3], want))
    print(f"Bekker get_forces vs numpy restatement: oracle {worst_o:.2e}, reference library {worst_r:.2e}" + ("" if have_ref else " (not built here)"))
    assert worst_o < 1e-11 and worst_r < 1e-11


def _bekker_short_rollout():
    import torch

    import torch_restatement as tr
    from auvsl_neural_ode_b200 import synth
    PB = po.bekker_default_params()
    fn = tr.bekker_wheel_forces(PB)
    x0, ctrl, _ = synth.make_batch(3, 4, seed=5, z_stable=0.065)
    ctrl = np.clip(ctrl, 0.5, 8.0)
    with torch.no_grad():
        Xt = torch.tensor(x0).T.contiguous()
        for i in range(3):
            Xt = tr.macro_step(None, Xt, torch.tensor(ctrl[i]).T, "flat", 10, fn)
    return PB, fn, x0, ctrl, Xt.numpy().T


def test_bekker_ode_and_rollout_vs_independent_restatements():
    """dense-algebra ODE (torch) around the numpy Bekker model against the oracle and the CPU build of the device math"""
    import torch

    import torch_restatement as tr
    PB, fn, x0, ctrl, xf_want = _bekker_short_rollout()
    worst_o = worst_h = 0.0
    for terr in ("flat", "slope"):
        X, U = (T_["ode_X"] if terr == "flat" else T_["ode_X_slope"]), T_["ode_U"]
        with torch.no_grad():
            Xd = tr.ode(None, torch.tensor(X), torch.tensor(U), terr, fn).numpy()
        assert np.abs(Xd[:, 11:17]).max() > 100.0  # wheels in the soil: the Bekker forces dominate the accelerations
        m = hc.model(1, TERR[terr], PB)
        for i in range(X.shape[0]):
            worst_o = max(worst_o, rel(po.ode(1, PB, po.terrain(TERR[terr]), X[i], U[i]), Xd[i]))
            worst_h = max(worst_h, rel(hc.ode(m, X[i][hc.LIVE], U[i], X[i][17:21]), Xd[i][hc.LIVE]))
    _, xf_o = po.rollout_batch(1, PB, po.terrain(0), x0, ctrl, 4)
    _, xf_h = hc.rollout(hc.model(1, 0, PB), x0, ctrl, 4, want_list=False)
    eo, eh = rel(xf_o, xf_want), rel(xf_h[hc.LIVE] if xf_h.shape[0] == 21 else xf_h, xf_want[hc.LIVE] if xf_h.shape[0] == 21 else xf_want)
    print(f"Bekker ODE: oracle {worst_o:.2e}, device math (CPU build) {worst_h:.2e}; 30 RK4 steps: oracle {eo:.2e}, device math {eh:.2e}")
    assert worst_o < 1e-11 and worst_h < 1e-9 and eo < 1e-11 and eh < 1e-9


@pytest.mark.gpu
def test_cuda_bekker_path_vs_independent_restatements():
    import torch

    import torch_restatement as tr
    from auvsl_neural_ode_b200 import capi
    PB, fn, x0, ctrl, xf_want = _bekker_short_rollout()
    ctx = capi.Context(0, capi.TIRE_BEKKER)
    ctx.set_bekker_params(PB)
    report = {}
    for terr in ("flat", "slope"):
        ctx.set_terrain(capi.FLAT if terr == "flat" else capi.SLOPE)
        X, U = (T_["ode_X"] if terr == "flat" else T_["ode_X_slope"]), T_["ode_U"]
        with torch.no_grad():
            Xd = tr.ode(None, torch.tensor(X), torch.tensor(U), terr, fn).numpy()
        report[f"ode {terr}"] = rel(ctx.ode(X.T.copy(), U.T.copy()).T, Xd)
    ctx.set_terrain(capi.FLAT)
    _, xf = ctx.rollout(x0, ctrl, 4, want_list=False)
    report["30 RK4 steps"] = rel(xf, xf_want)
    print("CUDA Bekker path vs numpy / torch restatements:", {k: f"{v:.2e}" for k, v in report.items()})
    for k, v in report.items():
        assert v < 1e-9, (k, v)


@pytest.mark.gpu
def test_cuda_path_vs_torch_autograd():
    """the CUDA path through the C ABI against the independent torch numbers: ODE values, every row of a 10-macro-step
    rollout (1e-9), per-sample losses and gradients at T = 20 on flat and sloped ground (1e-7 of max|g|)"""
    from auvsl_neural_ode_b200 import capi
    ctx = capi.Context(0, capi.TIRE_NN)
    P = capi.nn_default_params()
    ctx.set_nn_params(P)
    report = {}
    for terr in ("flat", "slope"):
        ctx.set_terrain(capi.FLAT if terr == "flat" else capi.SLOPE)
        X, U, Xd = (T_["ode_X"] if terr == "flat" else T_["ode_X_slope"]), T_["ode_U"], T_[f"ode_Xd_{terr}"]
        report[f"ode {terr}"] = rel(ctx.ode(X.T.copy(), U.T.copy()).T, Xd)
    ctx.set_terrain(capi.FLAT)
    xl, _ = ctx.rollout(T_["roll_x0"], T_["roll_ctrl"], 11)
    report["10 macro steps"] = rel(xl[:, :21], T_["roll_states"])
    for k, v in report.items():
        assert v < 1e-9, (k, v)
    ctx.set_nn_params(p416(T_["grad_params"]))
    for terr in ("flat", "slope"):
        ctx.set_terrain(capi.FLAT if terr == "flat" else capi.SLOPE)
        loss, _, gps = ctx.rollout_grad(T_["grad_x0"], T_["grad_ctrl"], T_["grad_gt"], 20, explode_thresh=1e30)
        G = T_[f"grad_grad_{terr}"]
        e = float(np.max(np.abs(gps[:104] - G) / np.abs(G).max(axis=0)))
        report[f"grad {terr}"] = e
        np.testing.assert_allclose(loss, T_[f"grad_loss_{terr}"], rtol=1e-9)
        assert e < 1e-7, (terr, e)
        assert np.all(gps[104:] == 0.0)
    print("CUDA path vs torch autograd:", {k: f"{v:.2e}" for k, v in report.items()})
