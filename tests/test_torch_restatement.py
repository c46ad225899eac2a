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
