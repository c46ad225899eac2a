"""Generates tests/golden/golden_torch_v1.npz with the independent torch-fp64 autograd restatement (tests/torch_restatement.py):
ODE values on flat and sloped ground, every row of a 10-macro-step rollout, and per-sample losses + gradients of the training
objective at T = 20 with perturbed weights.  Needs torch and numpy only (no reference, no oracle).  usage: python tests/golden/make_golden_torch.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import torch_restatement as tr  # noqa: E402
from auvsl_neural_ode_b200 import synth  # noqa: E402

out = {}
rng = np.random.default_rng(20231115)
p0 = tr.default_live_params()
out["live_params_default"] = p0

# ODE values: states around the operating point (unit quaternions with roll / pitch / yaw, wheels in and out of contact)
M = 12
ang = rng.uniform(-0.15, 0.15, (M, 2))
yaw = rng.uniform(-np.pi, np.pi, M)
X = np.zeros((M, 21))
for i in range(M):
    cr, sr, cp, sp, cy, sy = np.cos(ang[i, 0] / 2), np.sin(ang[i, 0] / 2), np.cos(ang[i, 1] / 2), np.sin(ang[i, 1] / 2), np.cos(yaw[i] / 2), np.sin(yaw[i] / 2)
    X[i, 0:4] = [sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy, cr * cp * cy + sr * sp * sy]
X[:, 4:6] = rng.uniform(-2, 2, (M, 2))
X[:, 6] = rng.uniform(0.055, 0.068, M)
X[:, 7:11] = rng.uniform(-3, 3, (M, 4))
X[:, 11:14] = rng.uniform(-0.5, 0.5, (M, 3))
X[:, 14:17] = rng.uniform(-0.6, 0.6, (M, 3))
U = rng.uniform(-2, 8, (M, 2))
X[:, 17], X[:, 19], X[:, 18], X[:, 20] = U[:, 0], U[:, 0], U[:, 1], U[:, 1]
out["ode_X"], out["ode_U"] = X, U
with torch.no_grad():
    for terr in ("flat", "slope"):
        Xs = X.copy()
        if terr == "slope":
            Xs[:, 6] += 0.1 * Xs[:, 4]  # keep the wheels near the surface z = 0.1 x
            out["ode_X_slope"] = Xs
        out[f"ode_Xd_{terr}"] = tr.ode(torch.tensor(p0), torch.tensor(Xs), torch.tensor(U), terr).numpy()

    # 10 macro steps (100 RK4 steps), every row
    x0, ctrl, gt = synth.make_batch(4, 11, seed=11)
    _, states = tr.rollout_loss(torch.tensor(p0), x0, ctrl, gt, "flat")
    out["roll_x0"], out["roll_ctrl"], out["roll_states"] = x0, ctrl, states.numpy()

# training objective at T = 20, perturbed weights
p = p0 + 0.05 * rng.standard_normal(104)
x0, ctrl, gt = synth.make_batch(4, 20, seed=77)
out["grad_params"], out["grad_x0"], out["grad_ctrl"], out["grad_gt"] = p, x0, ctrl, gt
for terr in ("flat", "slope"):
    xs = x0.copy()
    L, G, S = tr.loss_and_grad(p, xs, ctrl, gt, terr)
    out[f"grad_loss_{terr}"], out[f"grad_grad_{terr}"], out[f"grad_xfinal_{terr}"] = L, G, S[-1]
    print(terr, "loss", L, "|grad|max", np.abs(G).max())
np.savez_compressed(os.path.join(HERE, "golden_torch_v1.npz"), **out)
print("wrote golden_torch_v1.npz", {k: v.shape for k, v in out.items()})
