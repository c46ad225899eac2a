"""Regenerates tests/golden/*.npz from the CPU oracle (oracle/).  The reference's own tests pin no numeric
vectors for this path and the reference cannot be built in this image, so these fixtures
freeze the ORACLE's outputs ("parity unpinned"); they guard the oracle against regressions and give the GPU
tests fixed inputs/outputs that travel to the GPU box.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import pyoracle as po  # noqa: E402
from auvsl_neural_ode_b200 import synth  # noqa: E402


def rand_states(rng, n, bumpy=False):
    X = np.zeros((n, 21))
    U = rng.uniform(-2, 12, size=(n, 2))
    for i in range(n):
        q = rng.normal(size=4)
        q[:2] *= 0.05
        X[i, :4] = q / np.linalg.norm(q)
        X[i, 4:7] = [rng.uniform(1.5, 4.0) if bumpy else rng.normal(), rng.normal(), 0.058 + 0.004 * rng.normal()]
        X[i, 7:11] = rng.normal(size=4)
        X[i, 11:14] = rng.normal(size=3) * [0.2, 0.2, 1.0]
        X[i, 14:17] = rng.normal(size=3) * [1, 0.3, 0.1]
        X[i, 17:21] = [U[i, 0], U[i, 1], U[i, 0], U[i, 1]]
    return X, U


def main():
    rng = np.random.default_rng(20231115)
    p = po.nn_default_params()
    pb = po.bekker_default_params()
    out = {}
    # single ODE evaluations and RK4 steps, NN model, three analytic terrains
    for terr in (po.FLAT, po.SLOPE, po.BUMPY):
        t = po.terrain(terr)
        X, U = rand_states(rng, 32, bumpy=(terr == po.BUMPY))
        out[f"ode_nn_X_{terr}"], out[f"ode_nn_U_{terr}"] = X, U
        out[f"ode_nn_Xd_{terr}"] = np.stack([po.ode(po.TIRE_NN, p, t, x, u) for x, u in zip(X, U)])
        out[f"rk4_nn_X1_{terr}"] = np.stack([po.rk4(po.TIRE_NN, p, t, x, u) for x, u in zip(X, U)])
    # Bekker ODE, flat
    X, U = rand_states(rng, 32)
    X[:, 6] = 0.09 + 0.004 * rng.normal(size=32)
    U = np.abs(U) + 0.5
    X[:, 17:21] = np.stack([U[:, 0], U[:, 1], U[:, 0], U[:, 1]], axis=1)
    X[16:, 17:21] += rng.normal(0, 0.3, (16, 4))  # System::forward may be called with joint velocities != controls
    out["ode_bk_X"], out["ode_bk_U"] = X, U
    out["ode_bk_Xd"] = np.stack([po.ode(po.TIRE_BEKKER, pb, po.terrain(po.FLAT), x, u) for x, u in zip(X, U)])
    # settle()
    out["settle_nn"] = po.settle(po.TIRE_NN, p, po.terrain(po.FLAT))
    out["settle_bk"] = po.settle(po.TIRE_BEKKER, pb, po.terrain(po.FLAT))
    # forward rollouts, NN: x_list for T = 60, final states for T = 600
    x0, ctrl, gt = synth.make_batch(8, 60, seed=1)
    xl, xf = po.rollout_batch(po.TIRE_NN, p, po.terrain(po.FLAT), x0, ctrl, 60)
    out.update(roll60_x0=x0, roll60_ctrl=ctrl, roll60_xlist=xl, roll60_xfinal=xf)
    x0, ctrl, gt = synth.make_batch(4, 600, seed=2)
    _, xf = po.rollout_batch(po.TIRE_NN, p, po.terrain(po.FLAT), x0, ctrl, 600, want_list=False, nthreads=4)
    out.update(roll600_x0=x0, roll600_ctrl=ctrl, roll600_xfinal=xf, roll600_gt=gt)
    # Bumpy-terrain rollout (drives over the bump of TestTerrainMaps.cpp:18-28)
    x0, ctrl, gt = synth.make_batch(4, 120, seed=3)
    x0[0:4] = [[0.0], [0.0], [0.0], [1.0]]
    x0[4] = 1.7
    ctrl[:] = np.clip(ctrl, 3.0, 6.0)
    xl, xf = po.rollout_batch(po.TIRE_NN, p, po.terrain(po.BUMPY), x0, ctrl, 120)
    out.update(bumpy_x0=x0, bumpy_ctrl=ctrl, bumpy_xlist=xl)
    # loss + gradient (reverse tape), T = 20 (N = 4) and the reference's T = 200 (N = 2)
    x0, ctrl, gt = synth.make_batch(4, 20, seed=4)
    loss, gs, ls, nk, gps = po.rollout_grad_batch(p, po.terrain(po.FLAT), x0, ctrl, gt, 20, nthreads=4)
    out.update(grad20_x0=x0, grad20_ctrl=ctrl, grad20_gt=gt, grad20_loss=loss, grad20_gps=gps, grad20_sum=gs)
    x0, ctrl, gt = synth.make_batch(2, 200, seed=5)
    loss, gs, ls, nk, gps = po.rollout_grad_batch(p, po.terrain(po.FLAT), x0, ctrl, gt, 200, nthreads=2)
    out.update(grad200_x0=x0, grad200_ctrl=ctrl, grad200_gt=gt, grad200_loss=loss, grad200_gps=gps)
    # Bekker forward rollout, flat, T = 100
    x0, ctrl, gt = synth.make_batch(2, 100, seed=6, z_stable=float(out["settle_bk"][6]))
    ctrl[:] = np.clip(ctrl, 1.0, 8.0)
    xl, xf = po.rollout_batch(po.TIRE_BEKKER, pb, po.terrain(po.FLAT), x0, ctrl, 100, nthreads=2)
    out.update(bk_x0=x0, bk_ctrl=ctrl, bk_xlist=xl)
    np.savez_compressed(os.path.join(HERE, "golden_v1.npz"), **out)
    print("wrote", os.path.join(HERE, "golden_v1.npz"), {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
