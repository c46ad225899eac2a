"""Generates tests/golden/golden_ref_v1.npz: inputs and outputs of THE REFERENCE'S OWN CODE (oracle/_ref/libjackal_ref.so =
the reference's first-party sources compiled in place from /root/reference behind the stand-in headers of oracle/refshim/;
see oracle/ref_capi.cpp), in the SoA layout of include/avsim.h so that the CUDA path can be checked against them directly on
the GPU box, where /root/reference does not exist.  Run in the build container:  python tests/golden/make_golden_ref.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyref as pr  # noqa: E402

B = np.array([[0.0472, 0.0438], [0.0036, -0.0035], [-0.1744, 0.1748]])  # scripts/gen_fake_data.py:8-10
rng = np.random.default_rng(20231115)
P = pr.default_params(0)
PB = pr.default_params(1)
Z_NN = pr.default_initial_state(0, P, 0)[6]
Z_BK = pr.default_initial_state(1, PB, 0)[6]


def window(T, z, x_shift=0.0):
    """rows [T][10] = (time, vl, vr, x, y, z, yaw, wz, vx, vy) as Trainer::loadDataFile yields them"""
    t = np.arange(T) * 0.01
    vbar, delta, amp, f = rng.uniform(1, 7), rng.uniform(-2, 2), rng.uniform(0, 1), rng.uniform(0.05, 0.5)
    ph = rng.uniform(0, 2 * np.pi, 2)
    vl = np.clip(vbar + delta + amp * np.sin(2 * np.pi * f * t + ph[0]), -2, 12)
    vr = np.clip(vbar - delta + amp * np.sin(2 * np.pi * f * t + ph[1]), -2, 12)
    x, y, yaw = x_shift, 0.0, rng.uniform(-np.pi, np.pi) if x_shift == 0.0 else rng.uniform(-0.4, 0.4)
    rows = np.zeros((T, 10))
    for i in range(T):
        v = B @ np.array([vl[i], vr[i]])
        rows[i] = (t[i], vl[i], vr[i], x, y, z, yaw, v[2], v[0], v[1])
        for _ in range(10):
            x += 1e-3 * (np.cos(yaw) * v[0] - np.sin(yaw) * v[1])
            y += 1e-3 * (np.sin(yaw) * v[0] + np.cos(yaw) * v[1])
            yaw += 1e-3 * v[2]
    return rows


def soa(windows):
    """windows -> x0 [21][N], ctrl [T-1][2][N], gt [T][3][N] exactly as the host layer packs them (Trainer::pack)"""
    N, T = len(windows), windows[0].shape[0]
    x0, ctrl, gt = np.zeros((21, N)), np.zeros((T - 1, 2, N)), np.zeros((T, 3, N))
    for n, w in enumerate(windows):
        x0[:, n] = pr.initialize_state(w[0])[:21]
        ctrl[:, 0, n], ctrl[:, 1, n] = w[:-1, 1], w[:-1, 2]
        gt[:, 0, n], gt[:, 1, n], gt[:, 2, n] = w[:, 6], w[:, 3], w[:, 4]
    return x0, ctrl, gt


out = {"nn_params": P, "bekker_params": PB, "z_stable_nn": Z_NN, "z_stable_bekker": Z_BK,
       "settle_nn": pr.default_initial_state(0, P, 0), "settle_bekker": pr.default_initial_state(1, PB, 0)}

# ---- ODE spot values (System::forward)
for tire, params, z in ((0, P, 0.0590), (1, PB, 0.0620)):
    for terr in (0, 1, 2):
        Xs, Us, Xds = [], [], []
        for _ in range(12):
            q = np.array([rng.normal(0, 0.02), rng.normal(0, 0.02), rng.normal(0, 0.5), 1.0])
            q /= np.linalg.norm(q)
            X = np.zeros(21)
            X[0:4] = q
            X[4:7] = (rng.uniform(1.5, 4.0), rng.uniform(-1, 1), z + rng.normal(0, 0.002))
            if terr == 1:
                X[6] += 0.1 * X[4]
            X[11:14] = rng.normal(0, 0.3, 3)
            X[14:17] = (rng.uniform(-0.2, 0.8), rng.normal(0, 0.1), rng.normal(0, 0.05))
            u = rng.uniform(0.5, 8, 2)
            X[17:21] = (u[0], u[1], u[0], u[1])
            Xs.append(X); Us.append(u); Xds.append(pr.ode(tire, params, terr, X, u))
        out[f"ode_{tire}_{terr}_X"], out[f"ode_{tire}_{terr}_U"], out[f"ode_{tire}_{terr}_Xd"] = np.array(Xs), np.array(Us), np.array(Xds)

# ---- forward rollouts through Trainer::evaluateTrajectory (x_list of every row)
for tag, tire, params, z, terr, T, N, shift in (("roll_nn_flat", 0, P, Z_NN, 0, 120, 4, 0.0), ("roll_nn_bumpy", 0, P, Z_NN, 2, 120, 3, 1.9),
                                                ("roll_nn_flat600", 0, P, Z_NN, 0, 600, 3, 0.0), ("roll_bekker_flat", 1, PB, Z_BK, 0, 100, 3, 0.0)):
    ws = [window(T, z, shift) for _ in range(N)]
    if tire == 1:
        for w in ws:  # keep the Bekker drives away from the w R = vx sign switch (BekkerDynamics.cpp:93)
            w[:, 1:3] = np.clip(w[:, 1:3], 1.0, 8.0)
    x0, ctrl, gt = soa(ws)
    xl = np.zeros((T, 23, N))
    loss = np.zeros(N)
    for n, w in enumerate(ws):
        loss[n], x = pr.evaluate_trajectory(tire, params, terr, w)
        xl[:, :, n] = x
    out[tag + "_x0"], out[tag + "_ctrl"], out[tag + "_gt"], out[tag + "_eval_loss"] = x0, ctrl, gt, loss
    out[tag + "_xlist"] = xl if T <= 120 else xl[::50].copy()  # long windows: every 50th row (+ the last below)
    out[tag + "_xfinal"] = xl[-1, :21, :].copy()

# ---- Trainer::trainTrajectory: loss + gradient of the reference's own tape
p_pert = P.copy()
p_pert[:104] += rng.normal(0, 0.03, 104)
for tag, params, terr, T, N, shift in (("grad_flat20", P, 0, 20, 4, 0.0), ("grad_flat200", P, 0, 200, 2, 0.0), ("grad_slope20", p_pert, 1, 20, 2, 0.0),
                                       ("grad_bumpy20", p_pert, 2, 20, 3, 2.1)):
    ws = [window(T, Z_NN, shift) for _ in range(N)]
    x0, ctrl, gt = soa(ws)
    loss, grad = np.zeros(N), np.zeros((416, N))
    for n, w in enumerate(ws):
        loss[n], grad[:, n], _ = pr.train_trajectory(params, terr, w)
    out[tag + "_params"], out[tag + "_x0"], out[tag + "_ctrl"], out[tag + "_gt"] = params, x0, ctrl, gt
    out[tag + "_loss"], out[tag + "_grad"] = loss, grad

# ---- one RMSProp update (Trainer::updateParams)
g = out["grad_flat20_grad"].sum(axis=1) / 4
out["rmsprop_grad"] = g
out["rmsprop_params"], out["rmsprop_sq"] = pr.update_params(P, np.ones(416), g)

dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_ref_v1.npz")
np.savez_compressed(dst, **out)
print("wrote", dst, os.path.getsize(dst), "bytes;", len(out), "arrays")
