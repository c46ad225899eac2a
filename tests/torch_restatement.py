"""Independent restatement of the NN-tire rollout + training objective in torch fp64 with autograd.

TEST INFRASTRUCTURE, not product code.  Purpose: a second engine, by a different route, for what pins the gradients.  The
oracle (oracle/jackal_oracle.hpp) and the compiled reference sources (oracle/_ref) both differentiate with this repo's own
reverse tape (CppAD is absent and stood in for); this file shares neither code nor structure with them:

  * dense, generic spatial algebra (6x6 matrices built from first principles: Pluecker transforms, spatial inertias, crm/crf,
    articulated-body inertias by matrix products, a linear solve per evaluation) instead of the statement-by-statement walk of
    the generated forward_dynamics.cpp or the constant-folded kernels;
  * batched tensors [N, ...] instead of scalar code; torch.autograd instead of a tape of scalar nodes.

It follows the MATH of SURVEY.md Appendix A (A1-A8), which cites: generated/model_constants.h:24-100, transforms.cpp:194-239,
forward_dynamics.cpp:30-178, HybridDynamics.cpp:214-253, 321-417, 461-502, 519-631, TireNetwork.cpp:54-119, utils.cpp:20-52,
VehicleSystem.cpp:70-92, 113-148, Trainer.cpp:612-649.  Network weights are data and are read from csrc/avs_weights.h.
"""
import os
import re

import numpy as np
import torch

F64 = torch.float64
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# ---- A1: constants ----------------------------------------------------------------------------------------------------------
R_TIRE = 0.098
TX, TY, TZ = 0.13099999725818634, 0.1877949982881546, 0.03449999913573265
WHEEL_T = [(+TX, +TY, TZ), (+TX, -TY, TZ), (-TX, +TY, TZ), (-TX, -TY, TZ)]  # FL FR RL RR
M_B, COM_B = 16.52400016784668, (0.011999273672699928, 0.0, 0.06699594855308533)
IB = dict(xx=0.38783785700798035, yy=0.46875107288360596, zz=0.4509454071521759, xy=0.0011965520679950714, xz=-0.003115505911409855,
          yz=0.0031140821520239115)
M_W = 0.47699999809265137
IW = dict(xx=0.0013000000035390258, yy=0.0013000000035390258, zz=0.002400000113993883, xy=0.0, xz=0.0, yz=4.808253448174149e-11)
GRAVITY, DT = -9.81, 1e-3
IN_STD = (0.0008373582968488336, 0.5799984931945801, 0.576934278011322, 0.5774632096290588)
OUT_STD = (17.241097080572587, 15.9103406661298, 12.632573461434308)
AZ = 2  # joint axis index in [AX AY AZ LX LY LZ]


def _weights_from_header():
    txt = open(os.path.join(ROOT, "auvsl_neural_ode_b200", "csrc", "avs_weights.h")).read()
    out = {}
    for name, body in re.findall(r"static const double (\w+)\[\d+\] = \{(.*?)\};", txt, flags=re.S):
        out[name] = np.array([float(v) for v in body.replace("\n", " ").split(",") if v.strip()])
    return out


_W = _weights_from_header()
Z0 = torch.tensor(_W["kFixedZ0"], dtype=F64).reshape(8, 1)
Z2 = torch.tensor(_W["kFixedZ2"], dtype=F64).reshape(8, 8)
Z4 = torch.tensor(_W["kFixedZ4"], dtype=F64).reshape(1, 8)


def default_live_params():
    """the 104 live entries of the 416-vector (network 0: W0 | W2 | W4, row-major)"""
    return np.concatenate([_W["kDefaultW0"], _W["kDefaultW2"], _W["kDefaultW4"]])


def skew(v):
    x, y, z = v
    return torch.tensor([[0.0, -z, y], [z, 0.0, -x], [-y, x, 0.0]], dtype=F64)


def spatial_inertia(m, c, I):
    Ibar = torch.tensor([[I["xx"], -I["xy"], -I["xz"]], [-I["xy"], I["yy"], -I["yz"]], [-I["xz"], -I["yz"], I["zz"]]], dtype=F64)
    S = torch.zeros(6, 6, dtype=F64)
    S[:3, :3] = Ibar
    S[:3, 3:] = m * skew(c)
    S[3:, :3] = -m * skew(c)
    S[3:, 3:] = m * torch.eye(3, dtype=F64)
    return S


# ---- A2: wheel motion transforms at q = 0 -----------------------------------------------------------------------------------
E_WHEEL = torch.tensor([[1.0, 0.0, 0.0], [0.0, 0.0, -1.0], [0.0, 1.0, 0.0]], dtype=F64)


def wheel_X(t):
    X = torch.zeros(6, 6, dtype=F64)
    X[:3, :3] = E_WHEEL
    X[3:, :3] = -E_WHEEL @ skew(t)
    X[3:, 3:] = E_WHEEL
    return X


I_B = spatial_inertia(M_B, COM_B, IB)
I_W = spatial_inertia(M_W, (0.0, 0.0, 0.0), IW)
X_W = [wheel_X(t) for t in WHEEL_T]
U_W = I_W[:, AZ].clone()
D_W = U_W[AZ].clone()
IA_W = I_W - torch.outer(U_W, U_W) / D_W
AI_B = I_B + sum(X.T @ IA_W @ X for X in X_W)


def crm_cols(v):
    """batched spatial motion cross-product matrix crm(v): [N,6] -> [N,6,6]"""
    N = v.shape[0]
    w, l = v[:, :3], v[:, 3:]

    def sk(a):
        z = torch.zeros(N, dtype=F64)
        return torch.stack([torch.stack([z, -a[:, 2], a[:, 1]], 1), torch.stack([a[:, 2], z, -a[:, 0]], 1), torch.stack([-a[:, 1], a[:, 0], z], 1)], 1)

    M = torch.zeros(N, 6, 6, dtype=F64)
    M[:, :3, :3] = sk(w)
    M[:, 3:, :3] = sk(l)
    M[:, 3:, 3:] = sk(w)
    return M


def crf_apply(v, f):
    """v x* f = -crm(v)^T f"""
    return -(crm_cols(v).transpose(1, 2) @ f.unsqueeze(-1)).squeeze(-1)


def rot_from_quat(q):
    x, y, z, w = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    return torch.stack([
        torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], 1),
        torch.stack([2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)], 1),
        torch.stack([2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], 1)], 1)


def altitude(kind, x, y):
    if kind == "flat":
        return torch.zeros_like(x)
    if kind == "slope":
        return 0.1 * x
    raise ValueError(kind)


# ---- A7: tire network --------------------------------------------------------------------------------------------------------
def tire_net(p104, vx, vy, wheel_w, sink):
    """[N] each -> Fx, Fy, Fz [N]"""
    W0, W2, W4 = p104[:24].reshape(8, 3), p104[24:88].reshape(8, 8), p104[88:104].reshape(2, 8)
    s = torch.stack([sink / IN_STD[0], (wheel_w * R_TIRE - vx) / IN_STD[1], wheel_w / IN_STD[2], vy / IN_STD[3]], 1)  # [N,4]
    xy = torch.tanh(torch.tanh(s[:, 1:4] @ W0.T) @ W2.T) @ W4.T
    zz = (torch.tanh(torch.tanh(s[:, 0:1] @ Z0.T) @ Z2.T) @ Z4.T).squeeze(-1)
    zz = torch.where(zz > 0, zz, torch.zeros_like(zz))
    return xy[:, 0] * OUT_STD[0], xy[:, 1] * OUT_STD[1], zz * OUT_STD[2]


# ---- A3 + A4: one ODE evaluation ------------------------------------------------------------------------------------------------
def bekker_wheel_forces(p5):
    """tire function for ode(): the Bekker model of tests/bekker_restatement.py (numpy, values only — the reference's Bekker
    gradient is not finite, so there is nothing to differentiate against)"""
    import bekker_restatement as br

    def fn(sink, cv, wheel_w):
        out = np.zeros((sink.shape[0], 3))
        for n in range(sink.shape[0]):
            out[n] = br.wheel_force(p5, float(sink[n]), float(cv[n, 0]), float(cv[n, 1]), float(cv[n, 2]), float(wheel_w[n]))
        t = torch.tensor(out, dtype=F64)
        return t[:, 0], t[:, 1], t[:, 2]
    return fn


def ode(p104, X, u, terrain="flat", tire_fn=None):
    """X [N,21], u [N,2] -> Xd [N,21]; tire_fn(sink, cv, wheel_w) -> (Fx, Fy, Fz incl. damping) replaces the tire network"""
    N = X.shape[0]
    q, pos, w, v = X[:, 0:4], X[:, 4:7], X[:, 11:14], X[:, 14:17]
    qd = torch.stack([u[:, 0], u[:, 1], u[:, 0], u[:, 1]], 1)
    Rm = rot_from_quat(q)
    vb = torch.cat([w, v], 1)
    pb = crf_apply(vb, vb @ I_B.T)
    for i in range(4):
        e = torch.tensor([WHEEL_T[i][0], WHEEL_T[i][1], WHEEL_T[i][2] - R_TIRE], dtype=F64)
        cp = (Rm @ e) + pos
        sink = altitude(terrain, cp[:, 0], cp[:, 1]) - cp[:, 2]
        cv = v + torch.cross(w, e.expand(N, 3), dim=1)  # v - skew(e) w
        if tire_fn is None:
            fx, fy, fz = tire_net(p104, cv[:, 0], cv[:, 1], X[:, 17 + i], sink)
            fz = fz + (-20.0) * cv[:, 2]
        else:
            fx, fy, fz = tire_fn(sink, cv, X[:, 17 + i])
        zero = torch.zeros_like(fx)
        fext = torch.stack([zero, zero, zero, fx, -fz, fy], 1)
        vi = vb @ X_W[i].T
        vi = vi + torch.nn.functional.one_hot(torch.tensor(AZ), 6).to(F64) * qd[:, i:i + 1]
        ci = crm_cols(vi)[:, :, AZ] * qd[:, i:i + 1]
        pi = -fext + crf_apply(vi, vi @ I_W.T)
        ui = -pi[:, AZ]
        pa = pi + ci @ IA_W.T + U_W * (ui / D_W).unsqueeze(-1)
        pb = pb + pa @ X_W[i]  # X^T pa
    ab = -torch.linalg.solve(AI_B.expand(N, 6, 6), pb.unsqueeze(-1)).squeeze(-1)
    g_world = torch.tensor([0.0, 0.0, GRAVITY], dtype=F64)
    ab = ab + torch.cat([torch.zeros(N, 3, dtype=F64), Rm.transpose(1, 2) @ g_world], 1)
    x, y, z, ww = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    Q = torch.stack([torch.stack([ww, -z, y], 1), torch.stack([z, ww, -x], 1), torch.stack([-y, x, ww], 1), torch.stack([-x, -y, -z], 1)], 1)
    qdot = 0.5 * (Q @ w.unsqueeze(-1)).squeeze(-1)
    pdot = (Rm @ v.unsqueeze(-1)).squeeze(-1)
    return torch.cat([qdot, pdot, qd, ab, torch.zeros(N, 4, dtype=F64)], 1)


# ---- A5 + A6: RK4 with quaternion renormalisation, macro step -----------------------------------------------------------------------
def _renorm(X):
    return torch.cat([X[:, :4] / X[:, :4].norm(dim=1, keepdim=True), X[:, 4:]], 1)


def rk4(p104, X, u, terrain="flat", tire_fn=None):
    k1 = ode(p104, X, u, terrain, tire_fn)
    k2 = ode(p104, _renorm(X + 0.5 * DT * k1), u, terrain, tire_fn)
    k3 = ode(p104, _renorm(X + 0.5 * DT * k2), u, terrain, tire_fn)
    k4 = ode(p104, _renorm(X + DT * k3), u, terrain, tire_fn)
    return _renorm(X + (DT / 6.0) * (k1 + 2 * k2 + 2 * k3 + k4))


def macro_step(p104, X, u, terrain="flat", substeps=10, tire_fn=None):
    for _ in range(substeps):
        X = torch.cat([X[:, :17], u[:, 0:1], u[:, 1:2], u[:, 0:1], u[:, 1:2]], 1)
        X = rk4(p104, X, u, terrain, tire_fn)
    return X


# ---- A8: loss and training objective ---------------------------------------------------------------------------------------------
def row_loss(X, yaw_gt, x_gt, y_gt):
    x, y, z, w = X[:, 0], X[:, 1], X[:, 2], X[:, 3]
    yaw = torch.atan2(2 * (w * z + x * y), 1 - 2 * (y * y + z * z))
    e = yaw - yaw_gt
    return torch.atan2(torch.sin(e), torch.cos(e)).abs() + torch.sqrt((x_gt - X[:, 4]) ** 2 + (y_gt - X[:, 5]) ** 2)


def rollout_loss(p104, x0, ctrl, gt, terrain="flat"):
    """x0 [21][N], ctrl [T-1][2][N], gt [T][3][N] (the C-ABI's layouts) -> per-sample loss [N], states [T][21][N]"""
    T = gt.shape[0]
    X = torch.as_tensor(x0, dtype=F64).T.contiguous()
    ctrl_t, gt_t = torch.as_tensor(ctrl, dtype=F64), torch.as_tensor(gt, dtype=F64)
    states, loss = [X], torch.zeros(X.shape[0], dtype=F64)
    traj_len = torch.zeros(X.shape[0], dtype=F64)
    for i in range(1, T):
        X = macro_step(p104, X, ctrl_t[i - 1].T, terrain)
        states.append(X)
        loss = loss + row_loss(X, gt_t[i, 0], gt_t[i, 1], gt_t[i, 2])
        traj_len = traj_len + torch.sqrt((gt_t[i, 1] - gt_t[i - 1, 1]) ** 2 + (gt_t[i, 2] - gt_t[i - 1, 2]) ** 2)
    traj_len = torch.where(traj_len == 0, torch.ones_like(traj_len), traj_len)
    return loss / T / traj_len, torch.stack([s.T for s in states], 0)


def loss_and_grad(p104_np, x0, ctrl, gt, terrain="flat"):
    """per-sample loss [N], per-sample gradient [104][N], states [T][21][N]"""
    N = x0.shape[1]
    p = torch.tensor(p104_np, dtype=F64, requires_grad=True)
    L, states = rollout_loss(p, x0, ctrl, gt, terrain)
    grads = np.zeros((104, N))
    for n in range(N):
        g, = torch.autograd.grad(L[n], p, retain_graph=(n + 1 < N))
        grads[:, n] = g.numpy()
    return L.detach().numpy(), grads, states.detach().numpy()
