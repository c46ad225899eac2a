"""ctypes binding of the CPU oracle (oracle/_build/liboracle.so).

TEST INFRASTRUCTURE ONLY (parity status: oracle/README.md).  Importable from
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs — never from
the product package auvsl_neural_ode_b200.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")

TIRE_NN, TIRE_BEKKER = 0, 1
FLAT, SLOPE, BUMPY, GRID = 0, 1, 2, 3


def build(force=False):
    if force or not os.path.exists(_SO) or any(
        os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_SO)
        for f in ("oracle_capi.cpp", "jackal_oracle.hpp", "scalar.hpp")
    ):
        subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)
    return _SO


class Terrain(C.Structure):
    _fields_ = [("kind", C.c_int), ("nx", C.c_int), ("ny", C.c_int), ("x0", C.c_double), ("y0", C.c_double),
                ("dx", C.c_double), ("dy", C.c_double), ("height", C.c_void_p), ("soil", C.c_void_p), ("contact_search", C.c_int)]


def terrain(kind=FLAT, height=None, soil=None, x0=0.0, y0=0.0, dx=1.0, dy=1.0, contact_search=0):
    t = Terrain()
    t.kind = kind
    t.contact_search = int(contact_search)  # 1: HybridDynamics::get_tire_cpts_sinkages_fancy instead of the point under the hub
    t._keep = []
    if kind == GRID:
        h = np.ascontiguousarray(height, dtype=np.float64)
        t.ny, t.nx = h.shape
        t.x0, t.y0, t.dx, t.dy = x0, y0, dx, dy
        t.height = h.ctypes.data
        t._keep.append(h)
        if soil is not None:
            s = np.ascontiguousarray(soil, dtype=np.float64)
            assert s.shape == (5, t.ny, t.nx)
            t.soil = s.ctypes.data
            t._keep.append(s)
    return t


_lib = None
_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.orc_altitude.restype = C.c_double
        L.orc_altitude.argtypes = [C.POINTER(Terrain), C.c_double, C.c_double]
        L.orc_loss.restype = C.c_double
        L.orc_evaluate_trajectory.restype = C.c_double
        L.orc_train_trajectory.restype = C.c_double
        L.orc_train_trajectory_monolithic.restype = C.c_double
        L.orc_train_trajectory_dual8.restype = C.c_double
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def nn_default_params():
    p = np.zeros(416)
    lib().orc_nn_default_params(_p(p))
    return p


def bekker_default_params():
    p = np.zeros(5)
    lib().orc_bekker_default_params(_p(p))
    return p


def whole_body_com():
    c = np.zeros(3)
    lib().orc_whole_body_com(_p(c))
    return c


def altitude(t, x, y):
    return lib().orc_altitude(C.byref(t), float(x), float(y))


def tire_nn_forward(p416, in8):
    p416, in8 = _f64(p416), _f64(in8)
    out = np.zeros(4)
    lib().orc_tire_nn_forward(_p(p416), _p(in8), _p(out))
    return out


def bekker_forces(inputs4, soil5):
    inputs4, soil5 = _f64(inputs4), _f64(soil5)
    feat, f = np.zeros(3), np.zeros(4)
    lib().orc_bekker_forces(_p(inputs4), _p(soil5), _p(feat), _p(f))
    return feat, f


def ode(model, params, t, X21, u2):
    params, X21, u2 = _f64(params), _f64(X21), _f64(u2)
    out = np.zeros(21)
    lib().orc_ode(model, _p(params), C.byref(t), _p(X21), _p(u2), _p(out))
    return out


def rk4(model, params, t, X21, u2):
    params, X21, u2 = _f64(params), _f64(X21), _f64(u2)
    out = np.zeros(21)
    lib().orc_rk4(model, _p(params), C.byref(t), _p(X21), _p(u2), _p(out))
    return out


def fd(base_v6, g6, qd4, fext5x6):
    base_v6, g6, qd4, fe = _f64(base_v6), _f64(g6), _f64(qd4), _f64(fext5x6)
    a, qdd = np.zeros(6), np.zeros(4)
    lib().orc_fd(_p(base_v6), _p(g6), _p(qd4), _p(fe), _p(a), _p(qdd))
    return a, qdd


def integrate(model, params, t, X23, substeps=10):
    params, X23 = _f64(params), _f64(X23)
    out = np.zeros(23)
    lib().orc_integrate(model, _p(params), C.byref(t), int(substeps), _p(X23), _p(out))
    return out


def settle(model, params, t):
    params = _f64(params)
    out = np.zeros(21)
    lib().orc_settle(model, _p(params), C.byref(t), _p(out))
    return out


def init_state_com(start21):
    start21 = _f64(start21)
    out = np.zeros(21)
    lib().orc_init_state_com(_p(start21), _p(out))
    return out


def initialize_state(row10):
    row10 = _f64(row10)
    out = np.zeros(23)
    lib().orc_initialize_state(_p(row10), _p(out))
    return out


def loss(gt21, x23):
    gt21, x23 = _f64(gt21), _f64(x23)
    return lib().orc_loss(_p(gt21), _p(x23))


def evaluate(gt21, x23):
    gt21, x23 = _f64(gt21), _f64(x23)
    a, l = C.c_double(), C.c_double()
    lib().orc_evaluate(_p(gt21), _p(x23), C.byref(a), C.byref(l))
    return a.value, l.value


def evaluate_trajectory(model, params, t, rows):
    params, rows = _f64(params), _f64(rows)
    T = rows.shape[0]
    xl = np.zeros((T, 23))
    L = lib().orc_evaluate_trajectory(model, _p(params), C.byref(t), _p(rows), T, _p(xl))
    return L, xl


def train_trajectory(p416, t, rows):
    p416, rows = _f64(p416), _f64(rows)
    T = rows.shape[0]
    xl, g = np.zeros((T, 23)), np.zeros(416)
    L = lib().orc_train_trajectory(_p(p416), C.byref(t), _p(rows), T, _p(xl), _p(g))
    return L, g, xl


def train_trajectory_monolithic(p416, t, rows):
    p416, rows = _f64(p416), _f64(rows)
    g = np.zeros(416)
    L = lib().orc_train_trajectory_monolithic(_p(p416), C.byref(t), _p(rows), rows.shape[0], _p(g))
    return L, g


def bekker_integrate_vjp(p5, t, x23, lam21):
    """Reverse sweep through one Bekker integrate(): (x1[23], grad5, xbar21, number of non-finite entries)."""
    p5, x23, lam21 = _f64(p5), _f64(x23), _f64(lam21)
    x1, g, xb = np.zeros(23), np.zeros(5), np.zeros(21)
    lib().orc_bekker_integrate_vjp.restype = C.c_int
    bad = lib().orc_bekker_integrate_vjp(_p(p5), C.byref(t), _p(x23), _p(lam21), _p(x1), _p(g), _p(xb))
    return x1, g, xb, bad


def train_trajectory_dual8(p416, t, rows, idx8):
    p416, rows = _f64(p416), _f64(rows)
    idx = np.ascontiguousarray(idx8, dtype=np.int32)
    d = np.zeros(8)
    L = lib().orc_train_trajectory_dual8(_p(p416), C.byref(t), _p(rows), rows.shape[0], _p(idx), _p(d))
    return L, d


def rmsprop(params, sq, grad, lr=1e-3, l1=0.0):
    params, sq, grad = _f64(params).copy(), _f64(sq).copy(), _f64(grad)
    lib().orc_rmsprop(_p(params), _p(sq), _p(grad), params.size, C.c_double(lr), C.c_double(l1))
    return params, sq


def count_ops(model, params, X21, u2, what=0):
    params, X21, u2 = _f64(params), _f64(X21), _f64(u2)
    out = np.zeros(9, dtype=np.int64)
    lib().orc_count_ops(model, _p(params), _p(X21), _p(u2), what, _p(out))
    return dict(zip(["addsub", "mul", "div", "tanh", "sqrt", "trig", "exp", "pow", "other"], out.tolist()))


def rollout_batch(model, params, t, x0, ctrl, T, substeps=10, want_list=True, nthreads=1):
    """x0 [21][N], ctrl [T-1][2][N] -> x_list [T][23][N] (or None), x_final [21][N]"""
    params, x0, ctrl = _f64(params), _f64(x0), _f64(ctrl)
    N = x0.shape[1]
    xl = np.zeros((T, 23, N)) if want_list else None
    xf = np.zeros((21, N))
    lib().orc_rollout_batch(model, _p(params), C.byref(t), N, T, substeps, _p(x0), _p(ctrl),
                            _p(xl) if want_list else None, _p(xf), nthreads)
    return xl, xf


MARGIN_KINDS = ("contact |sinkage|", "Fx sign |wR - vx|", "slip clamps", "|pgc - pg| / pg", "|R - sinkage|", "arc |dtheta|")


def rollout_batch_margins(model, params, t, x0, ctrl, T, substeps=10, want_list=True, nthreads=1):
    """rollout_batch + margins [6][N]: how far every value-dependent branch of the Bekker path stayed from flipping
    (minimum over the rollout; kinds in MARGIN_KINDS, definitions in jackal_oracle.hpp)"""
    params, x0, ctrl = _f64(params), _f64(x0), _f64(ctrl)
    N = x0.shape[1]
    xl = np.zeros((T, 23, N)) if want_list else None
    xf = np.zeros((21, N))
    mg = np.zeros((6, N))
    lib().orc_rollout_batch_margins(model, _p(params), C.byref(t), N, T, substeps, _p(x0), _p(ctrl),
                                    _p(xl) if want_list else None, _p(xf), _p(mg), nthreads)
    return xl, xf, mg


def rollout_grad_batch(p416, t, x0, ctrl, gt, T, explode_thresh=100.0, explode_mode=0, nthreads=1, per_sample=True):
    """gt [T][3][N] (yaw,x,y) -> loss[N], grad_sum[416], loss_sum, n_kept, grad_per_sample[416][N]"""
    p416, x0, ctrl, gt = _f64(p416), _f64(x0), _f64(ctrl), _f64(gt)
    N = x0.shape[1]
    loss_ = np.zeros(N)
    gs = np.zeros(416)
    ls, nk = C.c_double(), C.c_longlong()
    gps = np.zeros((416, N)) if per_sample else None
    lib().orc_rollout_grad_batch(_p(p416), C.byref(t), N, T, 10, _p(x0), _p(ctrl), _p(gt), C.c_double(explode_thresh),
                                 explode_mode, _p(loss_), _p(gs), C.byref(ls), C.byref(nk),
                                 _p(gps) if per_sample else None, nthreads)
    return loss_, gs, ls.value, nk.value, gps


def fancy_sinkages(t, X21):
    X = _f64(X21)
    sk, pts = np.zeros(4), np.zeros((4, 3))
    lib().orc_fancy_sinkages(C.byref(t), _p(X), _p(sk), _p(pts))
    return sk, pts


def base_network_forward(p131, x):
    """BaseNetwork::forward (src/BaseNetwork.cpp:41-60), numpy restatement: x [3][N] = (wz, vx, vy) -> [3][N] = (nz, fx, fy);
    p131 = W0 (8x3), b0, W1 (8x8), b1, W2 (3x8), b2 row-major (:62-152)"""
    p, x = _f64(p131), np.atleast_2d(_f64(x).T).T
    W0, b0 = p[0:24].reshape(8, 3), p[24:32]
    W1, b1 = p[32:96].reshape(8, 8), p[96:104]
    W2, b2 = p[104:128].reshape(3, 8), p[128:131]
    l0 = np.tanh(W0 @ (x * 0.1) + b0[:, None])
    l1 = np.tanh(W1 @ l0 + b1[:, None])
    l2 = W2 @ l1 + b2[:, None]
    return (np.maximum(l2, 0.0) * (-x)) * 10.0


def legacy_net_forward(p435, x):
    """data/nn_pretrained.csv network (layout of scripts/pretrain.py:20-36): x [5][N] -> [3][N]"""
    p, x = _f64(p435), _f64(x)
    W0, b0 = p[0:80].reshape(16, 5), p[80:96]
    W2, b2 = p[96:352].reshape(16, 16), p[352:368]
    W4, b4 = p[368:416].reshape(3, 16), p[416:419]
    out_mean, out_std, in_mean, in_std = p[419:422], p[422:425], p[425:430], p[430:435]
    s = (x - in_mean[:, None]) / in_std[:, None]
    h = np.tanh(W2 @ np.tanh(W0 @ s + b0[:, None]) + b2[:, None])
    return (W4 @ h + b4[:, None]) * out_std[:, None] + out_mean[:, None]


def hardware_threads():
    return lib().orc_hardware_threads()
