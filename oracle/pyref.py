"""ctypes binding of oracle/_ref/libjackal_ref.so — the reference's OWN first-party sources compiled in place from
/root/reference behind the stand-in headers of oracle/refshim/ (see oracle/ref_capi.cpp and oracle/Makefile target `ref`).

TEST INFRASTRUCTURE ONLY.  Used to pin the restatement (oracle/jackal_oracle.hpp) and to generate the golden fixtures under
tests/golden/; never imported by the product package.  /root/reference exists in the build container only: `available()`
says whether the prebuilt library is there (it travels to the GPU box with the snapshot), `build()` needs the sources.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libjackal_ref.so")
REF = "/root/reference"

TIRE_NN, TIRE_BEKKER = 0, 1
FLAT, SLOPE, BUMPY = 0, 1, 2


def build():
    if os.path.isdir(os.path.join(REF, "src")):
        subprocess.check_call(["make", "-C", _HERE, "-j8", "ref"], stdout=subprocess.DEVNULL)
    return _SO


def available():
    return os.path.exists(_SO)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not available():
            build()
        L = C.CDLL(_SO)
        L.ref_altitude.restype = C.c_double
        L.ref_altitude.argtypes = [C.c_int, C.c_double, C.c_double]
        L.ref_loss.restype = C.c_double
        L.ref_train_trajectory.restype = C.c_double
        L.ref_evaluate_trajectory.restype = C.c_double
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def num_params(tire):
    return lib().ref_num_params(tire)


def default_params(tire):
    p = np.zeros(num_params(tire))
    lib().ref_default_params(tire, _p(p))
    return p


def whole_body_com():
    c = np.zeros(3)
    lib().ref_whole_body_com(_p(c))
    return c


def altitude(terr, x, y):
    return lib().ref_altitude(terr, float(x), float(y))


def tire_nn_forward(p416, in8):
    p, i = _f64(p416), _f64(in8)
    o = np.zeros(4)
    lib().ref_tire_nn_forward(_p(p), _p(i), _p(o))
    return o


def bekker_forces(inputs4, soil5):
    i, s = _f64(inputs4), _f64(soil5)
    f, w = np.zeros(3), np.zeros(4)
    lib().ref_bekker_forces(_p(i), _p(s), _p(f), _p(w))
    return f, w


def fd(base_v6, g6, qd4, fext5x6):
    v, g, qd, fe = _f64(base_v6), _f64(g6), _f64(qd4), _f64(fext5x6)
    a, qdd = np.zeros(6), np.zeros(4)
    lib().ref_fd(_p(v), _p(g), _p(qd), _p(fe), _p(a), _p(qdd))
    return a, qdd


def fancy_sinkages(terr, X21):
    """HybridDynamics::get_tire_cpts_sinkages_fancy -> sinkages [4], reported points [4][3]"""
    X = _f64(X21)
    sk, pts = np.zeros(4), np.zeros((4, 3))
    lib().ref_fancy_sinkages(terr, _p(X), _p(sk), _p(pts))
    return sk, pts


def base_network_forward(p131, in3):
    p, i = _f64(p131), _f64(in3)
    o = np.zeros(3)
    lib().ref_base_network_forward(_p(p), _p(i), _p(o))
    return o


def ode(tire, params, terr, X21, u2):
    x = np.concatenate([_f64(X21), _f64(u2)])
    p = _f64(params)
    xd = np.zeros(23)
    lib().ref_ode(tire, _p(p), terr, _p(x), _p(xd))
    return xd[:21]


def integrate(tire, params, terr, X23):
    x, p = _f64(X23), _f64(params)
    o = np.zeros(23)
    lib().ref_integrate(tire, _p(p), terr, _p(x), _p(o))
    return o


def default_initial_state(tire, params, terr):
    p = _f64(params)
    s = np.zeros(21)
    lib().ref_default_initial_state(tire, _p(p), terr, _p(s))
    return s


def initialize_state(row10):
    r = _f64(row10)
    x = np.zeros(23)
    lib().ref_initialize_state(_p(r), _p(x))
    return x


def loss(gt21, x23):
    g, x = _f64(gt21), _f64(x23)
    return lib().ref_loss(_p(g), _p(x))


def rollout(tire, params, terr, x0_23, ctrl, T):
    """x0_23 [23], ctrl [T-1][2] -> x_list [T][23]"""
    p, x, c = _f64(params), _f64(x0_23), _f64(ctrl)
    xl = np.zeros((T, 23))
    lib().ref_rollout(tire, _p(p), terr, T, _p(x), _p(c), _p(xl))
    return xl


def train_trajectory(p416, terr, rows):
    """rows [T][10] = (time, vl, vr, x, y, z, yaw, wz, vx, vy) -> loss, grad[416], x_list[T][23]  (Trainer::trainTrajectory)"""
    p, r = _f64(p416), _f64(rows)
    T = r.shape[0]
    xl, g = np.zeros((T, 23)), np.zeros(416)
    L = lib().ref_train_trajectory(_p(p), terr, _p(r), T, _p(xl), _p(g))
    return L, g, xl


def evaluate_trajectory(tire, params, terr, rows):
    p, r = _f64(params), _f64(rows)
    T = r.shape[0]
    xl = np.zeros((T, 23))
    L = lib().ref_evaluate_trajectory(tire, _p(p), terr, _p(r), T, _p(xl))
    return L, xl


def update_params(params, sq, grad):
    p, s, g = _f64(params).copy(), _f64(sq).copy(), _f64(grad)
    lib().ref_update_params(_p(p), _p(s), _p(g), p.size)
    return p, s


def trainer_eq_state(tire, terr):
    z, q = C.c_double(), np.zeros(4)
    lib().ref_trainer_eq_state(tire, terr, C.byref(z), _p(q))
    return z.value, q
