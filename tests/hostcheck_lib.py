"""ctypes binding of tests/_build/libhostcheck.so — the CPU build of the device math (WPT = 4)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libhostcheck.so")


class Model(C.Structure):
    _fields_ = [("tire", C.c_int), ("terr", C.c_int), ("nx", C.c_int), ("ny", C.c_int), ("x0", C.c_double), ("y0", C.c_double),
                ("dx", C.c_double), ("dy", C.c_double), ("params", C.c_void_p), ("zw", C.c_void_p), ("height", C.c_void_p),
                ("soil", C.c_void_p)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        subprocess.check_call(["make", "-C", os.path.join(_HERE, "hostcheck")], stdout=subprocess.DEVNULL)
        _lib = C.CDLL(_SO)
        _lib.hc_tanh.restype = C.c_double
        _lib.hc_tanh.argtypes = [C.c_double]
        _lib.hc_tanh_vec.restype = C.c_double
        _lib.hc_tanh_vec.argtypes = [C.c_double]
        _lib.hc_ztab.restype = C.c_double
        _lib.hc_ztab.argtypes = [C.c_double, C.c_void_p]
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def set_contact_search(on):
    """10-angle contact search (run-time-terrain instantiation of the device math) for the following ode / rollout calls"""
    lib().hc_set_contact_search(1 if on else 0)


def zweights():
    z = np.zeros(152)
    lib().hc_zweights(_p(z))
    return z


def ztab(s0):
    """(z4, d z4 / d s0) of the fixed z-network from the device's piecewise-polynomial table (ztab_eval, avs_model.cuh)"""
    d = np.zeros(1)
    v = lib().hc_ztab(float(s0), _p(d))
    return v, float(d[0])


def default_params():
    p = np.zeros(416)
    lib().hc_default_params(_p(p))
    return p


def aiinv():
    a = np.zeros((6, 6))
    lib().hc_aiinv(_p(a))
    return a


def model(tire=0, terr=0, params=None, height=None, soil=None, x0=0.0, y0=0.0, dx=1.0, dy=1.0):
    m = Model()
    m.tire, m.terr = tire, terr
    m._keep = []
    p = _f64(params if params is not None else default_params())
    z = zweights()
    m._keep += [p, z]
    m.params, m.zw = p.ctypes.data, z.ctypes.data
    m.x0, m.y0, m.dx, m.dy = x0, y0, dx, dy
    if height is not None:
        h = _f64(height)
        m.ny, m.nx = h.shape
        m.height = h.ctypes.data
        m._keep.append(h)
        if soil is not None:
            s = _f64(soil)
            m.soil = s.ctypes.data
            m._keep.append(s)
    return m


LIVE = [0, 1, 2, 3, 4, 5, 6, 11, 12, 13, 14, 15, 16]


def ode(m, X13, u, ws4=None):
    X13, u = _f64(X13), _f64(u)
    out = np.zeros(13)
    ws = None if ws4 is None else _f64(ws4)
    lib().hc_ode(C.byref(m), _p(X13), _p(u), _p(out), None if ws is None else _p(ws))
    return out


def rk4(m, X13, u):
    X13, u = _f64(X13), _f64(u)
    out, st = np.zeros(13), np.zeros((4, 13))
    lib().hc_rk4(C.byref(m), _p(X13), _p(u), _p(out), _p(st))
    return out, st


def ode_vjp(m, X13, u, Xdb):
    X13, u, Xdb = _f64(X13), _f64(u), _f64(Xdb)
    xb, g = np.zeros(13), np.zeros(104)
    lib().hc_ode_vjp(C.byref(m), _p(X13), _p(u), _p(Xdb), _p(xb), _p(g))
    return xb, g


def ode_vjp_layered(m, X13, u, Xdb):
    X13, u, Xdb = _f64(X13), _f64(u), _f64(Xdb)
    xb, g = np.zeros(13), np.zeros(104)
    lib().hc_ode_vjp_layered(C.byref(m), _p(X13), _p(u), _p(Xdb), _p(xb), _p(g))
    return xb, g


def rollout(m, x0, ctrl, T, substeps=10, want_list=True):
    x0, ctrl = _f64(x0), _f64(ctrl)
    N = x0.shape[1]
    xl = np.zeros((T, 23, N)) if want_list else None
    xf = np.zeros((21, N))
    lib().hc_rollout(C.byref(m), N, T, substeps, _p(x0), _p(ctrl), _p(xl) if want_list else None, _p(xf))
    return xl, xf


def rollout_margins(m, x0, ctrl, T, substeps=10):
    x0, ctrl = _f64(x0), _f64(ctrl)
    N = x0.shape[1]
    xl, xf, mg = np.zeros((T, 23, N)), np.zeros((21, N)), np.zeros((6, N))
    lib().hc_rollout_margins(C.byref(m), N, T, substeps, _p(x0), _p(ctrl), _p(xl), _p(xf), _p(mg))
    return xl, xf, mg


def rollout_grad(m, x0, ctrl, gt, T, substeps=10):
    x0, ctrl, gt = _f64(x0), _f64(ctrl), _f64(gt)
    N = x0.shape[1]
    loss, gps = np.zeros(N), np.zeros((104, N))
    lib().hc_rollout_grad(C.byref(m), N, T, substeps, _p(x0), _p(ctrl), _p(gt), _p(loss), _p(gps))
    return loss, gps
