"""ctypes binding of libavsim.so (C ABI: include/avsim.h).

Fails loudly when the CUDA library is missing — there is deliberately no fallback path.
Host entry points take numpy arrays; ``*_dev`` entry points take raw device pointers (ints, e.g.
``torch.Tensor.data_ptr()``) and enqueue on the context's own CUDA stream.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# AVS_LIB selects another build of the same library (development: A/B timing of kernel variants, the AVS_DEBUG build)
LIB_PATH = os.environ.get("AVS_LIB") or os.path.join(_HERE, "libavsim.so")

TIRE_NN, TIRE_BEKKER = 0, 1
FLAT, SLOPE, BUMPY, GRID = 0, 1, 2, 3
EXPLODE_ZERO_COMPONENT, EXPLODE_DROP_SAMPLE = 0, 1
STATE_DIM, NUM_NN_PARAMS, REDUCED_LEN, NUM_LIVE = 21, 416, 418, 104

EXPORTS = [
    "avs_create", "avs_destroy", "avs_last_error", "avs_strerror", "avs_nn_default_params", "avs_bekker_default_params",
    "avs_set_nn_params", "avs_get_nn_params", "avs_set_bekker_params", "avs_get_bekker_params", "avs_set_terrain_analytic",
    "avs_set_terrain_grid", "avs_set_checkpoint_budget", "avs_whole_body_com", "avs_init_state_com", "avs_settle", "avs_ode",
    "avs_rollout", "avs_rollout_dev", "avs_rollout_grad", "avs_rollout_grad_dev", "avs_rmsprop_step_dev", "avs_rmsprop_step", "avs_reset_optimizer",
    "avs_stream", "avs_set_stream", "avs_use_own_stream", "avs_sync", "avs_kernel_launches", "avs_last_kernel_ms", "avs_measure_fp64_peak",
    "avs_rollout_margins", "avs_comm_unique_id", "avs_comm_init_rank", "avs_comm_init_all", "avs_comm_destroy", "avs_comm_info",
    "avs_allreduce_dev", "avs_allreduce", "avs_alloc_pinned", "avs_free_pinned", "avs_set_reduced", "avs_set_contact_search", "avs_load_legacy_net_csv", "avs_mlp_forward",
]
NUM_MARGINS = 6
MARGIN_KINDS = ("contact |sinkage|", "Fx sign |wR - vx|", "slip clamps", "|pgc - pg| / pg", "|R - sinkage|", "arc |dtheta|")


class AvsError(RuntimeError):
    pass


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise AvsError(f"{LIB_PATH} not built (run `python -c 'import __graft_entry__ as g; g.build()'`); there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        vp, dbl, i32 = C.c_void_p, C.c_double, C.c_int
        L.avs_create.argtypes = [C.POINTER(vp), i32, i32]
        L.avs_destroy.argtypes = [vp]
        L.avs_destroy.restype = None
        L.avs_last_error.argtypes = [vp]
        L.avs_last_error.restype = C.c_char_p
        L.avs_strerror.argtypes = [i32]
        L.avs_strerror.restype = C.c_char_p
        for n in ("avs_nn_default_params", "avs_bekker_default_params", "avs_whole_body_com"):
            getattr(L, n).argtypes = [vp]
        for n in ("avs_set_nn_params", "avs_get_nn_params", "avs_set_bekker_params", "avs_get_bekker_params", "avs_settle"):
            getattr(L, n).argtypes = [vp, vp]
        L.avs_set_terrain_analytic.argtypes = [vp, i32]
        L.avs_set_terrain_grid.argtypes = [vp, vp, vp, i32, i32, dbl, dbl, dbl, dbl]
        L.avs_set_checkpoint_budget.argtypes = [vp, C.c_ulonglong]
        L.avs_init_state_com.argtypes = [vp, vp]
        L.avs_ode.argtypes = [vp, i32, vp, vp, vp]
        L.avs_rollout.argtypes = [vp, i32, i32, i32, vp, vp, vp, vp]
        L.avs_rollout_dev.argtypes = [vp, i32, i32, i32, vp, vp, vp, vp]
        L.avs_rollout_grad.argtypes = [vp, i32, i32, i32, vp, vp, vp, dbl, i32, vp, vp, vp]
        L.avs_rollout_grad_dev.argtypes = [vp, i32, i32, i32, vp, vp, vp, dbl, i32, vp, vp, vp]
        L.avs_rmsprop_step_dev.argtypes = [vp, vp, dbl, dbl]
        L.avs_rmsprop_step.argtypes = [vp, vp, dbl, dbl]
        L.avs_reset_optimizer.argtypes = [vp]
        L.avs_stream.argtypes = [vp]
        L.avs_stream.restype = vp
        L.avs_set_stream.argtypes = [vp, vp]
        L.avs_use_own_stream.argtypes = [vp]
        L.avs_sync.argtypes = [vp]
        L.avs_kernel_launches.argtypes = [vp]
        L.avs_kernel_launches.restype = C.c_longlong
        L.avs_last_kernel_ms.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.avs_measure_fp64_peak.argtypes = [vp, C.POINTER(dbl)]
        L.avs_rollout_margins.argtypes = [vp, i32, i32, i32, vp, vp, vp, vp, vp]
        L.avs_comm_unique_id.argtypes = [vp]
        L.avs_comm_init_rank.argtypes = [vp, vp, i32, i32]
        L.avs_comm_init_all.argtypes = [C.POINTER(vp), i32]
        L.avs_comm_destroy.argtypes = [vp]
        L.avs_comm_info.argtypes = [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]
        L.avs_allreduce_dev.argtypes = [vp, vp]
        L.avs_allreduce.argtypes = [vp, vp]
        L.avs_alloc_pinned.argtypes = [C.POINTER(vp), C.c_ulonglong]
        L.avs_free_pinned.argtypes = [vp]
        L.avs_set_reduced.argtypes = [vp, vp]
        L.avs_set_contact_search.argtypes = [vp, i32]
        L.avs_load_legacy_net_csv.argtypes = [C.c_char_p, vp]
        L.avs_mlp_forward.argtypes = [vp, i32, vp, i32, vp, vp]
        _lib = L
    return _lib


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def nn_default_params():
    p = np.zeros(NUM_NN_PARAMS)
    lib().avs_nn_default_params(_p(p))
    return p


def bekker_default_params():
    p = np.zeros(5)
    lib().avs_bekker_default_params(_p(p))
    return p


def whole_body_com():
    c = np.zeros(3)
    lib().avs_whole_body_com(_p(c))
    return c


def init_state_com(start21):
    s = _f64(start21)
    out = np.zeros(21)
    lib().avs_init_state_com(_p(s), _p(out))
    return out


MLP_BASE_NETWORK, MLP_LEGACY_TIRE = 0, 1


def load_legacy_net_csv(path):
    """data/nn_pretrained.csv -> the 435 numbers W0 (16x5), b0, W2 (16x16), b2, W4 (3x16), b4, out_mean, out_std, in_mean, in_std"""
    p = np.zeros(435)
    rc = lib().avs_load_legacy_net_csv(os.fsencode(path), _p(p))
    if rc != 0:
        raise AvsError(f"avs_load_legacy_net_csv({path}): {lib().avs_strerror(rc).decode()} (need exactly 435 numbers)")
    return p


def comm_unique_id():
    """128-byte NCCL unique id (rank 0 creates it and ships it to the other ranks by any transport)."""
    buf = C.create_string_buffer(128)
    rc = lib().avs_comm_unique_id(buf)
    if rc != 0:
        raise AvsError(f"avs_comm_unique_id: {lib().avs_strerror(rc).decode()} (libnccl.so.2 not loadable?)")
    return buf.raw


def initialize_state(yaw, x, y, z, wz, vx, vy):
    """System::initializeState (src/VehicleSystem.cpp:187-238): yaw-only quaternion, COM velocities -> base-link state."""
    s = np.zeros(21)
    s[2], s[3] = np.sin(yaw / 2.0), np.cos(yaw / 2.0)
    s[4:7] = (x, y, z)
    s[13], s[14], s[15] = wz, vx, vy
    return init_state_com(s)


class Context:
    """One simulator context per GPU (not thread-safe)."""

    def __init__(self, device=0, tire_model=TIRE_NN):
        self._h = C.c_void_p()
        rc = lib().avs_create(C.byref(self._h), int(device), int(tire_model))
        if rc != 0:
            raise AvsError(f"avs_create(device={device}) failed: {lib().avs_strerror(rc).decode()} — no usable sm_100 GPU; there is no CPU fallback")
        self.tire_model = tire_model

    def close(self):
        if self._h:
            lib().avs_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise AvsError(f"[{lib().avs_strerror(rc).decode()}] {lib().avs_last_error(self._h).decode()}")

    # ---- configuration
    def set_nn_params(self, p416):
        p = _f64(p416)
        assert p.size == NUM_NN_PARAMS
        self._ck(lib().avs_set_nn_params(self._h, _p(p)))

    def get_nn_params(self):
        p = np.zeros(NUM_NN_PARAMS)
        self._ck(lib().avs_get_nn_params(self._h, _p(p)))
        return p

    def set_bekker_params(self, p5):
        p = _f64(p5)
        assert p.size == 5
        self._ck(lib().avs_set_bekker_params(self._h, _p(p)))

    def get_bekker_params(self):
        p = np.zeros(5)
        self._ck(lib().avs_get_bekker_params(self._h, _p(p)))
        return p

    def set_terrain(self, kind):
        self._ck(lib().avs_set_terrain_analytic(self._h, int(kind)))

    def set_terrain_grid(self, height, soil=None, x0=0.0, y0=0.0, dx=1.0, dy=1.0):
        h = _f64(height)
        ny, nx = h.shape
        s = None
        if soil is not None:
            s = _f64(soil)
            assert s.shape == (5, ny, nx)
        self._ck(lib().avs_set_terrain_grid(self._h, _p(h), _p(s), nx, ny, x0, y0, dx, dy))

    def set_contact_search(self, on):
        """10-angle contact search (HybridDynamics::get_tire_cpts_sinkages_fancy) instead of the point under the hub; forward only."""
        self._ck(lib().avs_set_contact_search(self._h, 1 if on else 0))

    def mlp_forward(self, kind, params, x):
        """BaseNetwork::forward (kind 0, params[131], x [3][N]) or the nn_pretrained.csv network (kind 1, params[435], x [5][N]) -> [3][N]"""
        params, x = _f64(params), _f64(x)
        N = x.shape[1]
        out = np.zeros((3, N))
        self._ck(lib().avs_mlp_forward(self._h, int(kind), _p(params), N, _p(x), _p(out)))
        return out

    def set_checkpoint_budget(self, nbytes):
        self._ck(lib().avs_set_checkpoint_budget(self._h, int(nbytes)))

    # ---- host-buffer entry points (copies included)
    def settle(self):
        out = np.zeros(21)
        self._ck(lib().avs_settle(self._h, _p(out)))
        return out

    def ode(self, x, u):
        x, u = _f64(x), _f64(u)
        N = x.shape[1]
        xd = np.zeros((21, N))
        self._ck(lib().avs_ode(self._h, N, _p(x), _p(u), _p(xd)))
        return xd

    def rollout(self, x0, ctrl, T, substeps=10, want_list=True, want_final=True, out_list=None, out_final=None):
        x0, ctrl = _f64(x0), _f64(ctrl)
        N = x0.shape[1]
        assert x0.shape == (21, N) and (T == 1 or ctrl.shape == (T - 1, 2, N))
        xl = out_list if out_list is not None else (np.zeros((T, 23, N)) if want_list else None)
        xf = out_final if out_final is not None else (np.zeros((21, N)) if want_final else None)
        self._ck(lib().avs_rollout(self._h, N, T, substeps, _p(x0), _p(ctrl), _p(xl), _p(xf)))
        return xl, xf

    def rollout_margins(self, x0, ctrl, T, substeps=10, want_list=True):
        """Bekker only: rollout + branch margins [6][N] (kinds: MARGIN_KINDS; see include/avsim.h)."""
        x0, ctrl = _f64(x0), _f64(ctrl)
        N = x0.shape[1]
        xl = np.zeros((T, 23, N)) if want_list else None
        xf, mg = np.zeros((21, N)), np.zeros((NUM_MARGINS, N))
        self._ck(lib().avs_rollout_margins(self._h, N, T, substeps, _p(x0), _p(ctrl), _p(xl), _p(xf), _p(mg)))
        return xl, xf, mg

    def rollout_grad(self, x0, ctrl, gt, T, substeps=10, explode_thresh=100.0, explode_mode=0, per_sample=True):
        x0, ctrl, gt = _f64(x0), _f64(ctrl), _f64(gt)
        N = x0.shape[1]
        assert x0.shape == (21, N) and ctrl.shape == (T - 1, 2, N) and gt.shape == (T, 3, N)
        loss, red = np.zeros(N), np.zeros(REDUCED_LEN)
        gps = np.zeros((NUM_NN_PARAMS, N)) if per_sample else None
        self._ck(lib().avs_rollout_grad(self._h, N, T, substeps, _p(x0), _p(ctrl), _p(gt), float(explode_thresh), int(explode_mode),
                                        _p(loss), _p(red), _p(gps)))
        return loss, red, gps

    # ---- device-pointer entry points (async on the context's stream)
    def rollout_dev(self, N, T, substeps, d_x0, d_ctrl, d_x_list=0, d_x_final=0):
        self._ck(lib().avs_rollout_dev(self._h, N, T, substeps, d_x0, d_ctrl, d_x_list or None, d_x_final or None))

    def rollout_grad_dev(self, N, T, substeps, d_x0, d_ctrl, d_gt, d_loss, d_reduced, d_grad_live=0, explode_thresh=100.0, explode_mode=0):
        self._ck(lib().avs_rollout_grad_dev(self._h, N, T, substeps, d_x0, d_ctrl, d_gt, float(explode_thresh), int(explode_mode),
                                            d_loss or None, d_reduced or None, d_grad_live or None))

    def rmsprop_step_dev(self, d_reduced=0, lr=1e-3, l1_weight=0.0):
        self._ck(lib().avs_rmsprop_step_dev(self._h, d_reduced or None, float(lr), float(l1_weight)))

    def rmsprop_step(self, reduced_host, lr=1e-3, l1_weight=0.0):
        r = _f64(reduced_host)
        assert r.size == REDUCED_LEN
        self._ck(lib().avs_rmsprop_step(self._h, _p(r), float(lr), float(l1_weight)))

    # ---- multi-GPU (one NCCL rank per context; include/avsim.h "multi-GPU")
    def comm_init_rank(self, uid_bytes, world, rank):
        buf = C.create_string_buffer(bytes(uid_bytes), 128)
        self._ck(lib().avs_comm_init_rank(self._h, buf, int(world), int(rank)))

    def comm_destroy(self):
        self._ck(lib().avs_comm_destroy(self._h))

    def comm_info(self):
        w, r, v = C.c_int(), C.c_int(), C.c_int()
        self._ck(lib().avs_comm_info(self._h, C.byref(w), C.byref(r), C.byref(v)))
        return w.value, r.value, v.value

    def allreduce_dev(self, d_reduced=0):
        """In-place ncclAllReduce(sum) of 418 doubles on the context's stream (0 = the context's own reduced buffer)."""
        self._ck(lib().avs_allreduce_dev(self._h, d_reduced or None))

    def allreduce(self, want_host=True):
        out = np.zeros(REDUCED_LEN) if want_host else None
        self._ck(lib().avs_allreduce(self._h, _p(out)))
        return out

    def reset_optimizer(self):
        self._ck(lib().avs_reset_optimizer(self._h))

    @property
    def stream(self):
        return lib().avs_stream(self._h)

    def set_stream(self, cuda_stream_handle):
        """Enqueue on a caller-owned stream (0 / None = the legacy default stream)."""
        self._ck(lib().avs_set_stream(self._h, cuda_stream_handle or None))

    def use_own_stream(self):
        self._ck(lib().avs_use_own_stream(self._h))

    def sync(self):
        self._ck(lib().avs_sync(self._h))

    @property
    def kernel_launches(self):
        return int(lib().avs_kernel_launches(self._h))

    def last_kernel_ms(self):
        f, b = C.c_float(), C.c_float()
        self._ck(lib().avs_last_kernel_ms(self._h, C.byref(f), C.byref(b)))
        return f.value, b.value

    def measure_fp64_peak(self):
        t = C.c_double()
        self._ck(lib().avs_measure_fp64_peak(self._h, C.byref(t)))
        return t.value
