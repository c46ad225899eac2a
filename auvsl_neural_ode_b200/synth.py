"""Synthetic workloads of SURVEY.md §8(d) / BASELINE.json configs 2-5 (seed 20231115).

Per trajectory: yaw0 ~ U(-pi, pi), (x, y)0 = 0, z = z_stable; smooth random wheel-speed commands
vl, vr = clip(vbar +- delta + a sin(2 pi f t + phi), -2, 12) rad/s sampled at the 100 Hz row rate;
ground truth (x, y, yaw) by integrating the reference's 2-D linear kinematic model
(/root/reference/scripts/gen_fake_data.py:8-50: [vx, vy, wz] = B [vl, vr]) at 1 kHz; initial body
velocities from the same model.
"""
import numpy as np

# NOTE: no import of .capi here — bench.py's reference arm builds its inputs with this module and must not map libavsim.so.

SEED = 20231115
# whole-body COM (generated/miscellaneous.cpp:6-39 with model_constants.h:24-100; wheel COMs are the joint origins, whose x and
# y offsets cancel over the four wheels) and the Bekker soil defaults (src/BekkerSystem.cpp:21-28)
_M_BASE, _M_WHEEL = 16.52400016784668, 0.47699999809265137
_COM_BASE = np.array([0.011999273672699928, 0.0, 0.06699594855308533])
_TZ_WHEEL = 0.03449999913573265
WHOLE_BODY_COM = (_COM_BASE * _M_BASE + np.array([0.0, 0.0, 4 * _M_WHEEL * _TZ_WHEEL])) / (_M_BASE + 4 * _M_WHEEL)
BEKKER_DEFAULT = np.array([.2976, 2.083, 0.8, 0.0, .3927])
B = np.array([[0.0472, 0.0438], [0.0036, -0.0035], [-0.1744, 0.1748]])  # gen_fake_data.py:8-10
Z_STABLE_NN = 0.0584427891  # HybridDynamics::settle with the default NN weights on flat ground


def make_batch(N, T, seed=SEED, z_stable=Z_STABLE_NN, rate_hz=100.0):
    """Returns x0 [21][N], ctrl [T-1][2][N], gt [T][3][N] (yaw, x, y) as C-contiguous float64."""
    rng = np.random.default_rng(seed)
    yaw0 = rng.uniform(-np.pi, np.pi, N)
    vbar = rng.uniform(0.0, 8.0, N)
    delta = rng.uniform(-2.0, 2.0, N)
    amp = rng.uniform(0.0, 1.0, N)
    freq = rng.uniform(0.05, 0.5, N)
    phase = rng.uniform(0.0, 2 * np.pi, (2, N))
    t = np.arange(T)[:, None] / rate_hz
    vl = np.clip(vbar + delta + amp * np.sin(2 * np.pi * freq * t + phase[0]), -2.0, 12.0)  # [T][N]
    vr = np.clip(vbar - delta + amp * np.sin(2 * np.pi * freq * t + phase[1]), -2.0, 12.0)
    ctrl = np.ascontiguousarray(np.stack([vl[:-1], vr[:-1]], axis=1))  # [T-1][2][N]

    # ground truth: 10 Euler sub-steps of 1 ms per row with the row's command held (gen_fake_data.py:26-50)
    gt = np.zeros((T, 3, N))
    x = np.zeros(N)
    y = np.zeros(N)
    yaw = yaw0.copy()
    gt[0, 0], gt[0, 1], gt[0, 2] = yaw, x, y
    for i in range(1, T):
        v = B @ np.stack([vl[i - 1], vr[i - 1]])  # [3][N]: vx, vy, wz (body frame)
        for _ in range(10):
            c, s = np.cos(yaw), np.sin(yaw)
            x = x + 1e-3 * (c * v[0] - s * v[1])
            y = y + 1e-3 * (s * v[0] + c * v[1])
            yaw = yaw + 1e-3 * v[2]
        gt[i, 0], gt[i, 1], gt[i, 2] = yaw, x, y

    v0 = B @ np.stack([vl[0], vr[0]])
    com = WHOLE_BODY_COM
    x0 = np.zeros((21, N))
    x0[2], x0[3] = np.sin(yaw0 / 2.0), np.cos(yaw0 / 2.0)
    x0[6] = z_stable
    x0[13] = v0[2]
    # HybridDynamics::initStateCOM: v_base = v_com + com x w ... expressed with the skew matrix of :118-124
    w = np.stack([np.zeros(N), np.zeros(N), v0[2]])
    vc = np.stack([v0[0], v0[1], np.zeros(N)])
    x0[14] = vc[0] + (-com[2] * w[1] + com[1] * w[2])
    x0[15] = vc[1] + (com[2] * w[0] - com[0] * w[2])
    x0[16] = vc[2] + (-com[1] * w[0] + com[0] * w[1])
    return np.ascontiguousarray(x0), ctrl, np.ascontiguousarray(gt)


def make_grid_terrain(seed=SEED, n=512, cell=0.05, with_soil=True):
    """Config 4: height = sum_8 A_k sin(k_k . x + phi_k), sum A <= 0.05 m; soil grids = Bekker defaults x U(0.8, 1.2)."""
    rng = np.random.default_rng(seed + 1)
    x0 = y0 = -0.5 * (n - 1) * cell
    xs = x0 + cell * np.arange(n)
    X, Y = np.meshgrid(xs, xs)
    A = rng.uniform(0, 1, 8)
    A = 0.05 * A / A.sum()
    kx, ky = rng.uniform(-3, 3, 8), rng.uniform(-3, 3, 8)
    ph = rng.uniform(0, 2 * np.pi, 8)
    h = sum(A[k] * np.sin(kx[k] * X + ky[k] * Y + ph[k]) for k in range(8))
    soil = None
    if with_soil:
        d = BEKKER_DEFAULT
        soil = np.empty((5, n, n))
        for k in (0, 1, 2, 4):
            soil[k] = d[k] * rng.uniform(0.8, 1.2, (n, n))
        soil[3] = rng.uniform(0.0, 0.1, (n, n))
    return dict(height=np.ascontiguousarray(h), soil=soil, x0=x0, y0=y0, dx=cell, dy=cell)
