"""Independent numpy restatement of the Bekker tire model (reference src/auvsl_dynamics/BekkerTireModel.cpp:18-278 and the
per-wheel glue of BekkerDynamics.cpp:59-106).  TEST INFRASTRUCTURE, not product code; companion of tests/torch_restatement.py.

Different in structure from the oracle (a statement-by-statement scalar walk) and from the kernels (4-wide branch-free node
routine with its own log / exp / rsqrt): here every arc is evaluated as ONE numpy expression over its node vector
theta_k = lower + k dtheta (a product, not the reference's running sum), with numpy's libm, and the trapezoid rule is a dot
product.  Agreement with the oracle is therefore expected at rounding level (1e-12), not bit for bit."""
import numpy as np

R, B, K, C_SOIL, PG, K0, AU, NUM_STEPS = 0.098, 0.05, 0.0254, 8.62, 100.0, 2195.0, 192400.0, 10


def get_features(zr, vx, vy, w_r):
    """BekkerTireModel::get_features (:185-209)"""
    vx_c = vx if abs(vx) > 1e-3 else 1e-3
    wr_c = w_r if abs(w_r) > 1e-3 else 1e-3
    return zr, 1.0 - vx_c / wr_c, np.tanh(vy / abs(vx_c))


def _arc_nodes(upper, lower):
    """nodes and trapezoid weights of BekkerTireModel::integrate (:163-179): 9 panels from the lower bound, then the panel
    that ends exactly on the upper bound"""
    d = (upper - lower) / NUM_STEPS
    if not (lower < upper - 0.1 * d - d):  # the loop body never runs (zero-width arc)
        th = np.array([upper - d, upper])
        return th, np.array([0.5 * d, 0.5 * d])
    k = np.arange(NUM_STEPS)               # theta_0 .. theta_9
    th = np.concatenate([lower + k * d, [upper - d, upper]])
    w = np.zeros(NUM_STEPS + 2)
    w[0:NUM_STEPS - 1] += 0.5 * d          # left ends of panels 0..8
    w[1:NUM_STEPS] += 0.5 * d              # right ends of panels 0..8
    w[NUM_STEPS] += 0.5 * d
    w[NUM_STEPS + 1] += 0.5 * d
    return th, w


def get_forces(zr_in, slip_ratio, slip_angle, kc, kphi, n0, n1, phi):
    """BekkerTireModel::get_forces (:212-278): (Fx, Fy, Fz) in N"""
    zr = min(zr_in, R)
    n = n0 + n1 * abs(slip_ratio)
    kk = kc / B + kphi
    pgc = kk * zr ** n
    ze = zr if pgc < PG else (PG / kk) ** (1.0 / n)
    ku = K0 + AU * ze
    zu = (kk / ku) ** (1.0 / n) * ze
    tf, tr, tc = np.arccos(1 - zr / R), -np.arccos(1 - (zu + zr - ze) / R), np.arccos(1 - (zr - ze) / R)
    om, tsa, tphi = 1.0 - slip_ratio, np.tan(slip_angle), np.tan(phi)

    def shear(jx, jy, sigma):
        j = np.sqrt(jy ** 2 + jx ** 2)
        j = np.where(j == 0.0, 1e-6, j)
        s = (C_SOIL + sigma * tphi) * (1.0 - np.exp(-j / K))
        return -jx / j * s, -jy / j * s

    # front arc [tc, tf]
    th, w = _arc_nodes(tf, tc)
    sig = kk * (R * (np.cos(th) - np.cos(tf))) ** n
    jx = R * ((th - tf) - om * (np.sin(th) - np.sin(tf)))
    jy = R * om * tsa * (tf - th)
    tx, ty = shear(jx, jy, sig)
    fx_f, fy_f, fz_f = w @ (tx * np.cos(th) - sig * np.sin(th)), w @ ty, w @ (sig * np.cos(th) + tx * np.sin(th))
    # rear arc [tr, -tc]
    th, w = _arc_nodes(-tc, tr)
    sig = ku * (R * (np.cos(th) - np.cos(tr))) ** n
    jx = R * ((2 * tc + th - tf) - om * (np.sin(th) - np.sin(tf)) - 2 * np.sin(tc))
    jy = R * om * tsa * (tf - th - 2 * tc + 2 * np.sin(tc))
    tx, ty = shear(jx, jy, sig)
    fx_r, fy_r, fz_r = w @ (tx * np.cos(th) - sig * np.sin(th)), w @ ty, w @ (sig * np.cos(th) + tx * np.sin(th))
    # centre arc [-tc, tc] (flat bottom at depth ze)
    th, w = _arc_nodes(tc, -tc)
    sig = kk * ze ** n * np.ones_like(th)
    jx = R * ((tc - tf) + om * np.sin(tf) - np.sin(tc) + slip_ratio * np.cos(tc) * np.tan(th))
    jy = R * om * tsa * (tf - th + np.sin(tc) - np.cos(tc) * np.tan(th))
    tx, ty = shear(jx, jy, sig)
    sec2 = (1.0 / np.cos(th)) ** 2
    fx_c, fy_c = w @ (tx * sec2), w @ (ty * sec2)
    bt = R * B
    fx = bt * fx_f + bt * np.cos(tc) * fx_c + bt * fx_r
    fy = bt * fy_f + bt * np.cos(tc) * fy_c + bt * fy_r
    fz = bt * fz_f + bt * fz_r + 2 * bt * np.sin(tc) * PG
    return 1000.0 * fx, 1000.0 * fy, 1000.0 * fz


def wheel_force(p5, sink, cvx, cvy, cvz, wheel_w):
    """BekkerDynamics::get_tire_f_ext for one wheel (:59-106): soil 5-vector as passed to setParams -> (Fx, Fy, Fz + damping)"""
    kc, kphi, n0, n1, phi = 100.0 * p5[0], 1000.0 * p5[1], p5[2], max(p5[3], 0.0), p5[4]
    w_r = 0.098 * wheel_w
    zr, sr, sa = get_features(sink, cvx, cvy, w_r)
    fx, fy, fz = get_forces(zr, sr, sa, kc, kphi, n0, n1, phi) if sink > 0.0 else (0.0, 0.0, 0.0)
    fx = abs(fx) if (w_r - cvx) > 0.0 else -abs(fx)
    return fx, fy, fz + (-200.0) * cvz
