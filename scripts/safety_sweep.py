"""The reference's `safety_colormap` sweep (test/test_Terrain.cpp:71-122, 213-346) as ONE batched rollout.

The reference drives the vehicle over BumpyTerrainMap for 1000 rows (10 ms each) with every (vl, vr) of a
10 x 10 grid, 0..9 rad/s, one rollout after the other, and records the largest tilt angle
acos((R z)_z) of the body z axis; a rollout "fails" when the angle exceeds 0.7853 rad.  Here the 100 rollouts are
one `avs_rollout` call (N = 100, T = rows + 1); the tilt and the failure row are read from x_list afterwards.

    python scripts/safety_sweep.py [--rows 1000] [--terrain bumpy|flat|slope]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from auvsl_neural_ode_b200 import capi

MAX_TILT = 0.7853  # test_Terrain.cpp:116


def sweep_inputs(x_settled, rows, max_i=10, max_j=10, vel_scale=10.0):
    """x0 [21][N], ctrl [rows][2][N] of the (vl, vr) grid; sample index = i * max_j + j, vl = vel_scale*i/max_i."""
    N = max_i * max_j
    x0 = np.repeat(np.asarray(x_settled, dtype=np.float64)[:, None], N, axis=1)
    i, j = np.divmod(np.arange(N), max_j)
    vl, vr = vel_scale * i / max_i, vel_scale * j / max_j
    ctrl = np.empty((rows, 2, N))
    ctrl[:, 0, :] = vl
    ctrl[:, 1, :] = vr
    return np.ascontiguousarray(x0), ctrl


def tilt_angles(x_list):
    """acos of the zz entry of the rotation matrix (utils.cpp:45-52) for every row: [T][N]."""
    qx, qy = x_list[:, 0, :], x_list[:, 1, :]
    with np.errstate(invalid="ignore"):
        return np.arccos(np.clip(1.0 - 2.0 * (qx * qx + qy * qy), -1.0, 1.0))


def safety(x_list):
    """(max_angle [N], len [N]) with the reference's early exit: rows after the first angle > MAX_TILT do not count."""
    ang = tilt_angles(x_list)[1:]                      # the reference looks at integrate() outputs only
    ang = np.where(np.isfinite(ang), ang, np.pi)       # a blown-up rollout has certainly tipped over
    run_max = np.maximum.accumulate(ang, axis=0)
    failed = run_max > MAX_TILT
    rows = ang.shape[0]
    first = np.where(failed.any(axis=0), failed.argmax(axis=0), rows - 1)
    return run_max[first, np.arange(ang.shape[1])], first + 1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1000)
    ap.add_argument("--terrain", default="bumpy", choices=["flat", "slope", "bumpy"])
    args = ap.parse_args()
    ctx = capi.Context(0, capi.TIRE_NN)
    ctx.set_terrain({"flat": capi.FLAT, "slope": capi.SLOPE, "bumpy": capi.BUMPY}[args.terrain])
    x = ctx.settle()
    x[4] = x[5] = 0.0          # getDefaultInitialState keeps quaternion and z only (VehicleSystem.cpp:151-177)
    x[7:] = 0.0
    x0, ctrl = sweep_inputs(x, args.rows)
    x_list, _ = ctx.rollout(x0, ctrl, args.rows + 1, want_final=False)
    max_angle, length = safety(x_list)
    np.set_printoptions(precision=3, suppress=True, linewidth=160)
    print("max tilt [rad]; row = left speed 9..0 rad/s (top to bottom), column = right speed 0..9 rad/s")
    print(max_angle.reshape(10, 10)[::-1])
    print(f"{int((length < args.rows).sum())} of 100 rollouts exceed {MAX_TILT} rad before row {args.rows}")


if __name__ == "__main__":
    main()
