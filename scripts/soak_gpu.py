"""Soak test of the reverse kernel's producer / consumer hand-shake: many shapes (tail lanes, shortest windows, odd
substeps), every case run twice and compared bit for bit; a hang shows up as the caller's timeout."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from auvsl_neural_ode_b200 import capi, synth

ctx = capi.Context(0, capi.TIRE_NN)
rng = np.random.default_rng(0)
t0 = time.time()
cases = 0
for terr in (capi.FLAT, capi.SLOPE, capi.BUMPY):
    ctx.set_terrain(terr)
    for N in (1, 2, 3, 7, 8, 9, 31, 33, 64, 100, 255, 257, 1000):
        for T, sub in ((2, 1), (2, 10), (3, 3), (5, 10), (17, 7), (40, 10)):
            x0, ctrl, gt = synth.make_batch(N, T, seed=1000 + cases)
            a = ctx.rollout_grad(x0, ctrl, gt, T, substeps=sub)
            b = ctx.rollout_grad(x0, ctrl, gt, T, substeps=sub)
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[2], b[2]) and np.all(np.isfinite(a[2])), (terr, N, T, sub)
            cases += 1
# repeated full-size steps
ctx.set_terrain(capi.FLAT)
x0, ctrl, gt = synth.make_batch(4096, 200)
ref = None
for it in range(25):
    loss, red, _ = ctx.rollout_grad(x0, ctrl, gt, 200, per_sample=False)
    if ref is None:
        ref = (loss.copy(), red.copy())
    assert np.array_equal(loss, ref[0]) and np.array_equal(red, ref[1]), it
print(f"soak ok: {cases} small cases x 2 + 25 full-size steps, bit-identical repeats, {time.time() - t0:.1f} s")
