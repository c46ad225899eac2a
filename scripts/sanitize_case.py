"""Small end-to-end case for compute-sanitizer runs (memcheck / racecheck): all kernel families once."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from auvsl_neural_ode_b200 import capi, synth
N, T = 37, 5
x0, ctrl, gt = synth.make_batch(N, T, seed=1)
nn = capi.Context(0, capi.TIRE_NN)
xl, xf = nn.rollout(x0, ctrl, T)
loss, red, gps = nn.rollout_grad(x0, ctrl, gt, T, explode_mode=1)
nn.rmsprop_step(red)
nn.set_terrain(capi.BUMPY)
loss2, red2, _ = nn.rollout_grad(x0, ctrl, gt, T, per_sample=False)
g = synth.make_grid_terrain(n=64)
bk = capi.Context(0, capi.TIRE_BEKKER)
bk.set_terrain_grid(g["height"], g["soil"], g["x0"], g["y0"], g["dx"], g["dy"])
xb, _ = bk.rollout(x0, np.ascontiguousarray(np.clip(ctrl, 1, 6)[:2]), 3)
print("sanitize case ok", float(loss.sum()), float(red2[416]), bool(np.isfinite(xb).all()))
