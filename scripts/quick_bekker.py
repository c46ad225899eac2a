"""Bekker forward throughput (configs[0]/[3] shapes, scaled) — development aid."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from auvsl_neural_ode_b200 import capi, synth
N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
T = int(sys.argv[2]) if len(sys.argv) > 2 else 60
ctx = capi.Context(0, capi.TIRE_BEKKER)
g = synth.make_grid_terrain()
ctx.set_terrain_grid(g["height"], g["soil"], g["x0"], g["y0"], g["dx"], g["dy"])
x0, ctrl, gt = synth.make_batch(N, T, z_stable=0.065)
ctrl = np.clip(ctrl, 0.5, 8.0)
dev = torch.device("cuda:0")
dx0, dctrl = torch.from_numpy(x0).to(dev), torch.from_numpy(ctrl).to(dev)
dxf = torch.zeros((21, N), dtype=torch.float64, device=dev)
for it in range(3):
    t0 = time.time()
    ctx.rollout_dev(N, T, 10, dx0.data_ptr(), dctrl.data_ptr(), 0, dxf.data_ptr())
    ctx.sync()
    dt = time.time() - t0
    print(f"bekker grid N={N} T={T}: {dt*1e3:.1f} ms -> {N*(T-1)*10/dt:.3e} vehicle-steps/s; nan={torch.isnan(dxf).any().item()} z mean {dxf[6].mean().item():.4f}")
