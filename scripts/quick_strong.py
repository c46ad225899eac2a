"""configs[4] quick timing on one GPU: 65536 windows x 200 rows per training step (chunked)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from auvsl_neural_ode_b200 import capi, synth
N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
T = 200
dev = torch.device("cuda:0")
ctx = capi.Context(0, capi.TIRE_NN)
x0, ctrl, gt = synth.make_batch(N, T)
d = [torch.from_numpy(a).to(dev) for a in (x0, ctrl, gt)]
loss = torch.zeros(N, dtype=torch.float64, device=dev); red = torch.zeros(418, dtype=torch.float64, device=dev)
for it in range(3):
    t0 = time.time()
    ctx.rollout_grad_dev(N, T, 10, d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), loss.data_ptr(), red.data_ptr(), 0, 100.0, 1)
    ctx.sync()
    dt = time.time() - t0
    print(f"N={N} T={T}: {dt*1e3:.1f} ms -> {N*(T-1)*10/dt:.3e} vehicle-steps/s, n_kept {red[417].item()}")
