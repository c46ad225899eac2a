"""Quick GPU sanity + timing (development aid)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from auvsl_neural_ode_b200 import capi, synth

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
Tf = int(sys.argv[2]) if len(sys.argv) > 2 else 600
Tb = int(sys.argv[3]) if len(sys.argv) > 3 else 200
ctx = capi.Context(0, capi.TIRE_NN)
print("fp64 peak TFLOP/s:", ctx.measure_fp64_peak())
dev = torch.device("cuda:0")
for T, grad in ((Tf, False), (Tb, True)):
    x0, ctrl, gt = synth.make_batch(N, T)
    dx0, dctrl, dgt = (torch.from_numpy(a).to(dev) for a in (x0, ctrl, gt))
    dxf = torch.zeros((21, N), dtype=torch.float64, device=dev)
    dloss = torch.zeros(N, dtype=torch.float64, device=dev)
    dred = torch.zeros(418, dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    for it in range(int(sys.argv[4]) if len(sys.argv) > 4 else 3):
        t0 = time.time()
        if grad:
            ctx.rollout_grad_dev(N, T, 10, dx0.data_ptr(), dctrl.data_ptr(), dgt.data_ptr(), dloss.data_ptr(), dred.data_ptr())
        else:
            ctx.rollout_dev(N, T, 10, dx0.data_ptr(), dctrl.data_ptr(), 0, dxf.data_ptr())
        ctx.sync()
        dt = time.time() - t0
        f, b = ctx.last_kernel_ms()
        steps = N * (T - 1) * 10
        print(f"grad={grad} N={N} T={T} wall={dt*1e3:.2f} ms fwd_kernel={f:.2f} ms bwd_kernel={b:.2f} ms -> {steps/dt:.3e} vehicle-steps/s")
    if grad:
        print("loss mean", dloss.mean().item(), "n_kept", dred[417].item(), "gradnorm", dred[:416].abs().sum().item())
    else:
        print("final x mean", dxf[4].mean().item(), "nan:", torch.isnan(dxf).any().item())
