"""A/B timing of library variants (AVS_LIB) on four workloads: NN forward-only 4096 x 600, NN fwd+adjoint 4096 x 200 and
8192 x 200, Bekker grid 8192 x 200.   usage: ab_blocks.py tag1 tag2 ...   (tag 'main' = libavsim.so)"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, numpy as np, torch
sys.path.insert(0, %r)
from auvsl_neural_ode_b200 import capi, synth
dev = torch.device("cuda:0")
nn = capi.Context(0, capi.TIRE_NN)
def fwd(ctx, N, T, z=synth.Z_STABLE_NN, clip=False):
    x0, ctrl, _ = synth.make_batch(N, T, z_stable=z)
    if clip: ctrl = np.ascontiguousarray(np.clip(ctrl, 0.5, 8.0))
    a, b = torch.from_numpy(x0).to(dev), torch.from_numpy(ctrl).to(dev)
    xf = torch.zeros((21, N), dtype=torch.float64, device=dev)
    best = 1e9
    for _ in range(3):
        ctx.rollout_dev(N, T, 10, a.data_ptr(), b.data_ptr(), 0, xf.data_ptr()); ctx.sync()
        best = min(best, ctx.last_kernel_ms()[0])
    return best
def grad(N, T):
    x0, ctrl, gt = synth.make_batch(N, T)
    d = [torch.from_numpy(v).to(dev) for v in (x0, ctrl, gt)]
    loss = torch.zeros(N, dtype=torch.float64, device=dev); red = torch.zeros(418, dtype=torch.float64, device=dev)
    bf = bb = 1e9
    for _ in range(3):
        nn.rollout_grad_dev(N, T, 10, d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), loss.data_ptr(), red.data_ptr()); nn.sync()
        f, b = nn.last_kernel_ms(); bf, bb = min(bf, f), min(bb, b)
    return bf, bb
f600 = fwd(nn, 4096, 600)
g4 = grad(4096, 200)
g8 = grad(8192, 200)
bk = capi.Context(0, capi.TIRE_BEKKER)
g = synth.make_grid_terrain()
bk.set_terrain_grid(g["height"], g["soil"], g["x0"], g["y0"], g["dx"], g["dy"])
b200 = fwd(bk, 8192, 200, z=0.065, clip=True)
print("RESULT fwd-only(4096x600) %%.2f | fwd+store/bwd(4096x200) %%.2f / %%.2f | (8192x200) %%.2f / %%.2f | bekker grid (8192x200) %%.2f ms" %% (f600, g4[0], g4[1], g8[0], g8[1], b200))
''' % ROOT
for tag in sys.argv[1:]:
    lib = os.path.join(ROOT, "auvsl_neural_ode_b200", "libavsim.so" if tag == "main" else f"libavsim_{tag}.so")
    out = subprocess.run([sys.executable, "-c", CHILD], capture_output=True, text=True, env=dict(os.environ, AVS_LIB=lib))
    res = [l for l in out.stdout.splitlines() if l.startswith("RESULT")]
    print(f"{tag:8s}", res[0][7:] if res else ("FAILED " + out.stderr[-400:]), flush=True)
