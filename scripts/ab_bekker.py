"""A/B timing of library variants on the Bekker grid-soil workload (configs[3] shape, shortened): each variant runs in its own
process through AVS_LIB and prints the best of three launches plus a checksum of the final states.
usage: ab_bekker.py tag1 tag2 ...   (tag 'main' = libavsim.so)"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import sys, os, time
sys.path.insert(0, %r)
import numpy as np, torch
from auvsl_neural_ode_b200 import capi, synth
N, T = 8192, 100
ctx = capi.Context(0, capi.TIRE_BEKKER)
g = synth.make_grid_terrain()
ctx.set_terrain_grid(g["height"], g["soil"], g["x0"], g["y0"], g["dx"], g["dy"])
x0, ctrl, gt = synth.make_batch(N, T, z_stable=0.065)
ctrl = np.clip(ctrl, 0.5, 8.0)
dev = torch.device("cuda:0")
dx0, dctrl = torch.from_numpy(x0).to(dev), torch.from_numpy(ctrl).to(dev)
dxf = torch.zeros((21, N), dtype=torch.float64, device=dev)
best = 1e9
for it in range(4):
    torch.cuda.synchronize(); t0 = time.time()
    ctx.rollout_dev(N, T, 10, dx0.data_ptr(), dctrl.data_ptr(), 0, dxf.data_ptr())
    ctx.sync(); best = min(best, time.time() - t0)
xf = dxf.cpu().numpy()
print(f"RESULT {best*1e3:.2f} ms  {N*(T-1)*10/best:.4e} vehicle-steps/s  finite={np.isfinite(xf).all()}  sum|x|={np.abs(xf[:13]).sum():.12e}")
np.save(os.environ.get("AVS_AB_OUT", "/tmp/ab_xf.npy"), xf)
''' % ROOT
ref = None
for tag in sys.argv[1:]:
    lib = os.path.join(ROOT, "auvsl_neural_ode_b200", "libavsim.so" if tag == "main" else f"libavsim_{tag}.so")
    out_npy = f"/tmp/ab_xf_{tag}.npy"
    env = dict(os.environ, AVS_LIB=lib, AVS_AB_OUT=out_npy)
    r = subprocess.run([sys.executable, "-c", CODE], capture_output=True, text=True, env=env)
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT")]
    extra = ""
    try:
        import numpy as np
        xf = np.load(out_npy)
        if ref is None: ref = xf
        else:
            d = np.abs(xf[:13] - ref[:13]) / (1e-3 + np.abs(ref[:13]))
            extra = f"  max rel diff vs first {d.max():.2e}  (p99.9 {np.quantile(d, 0.999):.2e})"
    except Exception as e:
        extra = f"  ({e})"
    print(f"{tag:8s} {line[0] if line else r.stderr[-400:]}{extra}", flush=True)
