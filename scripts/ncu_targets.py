"""Launches every kernel family of bench.py at its bench configuration, twice each, for ONE `ncu --set full` capture:
    rollout_fwd_kernel<NN,flat,STORE> + rollout_bwd2_kernel<flat>   configs[2]: 4096 x 200 rows, forward + adjoint
    rollout_fwd_kernel<NN,flat>                                     configs[1]: 4096 x 600 rows, forward only
    rollout_fwd_kernel<Bekker,dyn>                                  configs[3]: 8192 x 600 rows, grid height + soil
Writes gpurun_out/ncu_targets_manifest.json (vehicle-steps per launch of each family) for scripts/ncu_counters.py.
Run it plain first, then under
    ncu --set full --metrics smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,\
smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,sm__inst_executed_pipe_fp64.sum --clock-control none --import-source on \
        -k regex:rollout -c 4 -o gpurun_out/prof_r2 python scripts/ncu_targets.py once"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

from auvsl_neural_ode_b200 import capi, synth

small = len(sys.argv) > 1 and sys.argv[1] == "small"  # quick functional check
once = "once" in sys.argv[1:]                           # one launch per family (an ncu capture replays each ~40 times)
REPS = 1 if once else 2
# the Bekker family is captured on 200 rows instead of 600 (same per-vehicle-step counts, a third of the replay time)
N, T, Tf, Nb, Tb = (4096, 200, 600, 8192, 200 if once else 600) if not small else (256, 20, 30, 256, 30)
nn = capi.Context(0, capi.TIRE_NN)
x0, ctrl, gt = synth.make_batch(N, T)
for _ in range(REPS):
    loss, red, _ = nn.rollout_grad(x0, ctrl, gt, T, explode_mode=1, per_sample=False)
fx0, fctrl, _ = synth.make_batch(N, Tf, seed=synth.SEED + 100)
for _ in range(REPS):
    _, xf = nn.rollout(fx0, fctrl, Tf, want_list=False)
bk = capi.Context(0, capi.TIRE_BEKKER)
g = synth.make_grid_terrain()
bk.set_terrain_grid(g["height"], g["soil"], g["x0"], g["y0"], g["dx"], g["dy"])
bx0, bctrl, _ = synth.make_batch(Nb, Tb, seed=synth.SEED + 200, z_stable=0.065)
bctrl = np.ascontiguousarray(np.clip(bctrl, 0.5, 8.0))
for _ in range(REPS):
    _, bxf = bk.rollout(bx0, bctrl, Tb, want_list=False)
ok = bool(np.isfinite(loss).all() and np.isfinite(xf).all() and np.isfinite(bxf).all())
manifest = {"rollout_fwd_kernel<NN,flat,STORE>": N * (T - 1) * 10, "rollout_bwd2_kernel<flat>": N * (T - 1) * 10,
            "rollout_fwd_kernel<NN,flat>": N * (Tf - 1) * 10, "rollout_fwd_kernel<Bekker,dyn>": Nb * (Tb - 1) * 10,
            "shapes": {"N": N, "T": T, "T_forward": Tf, "N_bekker": Nb, "T_bekker": Tb}}
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(manifest, open(os.path.join(ROOT, "gpurun_out", "ncu_targets_manifest.json"), "w"), indent=1)
print("ncu targets ok" if ok else "ncu targets: NON-FINITE RESULT", red[417], float(loss.mean()))
sys.exit(0 if ok else 1)
