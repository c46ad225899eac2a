"""Turns an ncu capture of scripts/ncu_targets.py into profiles/r2_counters.json: per kernel family the FP64 instruction
counts, DRAM bytes and pipe utilisation, normalised PER VEHICLE-STEP, plus a hash of the kernel sources they belong to.
bench.py multiplies the per-step counts by the units of its own launches and divides by its own measured durations.

usage: python scripts/ncu_counters.py gpurun_out/prof_r2.ncu-rep gpurun_out/ncu_targets_manifest.json [out.json]"""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import source_hash  # noqa: E402

rep, manifest = sys.argv[1], json.load(open(sys.argv[2]))
out = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "profiles", "r2_counters.json")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}

TAGS = [(r"rollout_fwd_kernel<0, 0, 1", "rollout_fwd_kernel<NN,flat,STORE>"), (r"rollout_fwd_kernel<0, 0, 0", "rollout_fwd_kernel<NN,flat>"),
        (r"rollout_fwd_kernel<1, 7, 0", "rollout_fwd_kernel<Bekker,dyn>"), (r"rollout_bwd2_kernel<0>", "rollout_bwd2_kernel<flat>")]
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}


def val(r, name, scale=False):
    if name not in idx:
        return None
    try:
        v = float(r[idx[name]].replace(",", ""))
    except ValueError:
        return None
    return v * SCALE.get(units[idx[name]], 1.0) if scale else v


kernels = {}
for r in data:
    name = r[idx["Kernel Name"]]
    tag = next((t for pat, t in TAGS if re.search(re.escape(pat), name)), None)
    if tag is None or tag not in manifest:
        continue
    steps = float(manifest[tag])
    dfma = val(r, "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum")
    if dfma is None or dfma != dfma:  # an interrupted replay leaves NaNs: keep the previous (complete) launch of the family
        continue
    dmul = val(r, "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum")
    dadd = val(r, "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum")
    k = {
        "ncu_name": name, "vehicle_steps_per_launch": steps,
        "duration_ms": val(r, "gpu__time_duration.sum", True), "grid": val(r, "launch__grid_size"), "block": val(r, "launch__block_size"),
        "registers": val(r, "launch__registers_per_thread"), "dyn_smem_bytes": val(r, "launch__shared_mem_per_block_dynamic", True),
        "pipe_fp64_active_pct": val(r, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
        # the same counter over ELAPSED cycles of all SMs (idle SMs of a grid smaller than the chip count): the whole-chip figure
        "pipe_fp64_elapsed_pct": val(r, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed"),
        "issue_active_pct": val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "warps_active_pct": val(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
        "warp_inst_executed": val(r, "smsp__inst_executed.sum"), "warp_inst_pipe_fp64": val(r, "sm__inst_executed_pipe_fp64.sum"),
        "shared_bank_conflicts": val(r, "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
        "per_vehicle_step": {
            "dfma": dfma / steps if dfma is not None else None, "dmul": dmul / steps if dmul is not None else None,
            "dadd": dadd / steps if dadd is not None else None,
            "dram_read_bytes": val(r, "dram__bytes_read.sum", True) / steps, "dram_write_bytes": val(r, "dram__bytes_write.sum", True) / steps,
            "warp_inst": (val(r, "smsp__inst_executed.sum") or 0) / steps,
        },
        "stalls_per_issue": {h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]: val(r, h)
                             for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")
                             and (val(r, h) or 0) >= 0.05},
    }
    kernels[tag] = k  # the second launch of a family overwrites the first (warm)
json.dump({"source_hash": source_hash(), "capture": os.path.basename(rep), "shapes": manifest.get("shapes"), "kernels": kernels,
           "how": "ncu --set full + smsp__sass_thread_inst_executed_op_{dfma,dmul,dadd}_pred_on.sum, --clock-control none, of scripts/ncu_targets.py on one B200"},
          open(out, "w"), indent=1)
for t, k in kernels.items():
    p = k["per_vehicle_step"]
    print(f"{t}: {k['duration_ms']:.3f} ms, regs {k['registers']:.0f}, fp64 pipe {k['pipe_fp64_active_pct']:.1f} %, per step: dfma {p['dfma']}, dmul {p['dmul']}, dadd {p['dadd']}, "
          f"dram {p['dram_read_bytes'] + p['dram_write_bytes']:.0f} B")
