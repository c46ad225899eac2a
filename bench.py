#!/usr/bin/env python
"""bench.py — vehicle-steps/s (fwd+adjoint, fp64) of batched Jackal rollouts on N B200s.

  python bench.py --gpus N --steps K --warmup W             (N > 1: launched under torchrun by the driver)
  python bench.py --impl reference --gpus N --steps K --warmup W

A "step" is one Trainer step over one synthetic batch (BASELINE.json configs[2], per GPU): forward
rollout with checkpoints + reverse sweep (BPTT) + per-sample explosion filter + reduction
(+ one all-reduce of 418 doubles when N > 1) + RMSProp.  1 vehicle-step = one RK4 step (dt = 1e-3 s) of
one vehicle = 4 ODE evaluations.  `value` counts the vehicle-steps of ALL ranks divided by the
max-over-ranks device time with inputs resident in HBM; `e2e` is the same through the host-buffer C-ABI
call (pinned H2D of the batch + D2H of losses/gradient inside the timed region).  The same JSON line carries three more
blocks, each with its own kernel time, clocks and roofline: `configs1_forward_only` (4096 x 600 forward), `configs3_bekker_grid`
(Bekker on height / soil grids, 8192 per GPU x 600) and `strong_scaling` (configs[4]: 65536 windows per step split over the ranks).
`roofline` comes from hardware counters: FP64 instruction counts per vehicle-step of the committed ncu capture of THIS build
(profiles/r2_counters.json, tied by a hash of the kernel sources) over kernel durations measured live.
The reference arm (and `cpu_baseline`) time the reference's OWN Trainer::trainTrajectory on the box's host cores: oracle/_ref is
built from the reference's first-party sources in place, behind stand-in headers for its absent third-party libraries (Eigen,
CppAD, iit-rbd: oracle/refshim/); without that prebuilt library they fall back to the oracle restatement (kind "port").
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "vehicle-steps/sec (fwd+adjoint, fp64) batched rollouts"
UNIT = "vehicle-steps/s"
COUNTERS = os.path.join(ROOT, "profiles", "r2_counters.json")  # written by scripts/ncu_counters.py from an ncu capture of this build
# algorithmic HBM bytes per vehicle-step (controls 1.6 + x_list 18.4 + macro-step checkpoints 2 x 16.8)
ALGORITHMIC_HBM_B_PER_STEP = 55.0
# flop-equivalent weights of SURVEY.md §8(d) (add/mul = 1, fma = 2, tanh = 40, sqrt = div = 20), applied to the op counts of the
# reference formulation measured by the oracle's counting scalar (pyoracle.count_ops) in the cpu_baseline leg
FLOP_EQ_W = {"addsub": 1, "mul": 1, "div": 20, "tanh": 40, "sqrt": 20, "trig": 40, "exp": 40, "pow": 80, "other": 1}


def source_hash():
    """sha256 over the kernel sources of libavsim.so: ties the committed ncu counters to the build that is running."""
    import glob
    import hashlib
    h = hashlib.sha256()
    d = os.path.join(ROOT, "auvsl_neural_ode_b200", "csrc")
    for f in sorted(glob.glob(os.path.join(d, "*.cu")) + glob.glob(os.path.join(d, "*.cuh")) + glob.glob(os.path.join(d, "*.h"))):
        h.update(os.path.basename(f).encode())
        h.update(open(f, "rb").read())
    return h.hexdigest()[:16]


def load_counters():
    try:
        c = json.load(open(COUNTERS))
    except Exception:
        return None
    c["matches_build"] = (c.get("source_hash") == source_hash())
    return c


N_SCHEDULERS = 148 * 4  # SM sub-partitions of a B200, one issue slot per cycle each
SM_CLOCK_MHZ = 1965.0    # maximum SM clock of the pool's B200s (the clocks block of the line reports what the run saw)


def roofline_block(counters, kernel, units, kernel_ms, fp64_peak_tflops, hbm_peak_gbs):
    """Roofline of one kernel launch from HARDWARE counters.  The FP64 instruction counts per vehicle-step come from the
    committed ncu capture of this build (profiles/r2_counters.json: smsp__sass_thread_inst_executed_op_{dfma,dmul,dadd}_pred_on,
    deterministic per vehicle-step), the duration is measured live in this run with CUDA events on the launching stream.
      achieved  = FP64 thread-instructions per second x 2 flop  (every FP64 instruction holds a pipe slot that could be a DFMA)
      frac      = achieved / DFMA peak measured in this process  (tracks ncu's sm__pipe_fp64_cycles_active)
      executed_flops_* = the same with DMUL / DADD counted as 1 flop (the "true flops" view)
    """
    blk = {"bound": "fp64", "kernel": kernel, "kernel_ms": kernel_ms, "peak": fp64_peak_tflops, "unit": "TFLOP/s",
           "peak_source": "DFMA-chain microbenchmark in this process (MEASURED_PEAKS.json has no fp64 entry; datasheet 37.2 at 1965 MHz)",
           "achieved": None, "frac": None, "traffic": None}
    k = (counters or {}).get("kernels", {}).get(kernel)
    if not k or not kernel_ms or kernel_ms <= 0:
        blk["note"] = "no committed counters for this kernel (profiles/r2_counters.json)"
        return blk
    per = k["per_vehicle_step"]
    sec = kernel_ms * 1e-3
    fp64_inst = (per["dfma"] + per["dmul"] + per["dadd"]) * units
    flops = (2 * per["dfma"] + per["dmul"] + per["dadd"]) * units
    traffic = (per["dram_read_bytes"] + per["dram_write_bytes"]) * units
    blk.update({
        "achieved": 2.0 * fp64_inst / sec / 1e12, "frac": 2.0 * fp64_inst / sec / 1e12 / fp64_peak_tflops,
        "achieved_definition": "FP64 thread-instructions/s x 2 flop (issue slots of the FP64 pipe)",
        "executed_flops_tflops": flops / sec / 1e12, "executed_flops_frac": flops / sec / 1e12 / fp64_peak_tflops,
        "fp64_thread_inst_per_vehicle_step": {"dfma": per["dfma"], "dmul": per["dmul"], "dadd": per["dadd"]},
        # ncu's own reading of the same pipe at capture time: share of ELAPSED cycles over all SMs (comparable with frac, which is
        # relative to the measured DFMA peak = 94 % of the datasheet rate) and share of cycles of the SMs that were active
        "ncu_pipe_fp64_pct_of_elapsed_at_capture": k.get("pipe_fp64_elapsed_pct"),
        "ncu_pipe_fp64_pct_of_active_at_capture": k.get("pipe_fp64_active_pct"),
        "traffic": traffic, "traffic_unit": "B per launch (dram__bytes_read + dram__bytes_write, ncu)",
        "traffic_over_algorithmic": (per["dram_read_bytes"] + per["dram_write_bytes"]) / ALGORITHMIC_HBM_B_PER_STEP,
        "hbm_gbs": traffic / sec / 1e9, "hbm_frac_of_measured_peak": traffic / sec / 1e9 / hbm_peak_gbs if hbm_peak_gbs else None,
        "counters": {"file": os.path.relpath(COUNTERS, ROOT), "source_hash": counters.get("source_hash"), "matches_build": counters.get("matches_build"),
                     "captured_ms": k.get("duration_ms"), "registers": k.get("registers"), "grid": k.get("grid"), "block": k.get("block")},
    })
    # Issue-slot view: on this chip a warp's FP64 instruction occupies its scheduler for two issue cycles, any
    # other instruction for one, and both rollout kernels are bound by exactly that (measured round 2: 216 extra FP64 instructions
    # per stage cost the forward kernel their issue time wherever they were placed).  slots = all warp instructions + the FP64 ones
    # once more; frac_of_chip = slots / (duration x SM clock x 592 schedulers).
    if k.get("warp_inst_executed") and k.get("warp_inst_pipe_fp64") and k.get("vehicle_steps_per_launch"):
        per_step = (k["warp_inst_executed"] + k["warp_inst_pipe_fp64"]) / k["vehicle_steps_per_launch"]
        blk["issue_slots"] = {"warp_inst_per_vehicle_step": k["warp_inst_executed"] / k["vehicle_steps_per_launch"],
                              "fp64_warp_inst_per_vehicle_step": k["warp_inst_pipe_fp64"] / k["vehicle_steps_per_launch"],
                              "slots_per_vehicle_step": per_step,
                              "frac_of_chip": per_step * units / (sec * SM_CLOCK_MHZ * 1e6 * N_SCHEDULERS),
                              "definition": "(warp instructions + FP64 warp instructions) / (duration x 1965 MHz x 592 schedulers)"}
    return blk


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="trajectories per GPU (configs[2]: 4096)")
    ap.add_argument("--rows", type=int, default=200, help="data rows per window (Trainer::m_train_steps)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the configs[1]/[3]/[4] blocks (headline only)")
    ap.add_argument("--strong-windows", type=int, default=65536, help="configs[4]: total windows per training step, split over the ranks")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.t.join(timeout=2)
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm), "reasons": sorted(reasons)}


def _windows_for_reference(n_traj, T, seed):
    """synthetic batch -> GroundTruthDataRow windows [n][T][10] (time, vl, vr, x, y, z, yaw, wz, vx, vy) as Trainer::loadDataFile
    yields them; Trainer's initializeState on row 0 reproduces synth.make_batch's x0 exactly"""
    import numpy as np
    from auvsl_neural_ode_b200 import synth
    x0, ctrl, gt = synth.make_batch(n_traj, T, seed=seed)
    vl = np.concatenate([ctrl[:, 0, :], ctrl[-1:, 0, :]])
    vr = np.concatenate([ctrl[:, 1, :], ctrl[-1:, 1, :]])
    rows = np.zeros((n_traj, T, 10))
    for n in range(n_traj):
        v = synth.B @ np.stack([vl[:, n], vr[:, n]])
        rows[n, :, 0] = np.arange(T) * 0.01
        rows[n, :, 1], rows[n, :, 2] = vl[:, n], vr[:, n]
        rows[n, :, 3], rows[n, :, 4], rows[n, :, 5], rows[n, :, 6] = gt[:, 1, n], gt[:, 2, n], synth.Z_STABLE_NN, gt[:, 0, n]
        rows[n, :, 7], rows[n, :, 8], rows[n, :, 9] = v[2], v[0], v[1]
    return rows


def cpu_baseline_fwd_adj(T, threads, n_traj, repeats=1):
    """The reference's CPU path on `threads` host threads, one trajectory per thread like Trainer::Worker.
    kind "reference": oracle/_ref — the reference's OWN Trainer::trainTrajectory (CppAD-style tape + Reverse(1)), compiled in
    place from its sources behind the stand-in Eigen / CppAD / iit-rbd headers of oracle/refshim/;
    kind "port": the oracle restatement, when the prebuilt reference library is not there.
    Returns (vehicle-steps/s, seconds, kind)."""
    import numpy as np
    from auvsl_neural_ode_b200 import synth
    from oracle import pyref
    if pyref.available():
        from concurrent.futures import ThreadPoolExecutor
        rows = _windows_for_reference(n_traj, T, synth.SEED + 17)
        P = pyref.default_params(0)
        pyref.train_trajectory(P, 0, rows[0][:3])  # load + first-call costs outside the timed region
        best = None
        for _ in range(repeats):
            t0 = time.perf_counter()
            with ThreadPoolExecutor(max_workers=threads) as ex:  # ctypes releases the GIL; the tape is thread_local
                list(ex.map(lambda w: pyref.train_trajectory(P, 0, w)[0], rows))
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        return n_traj * (T - 1) * 10 / best, best, "reference"
    from oracle import pyoracle as po
    x0, ctrl, gt = synth.make_batch(n_traj, T, seed=synth.SEED + 17)
    p = po.nn_default_params()
    flat = po.terrain(po.FLAT)
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        po.rollout_grad_batch(p, flat, x0, ctrl, gt, T, nthreads=threads, per_sample=False)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return n_traj * (T - 1) * 10 / best, best, "port"


CPU_KIND_NOTE = {
    "reference": "the reference's own first-party sources (Trainer::trainTrajectory -> VehicleSystem::integrate -> HybridDynamics::RK4/ODE -> "
                 "TireNetwork / generated ForwardDynamics), compiled in place (-O2) against stand-in headers for its absent third-party libraries "
                 "(Eigen, CppAD, iit-rbd: oracle/refshim/); its CMake build cannot run here",
    "port": "oracle restatement (the prebuilt oracle/_ref library is absent); flatters the CPU: plain double forward + a minimal tape",
}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    T = args.rows
    n_traj = 2 * cores  # a bounded sample per step (a few seconds), so the whole run ends within minutes
    if args.warmup > 0:  # one short warm-up pass is enough for a CPU code path
        cpu_baseline_fwd_adj(min(T, 20), cores, cores)
    t0 = time.perf_counter()
    kind = "port"
    for _ in range(args.steps):
        _, _, kind = cpu_baseline_fwd_adj(T, cores, n_traj)
    dt = time.perf_counter() - t0
    units = args.steps * n_traj * (T - 1) * 10
    value = units / dt
    sample = f"{n_traj} trajectories x {T} rows per step (same synthetic distribution), forward + reverse-tape gradient, one trajectory per thread"
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"configs[2] sample: NN tire model + adjoint, T={T}, {n_traj} trajectories/step on host CPU", "rows": T,
                   "substeps": 10, "terrain": "flat"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": CPU_KIND_NOTE[kind],
    }
    print(json.dumps(out))


def cpu_reference_op_counts():
    """Op counts of ONE RK4 step as the reference formulation executes it (oracle counting scalar; cpu_baseline leg only)."""
    import numpy as np
    from oracle import pyoracle as po
    X = np.zeros(21)
    X[3], X[6], X[14] = 1.0, 0.0584, 0.3
    u = np.array([4.0, 3.0])
    nn = po.count_ops(po.TIRE_NN, po.nn_default_params(), X, u, 1)
    bk = po.count_ops(po.TIRE_BEKKER, po.bekker_default_params(), X, u, 1)
    eq = lambda c: float(sum(FLOP_EQ_W[k] * v for k, v in c.items()))
    return {"nn_rk4_step": nn, "bekker_rk4_step": bk, "nn_flop_equiv": eq(nn), "bekker_flop_equiv": eq(bk), "weights": FLOP_EQ_W}


def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from auvsl_neural_ode_b200 import capi, synth
    from auvsl_neural_ode_b200.trainer import BatchTrainer, shard_range

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    stdout_saved = None
    if world > 1:
        # NCCL prints its version banner on stdout when a communicator comes up; keep stdout for the ONE JSON line
        sys.stdout.flush()
        stdout_saved = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)
        warm = torch.zeros(1, device=dev)
        dist.all_reduce(warm)
        torch.cuda.synchronize()
    N, T, sub = args.batch, args.rows, 10

    ctx = capi.Context(local, capi.TIRE_NN)
    ctx.set_terrain(capi.FLAT)
    ctx.set_nn_params(capi.nn_default_params())
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    fp64_peak = ctx.measure_fp64_peak()
    counters = load_counters()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    hbm_src = "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"

    # this rank's shard of the synthetic batch (weak scaling: N trajectories per GPU, distinct seeds per rank)
    x0, ctrl, gt = synth.make_batch(N, T, seed=synth.SEED + rank)
    h_x0, h_ctrl, h_gt = (torch.from_numpy(a).pin_memory() for a in (x0, ctrl, gt))
    d_x0, d_ctrl, d_gt = (t.to(dev) for t in (h_x0, h_ctrl, h_gt))
    trainer = BatchTrainer(ctx, dev, explode_mode=capi.EXPLODE_DROP_SAMPLE)
    trainer.init_native_comm()  # world > 1: the all-reduce is avs_allreduce_dev (ncclAllReduce on the context's stream, C ABI)
    if stdout_saved is not None:
        import ctypes
        sys.stdout.flush()
        ctypes.CDLL(None).fflush(None)  # NCCL writes its banner through C stdio: push it out while fd 1 still points at stderr
        os.dup2(stdout_saved, 1)
        os.close(stdout_saved)
    steps_per_call = N * (T - 1) * sub

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(fn, steps, warmup):
        """W untimed + K timed calls of fn, bracketed by barrier + synchronize; returns (max-over-ranks ms, clocks on rank 0)."""
        for _ in range(warmup):
            fn()
        barrier()
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        barrier()
        t0 = time.perf_counter()
        ev0.record(stream)
        for _ in range(steps):
            fn()
        ev1.record(stream)
        barrier()
        wall_ms = (time.perf_counter() - t0) * 1e3
        clocks = sampler.stop() if rank == 0 else None
        return max_over_ranks(ev0.elapsed_time(ev1)), max_over_ranks(wall_ms), clocks

    # ---------------- headline, resident-input arm (value): configs[2] per GPU
    launches0 = ctx.kernel_launches
    W = max(args.warmup, 3)
    ms_max, _, clocks = timed(lambda: trainer.train_step(d_x0, d_ctrl, d_gt, sub), args.steps, W)
    launches = (ctx.kernel_launches - launches0) * args.steps // (args.steps + W)
    f_ms, b_ms = ctx.last_kernel_ms()  # CUDA events around each rollout kernel inside the C-ABI call (last step)
    value = world * steps_per_call * args.steps / (ms_max * 1e-3)
    loss_mean = float(trainer.reduced[416].item() / max(trainer.reduced[417].item(), 1.0))

    # ---------------- end-to-end arm: host-buffer C-ABI call, pinned H2D + D2H inside the timed region
    ctx.set_nn_params(capi.nn_default_params())
    ctx.reset_optimizer()
    h_loss = torch.empty(N, dtype=torch.float64).pin_memory()
    h_red = torch.empty(capi.REDUCED_LEN, dtype=torch.float64).pin_memory()
    lib, h = capi.lib(), ctx._h

    def e2e_step():
        rc = lib.avs_rollout_grad(h, N, T, sub, h_x0.data_ptr(), h_ctrl.data_ptr(), h_gt.data_ptr(), 100.0, capi.EXPLODE_DROP_SAMPLE,
                                  h_loss.data_ptr(), h_red.data_ptr(), None)
        if rc:
            raise RuntimeError(lib.avs_last_error(h).decode())
        if world > 1:
            ctx.allreduce_dev(0)     # the context's own reduced buffer, on the device: no host hop
        ctx.rmsprop_step_dev(0)      # RMSProp on the (all-reduced) device buffer
        ctx.sync()

    e2e_ev_ms, e2e_wall_ms, _ = timed(e2e_step, args.steps, W)
    e2e_value = world * steps_per_call * args.steps / (max(e2e_ev_ms, e2e_wall_ms) * 1e-3)
    h2d = (21 + 2 * (T - 1) + 3 * T) * 8 * N
    d2h = (N + capi.REDUCED_LEN) * 8

    extra = {}
    if not args.no_extra:
        # ---------------- configs[1]: forward-only rollouts, N x 600 rows, NN tire model (every rank its own batch: weak)
        Tf = 600
        fx0, fctrl, _ = synth.make_batch(N, Tf, seed=synth.SEED + 100 + rank)
        dfx0, dfctrl = torch.from_numpy(fx0).to(dev), torch.from_numpy(fctrl).to(dev)
        dxf = torch.empty((21, N), dtype=torch.float64, device=dev)
        fsteps = N * (Tf - 1) * sub
        fms, _, fclk = timed(lambda: ctx.rollout_dev(N, Tf, sub, dfx0.data_ptr(), dfctrl.data_ptr(), 0, dxf.data_ptr()), 3, 3)
        kf_ms, _ = ctx.last_kernel_ms()
        if rank == 0:
            extra["configs1_forward_only"] = {
                "workload": f"configs[1] per GPU: {N} rollouts x {Tf} rows, NN tire model (load_model weights), flat terrain, forward only",
                "value": world * fsteps * 3 / (fms * 1e-3), "unit": UNIT, "n_gpus": world, "scaling": "weak", "ms_per_rollout_batch": fms / 3,
                "kernel_ms": {"rollout_fwd_kernel<NN,flat>": kf_ms}, "clocks": fclk,
                "roofline": roofline_block(counters, "rollout_fwd_kernel<NN,flat>", fsteps, kf_ms, fp64_peak, hbm_peak)}
        del dfx0, dfctrl, dxf

        # ---------------- configs[3]: Bekker model on randomised height / soil grids, 65536 rollouts over 8 GPUs = 8192 per GPU x 600 rows
        Nb = 8192
        bk = capi.Context(local, capi.TIRE_BEKKER)
        bk.set_stream(stream.cuda_stream)
        g = synth.make_grid_terrain()
        bk.set_terrain_grid(g["height"], g["soil"], g["x0"], g["y0"], g["dx"], g["dy"])
        bx0, bctrl, _ = synth.make_batch(Nb, Tf, seed=synth.SEED + 200 + rank, z_stable=0.065)
        bctrl = np.ascontiguousarray(np.clip(bctrl, 0.5, 8.0))
        dbx0, dbctrl = torch.from_numpy(bx0).to(dev), torch.from_numpy(bctrl).to(dev)
        dbf = torch.empty((21, Nb), dtype=torch.float64, device=dev)
        bsteps = Nb * (Tf - 1) * sub
        bms, _, bclk = timed(lambda: bk.rollout_dev(Nb, Tf, sub, dbx0.data_ptr(), dbctrl.data_ptr(), 0, dbf.data_ptr()), 2, 1)
        kb_ms, _ = bk.last_kernel_ms()
        finite = bool(torch.isfinite(dbf).all().item())
        # the same workload with the 10-angle contact search (HybridDynamics::get_tire_cpts_sinkages_fancy) switched on
        bk.set_contact_search(True)
        cms, _, _ = timed(lambda: bk.rollout_dev(Nb, Tf, sub, dbx0.data_ptr(), dbctrl.data_ptr(), 0, dbf.data_ptr()), 1, 1)
        bk.set_contact_search(False)
        if rank == 0:
            extra["configs3_bekker_grid"] = {
                "workload": f"configs[3] per GPU: {Nb} rollouts x {Tf} rows, Bekker tire model, 512^2 randomised height + per-cell soil grids (65536 rollouts over 8 GPUs)",
                "value": world * bsteps * 2 / (bms * 1e-3), "unit": UNIT, "n_gpus": world, "scaling": "weak", "ms_per_rollout_batch": bms / 2,
                "kernel_ms": {"rollout_fwd_kernel<Bekker,dyn>": kb_ms}, "clocks": bclk, "all_finite": finite,
                "with_contact_search": {"value": world * bsteps / (cms * 1e-3), "unit": UNIT, "ms_per_rollout_batch": cms,
                                        "what": "same batch, sinkage from the 10-angle contact search (40 height lookups per ODE evaluation instead of 4)"},
                "roofline": roofline_block(counters, "rollout_fwd_kernel<Bekker,dyn>", bsteps, kb_ms, fp64_peak, hbm_peak)}
        bk.close()
        del dbx0, dbctrl, dbf

        # ---------------- configs[4]: end-to-end training step over 65536 windows, split over the ranks (STRONG scaling)
        Ns_total = args.strong_windows
        sb, se = shard_range(Ns_total, rank, world)
        Ns = se - sb
        sx0, sctrl, sgt = synth.make_batch(Ns, T, seed=synth.SEED + 300 + rank)
        dsx0, dsctrl, dsgt = (torch.from_numpy(a).to(dev) for a in (sx0, sctrl, sgt))
        ctx.set_nn_params(capi.nn_default_params())
        ctx.reset_optimizer()
        tr2 = BatchTrainer(ctx, dev, explode_mode=capi.EXPLODE_DROP_SAMPLE)
        tr2.native_comm = trainer.native_comm
        sms, _, sclk = timed(lambda: tr2.train_step(dsx0, dsctrl, dsgt, sub), 3, 1)
        n_kept = float(tr2.reduced[417].item())
        if rank == 0:
            extra["strong_scaling"] = {
                "workload": f"configs[4]: {Ns_total} windows x {T} rows per training step (forward + BPTT + all-reduce + RMSProp), split over {world} GPU(s)",
                "total_windows": Ns_total, "windows_per_gpu": Ns, "n_gpus": world, "scaling": "strong", "value": Ns_total * (T - 1) * sub * 3 / (sms * 1e-3),
                "unit": UNIT, "ms_per_step": sms / 3, "n_kept": n_kept, "clocks": sclk,
                "note": "the record streams of one GPU's share (6 KB per vehicle-step) are processed in 256-vehicle-aligned chunks when they exceed 60 % of free HBM"}
        del dsx0, dsctrl, dsgt

    if rank == 0:
        roofline = roofline_block(counters, "rollout_bwd2_kernel<flat>", steps_per_call, b_ms, fp64_peak, hbm_peak)
        roofline["hbm_peak"] = {"gbs": hbm_peak, "source": hbm_src}
        roofline["fwd_kernel"] = roofline_block(counters, "rollout_fwd_kernel<NN,flat,STORE>", steps_per_call, f_ms, fp64_peak, hbm_peak)
        roofline["kernel_share_of_step"] = {"rollout_fwd_kernel<NN,flat,STORE>": f_ms / (ms_max / args.steps), "rollout_bwd2_kernel<flat>": b_ms / (ms_max / args.steps)}
        cpu = None
        if not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            n_traj = 8 * cores  # ~10-20 s of CPU work on all host threads
            v_all, dt_all, kind = cpu_baseline_fwd_adj(T, cores, n_traj)
            v_one, dt_one, _ = cpu_baseline_fwd_adj(T, 1, 2)
            ops = cpu_reference_op_counts()
            cpu = {"value": v_all, "unit": UNIT, "cores": cores, "kind": kind,
                   "sample": f"{n_traj} trajectories x {T} rows, fwd + reverse-tape gradient, {cores} threads ({dt_all:.1f} s); "
                             f"single thread: {v_one:.1f} {UNIT} on 2 trajectories ({dt_one:.1f} s)",
                   "single_thread_value": v_one, "what": CPU_KIND_NOTE[kind], "reference_formulation_ops": ops}
            # secondary, clearly named: useful work measured in the REFERENCE formulation's operations (what CppAD would execute per
            # RK4 step: dense 6x6 products, LLT every call, the z-network's 256 tanh ...), not what the kernels execute
            fe = ops["nn_flop_equiv"]
            roofline["reference_formulation_work"] = {
                "flop_equiv_per_vehicle_step_forward": fe,
                "definition": "NOT a roofline fraction: the rate at which the run retires the REFERENCE formulation's operations (oracle count_ops of "
                              "one RK4 step: dense 6x6 products, LLT per call, 512 tanh ... with SURVEY 8(d) weights; fwd+adjoint taken as 3 x forward), "
                              "in TFLOP/s-equivalents.  It exceeds the FP64 peak because the kernels fold most of that work away (DESIGN.md section 2)",
                "whole_step_equiv_tflops": 3.0 * fe * value / world / 1e12,
                "forward_only_equiv_tflops": (fe * extra["configs1_forward_only"]["value"] / world / 1e12) if "configs1_forward_only" in extra else None,
                "fp64_peak_tflops": fp64_peak}
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": W,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": f"configs[2] per GPU: {N} rollouts x {T} rows, NN tire model (load_model weights), flat terrain, forward + adjoint "
                                   "w.r.t. tire-NN weights + explosion filter + (all-reduce) + RMSProp",
                       "trajectories_per_gpu": N, "rows": T, "substeps": sub, "vehicle_steps_per_step": world * steps_per_call,
                       "l2": "no explicit flush: every step streams 22.4 GB of state + activation records (>> 126 MB L2) out to HBM in the "
                             "forward sweep and back in the reverse sweep",
                       "parallelism": f"dp{world} (shard by trajectory, one ncclAllReduce of 418 fp64 per step behind the C ABI: avs_allreduce_dev)"},
            "clocks": clocks, "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu, "loss_mean": loss_mean,
        }
        out.update(extra)
        print(json.dumps(out))
        sys.stdout.flush()
    if world > 1:
        ctx.comm_destroy()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
