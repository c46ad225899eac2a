#!/usr/bin/env python
"""bench.py — vehicle-steps/s (fwd+adjoint, fp64) of batched Jackal rollouts on N B200s.

  python bench.py --gpus N --steps K --warmup W             (N > 1: launched under torchrun by the driver)
  python bench.py --impl reference --gpus N --steps K --warmup W

A "step" is one Trainer step over one synthetic batch (BASELINE.json configs[2], per GPU): forward
rollout with checkpoints + reverse sweep (BPTT) + per-sample explosion filter + reduction
(+ one all-reduce of 418 doubles when N > 1) + RMSProp.  1 vehicle-step = one RK4 step (dt = 1e-3 s) of
one vehicle = 4 ODE evaluations.  `value` counts the vehicle-steps of ALL ranks divided by the
max-over-ranks device time with inputs resident in HBM; `e2e` is the same through the host-buffer C-ABI
call (pinned H2D of the batch + D2H of losses/gradient inside the timed region).
The reference arm times the reference algorithm's CPU restatement (oracle/, the reference itself cannot be
built here: CppAD / iit-rbd / Eigen absent) on the box's host cores.
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
# algorithmic work per vehicle-step of the NN-tire path, in flop-equivalents
# (add/mul = 1, fma = 2, tanh = 40, sqrt = div = 20)
FLOP_EQ_FWD = 29.5e3
FLOP_EQ_FWD_ADJ = 76.8e3


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="trajectories per GPU (configs[2]: 4096)")
    ap.add_argument("--rows", type=int, default=200, help="data rows per window (Trainer::m_train_steps)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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


def cpu_baseline_fwd_adj(T, threads, n_traj, repeats=1):
    """Oracle (reference restatement) fwd + reverse-tape gradient on `threads` host threads."""
    import numpy as np
    from auvsl_neural_ode_b200 import synth
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
    return n_traj * (T - 1) * 10 / best, best


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import pyoracle as po
    cores = po.hardware_threads() or os.cpu_count() or 1
    T = args.rows
    n_traj = 8 * cores  # a bounded sample per step (a few seconds), so the whole run ends within minutes
    for _ in range(max(args.warmup, 0) and 1):  # one short warm-up pass is enough for a CPU code path
        cpu_baseline_fwd_adj(min(T, 20), cores, cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_baseline_fwd_adj(T, cores, n_traj)
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
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference cannot be built here (CppAD, iit-rbd, Eigen3 absent); this is the oracle restatement, which flatters the CPU "
                "(plain double forward + minimal tape instead of CppAD::AD<double>)",
    }
    print(json.dumps(out))


def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from auvsl_neural_ode_b200 import capi, synth
    from auvsl_neural_ode_b200.trainer import BatchTrainer

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        # NCCL prints its version banner on stdout when the first communicator comes up; keep stdout for the ONE JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            warm = torch.zeros(1, device=dev)
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    N, T, sub = args.batch, args.rows, 10

    ctx = capi.Context(local, capi.TIRE_NN)
    ctx.set_terrain(capi.FLAT)
    ctx.set_nn_params(capi.nn_default_params())
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    fp64_peak = ctx.measure_fp64_peak()

    # this rank's shard of the synthetic batch (weak scaling: N trajectories per GPU, distinct seeds per rank)
    x0, ctrl, gt = synth.make_batch(N, T, seed=synth.SEED + rank)
    h_x0, h_ctrl, h_gt = (torch.from_numpy(a).pin_memory() for a in (x0, ctrl, gt))
    d_x0, d_ctrl, d_gt = (t.to(dev) for t in (h_x0, h_ctrl, h_gt))
    trainer = BatchTrainer(ctx, dev, explode_mode=capi.EXPLODE_DROP_SAMPLE)
    steps_per_call = N * (T - 1) * sub

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- resident-input arm (value)
    for _ in range(max(args.warmup, 3)):
        trainer.train_step(d_x0, d_ctrl, d_gt, sub)
    barrier()
    launches0 = ctx.kernel_launches
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fwd_ms, bwd_ms = [], []
    barrier()
    ev0.record(stream)
    for _ in range(args.steps):
        trainer.train_step(d_x0, d_ctrl, d_gt, sub)
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = ctx.kernel_launches - launches0
    # per-kernel durations (CUDA events recorded on the launching stream inside the C-ABI call; last step)
    f_ms, b_ms = ctx.last_kernel_ms()
    clocks = sampler.stop() if rank == 0 else None
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_max = float(t_ms.item())
    value = world * steps_per_call * args.steps / (ms_max * 1e-3)
    loss_mean = float(trainer.reduced[416].item() / max(trainer.reduced[417].item(), 1.0))

    # ---------------- end-to-end arm: host-buffer C-ABI call, pinned H2D + D2H inside the timed region
    ctx.set_nn_params(capi.nn_default_params())
    ctx.reset_optimizer()
    h_loss = torch.empty(N, dtype=torch.float64).pin_memory()
    h_red = torch.empty(capi.REDUCED_LEN, dtype=torch.float64).pin_memory()
    d_red = torch.empty(capi.REDUCED_LEN, dtype=torch.float64, device=dev)
    lib, h = capi.lib(), ctx._h

    def e2e_step():
        rc = lib.avs_rollout_grad(h, N, T, sub, h_x0.data_ptr(), h_ctrl.data_ptr(), h_gt.data_ptr(), 100.0, capi.EXPLODE_DROP_SAMPLE,
                                  h_loss.data_ptr(), h_red.data_ptr(), None)
        if rc:
            raise RuntimeError(lib.avs_last_error(h).decode())
        if world > 1:
            d_red.copy_(h_red, non_blocking=True)
            dist.all_reduce(d_red, op=dist.ReduceOp.SUM)
            ctx.rmsprop_step_dev(d_red.data_ptr())
        else:
            ctx.rmsprop_step(h_red.numpy())
        ctx.sync()

    for _ in range(max(args.warmup, 3)):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    ev0.record(stream)
    for _ in range(args.steps):
        e2e_step()
    ev1.record(stream)
    barrier()
    e2e_ms = max(ev0.elapsed_time(ev1), (time.perf_counter() - t0) * 1e3)
    t_ms = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    e2e_value = world * steps_per_call * args.steps / (float(t_ms.item()) * 1e-3)
    h2d = (21 + 2 * (T - 1) + 3 * T) * 8 * N
    d2h = (N + capi.REDUCED_LEN) * 8

    # ---------------- context numbers (not the headline): configs[1] forward-only rollouts, same batch, T = 600 rows
    fwd_only = None
    if rank == 0:
        Tf = 600
        fx0, fctrl, _ = synth.make_batch(N, Tf, seed=synth.SEED)
        dfx0, dfctrl = torch.from_numpy(fx0).to(dev), torch.from_numpy(fctrl).to(dev)
        dxf = torch.empty((21, N), dtype=torch.float64, device=dev)
        for _ in range(2):
            ctx.rollout_dev(N, Tf, sub, dfx0.data_ptr(), dfctrl.data_ptr(), 0, dxf.data_ptr())
        torch.cuda.synchronize()
        ev0.record(stream)
        for _ in range(3):
            ctx.rollout_dev(N, Tf, sub, dfx0.data_ptr(), dfctrl.data_ptr(), 0, dxf.data_ptr())
        ev1.record(stream)
        torch.cuda.synchronize()
        fms = ev0.elapsed_time(ev1) / 3
        fwd_only = {"workload": f"configs[1]: {N} rollouts x {Tf} rows, NN tire model, forward only, 1 GPU", "value": N * (Tf - 1) * sub / (fms * 1e-3),
                    "unit": UNIT, "ms_per_rollout_batch": fms, "roofline_frac_flop_equiv": FLOP_EQ_FWD * N * (Tf - 1) * sub / (fms * 1e-3) / 1e12 / fp64_peak}

    if rank == 0:
        # roofline of the dominant kernel (reverse sweep): SURVEY §8(d) flop-equivalents per vehicle-step x units / duration
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        bwd_flops = (FLOP_EQ_FWD_ADJ - FLOP_EQ_FWD) * steps_per_call
        fwd_flops = FLOP_EQ_FWD * steps_per_call
        ach_bwd = bwd_flops / (b_ms * 1e-3) / 1e12 if b_ms > 0 else None
        ach_fwd = fwd_flops / (f_ms * 1e-3) / 1e12 if f_ms > 0 else None
        roofline = {
            "bound": "fp64", "kernel": "rollout_bwd_kernel", "achieved": ach_bwd, "peak": fp64_peak, "unit": "TFLOP/s",
            "frac": (ach_bwd / fp64_peak) if ach_bwd else None,
            # ncu (profiles/r1_ncu_summary.md): dram read+write of the reverse kernel = 3.02 KB per vehicle-step
            # (algorithmic: 448 B state records + 2560 B activation records read once)
            "traffic": 3020.0 * steps_per_call, "traffic_unit": "B per launch", "hbm_frac_of_measured_peak":
                (3020.0 * steps_per_call / (b_ms * 1e-3) / 1e9 / peaks.get("hbm_gbs", 6543.4)) if b_ms > 0 else None,
            "peak_source": "DFMA-chain microbenchmark run in this process (MEASURED_PEAKS.json has no fp64 entry); "
                           f"HBM peak of record {peaks.get('hbm_gbs')} GB/s is not the bound: ~6 KB of HBM traffic per vehicle-step (fwd writes + bwd reads) vs ~77 kflop-equivalents",
            "kernel_ms": {"rollout_fwd_kernel": f_ms, "rollout_bwd_kernel": b_ms},
            "fwd_kernel": {"achieved": ach_fwd, "frac": (ach_fwd / fp64_peak) if ach_fwd else None},
            # hardware-side view (ncu, profiles/README.md): share of cycles the FP64 pipe is busy
            "ncu_fp64_pipe_active_pct": {"rollout_fwd_kernel": 56.2, "rollout_bwd_kernel": 52.6, "source": "profiles/r1_ncu_bench_config.txt (not measured in this run)"},
            "flop_equiv_per_vehicle_step": {"fwd": FLOP_EQ_FWD, "fwd_adjoint": FLOP_EQ_FWD_ADJ},
            "whole_step_frac": FLOP_EQ_FWD_ADJ * value / world / 1e12 / fp64_peak,
        }
        cpu = None
        if not args.no_cpu_baseline:
            from oracle import pyoracle as po
            cores = po.hardware_threads() or os.cpu_count() or 1
            n_traj = 32 * cores  # ~10 s of CPU work on all host threads
            v_all, dt_all = cpu_baseline_fwd_adj(T, cores, n_traj)
            v_one, dt_one = cpu_baseline_fwd_adj(T, 1, 2)
            cpu = {"value": v_all, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"{n_traj} trajectories x {T} rows, fwd + reverse-tape gradient, {cores} threads ({dt_all:.1f} s); "
                             f"single thread: {v_one:.1f} {UNIT} on 2 trajectories ({dt_one:.1f} s)",
                   "single_thread_value": v_one}
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": f"configs[2] per GPU: {N} rollouts x {T} rows, NN tire model (load_model weights), flat terrain, forward + adjoint "
                                   "w.r.t. tire-NN weights + explosion filter + (all-reduce) + RMSProp",
                       "trajectories_per_gpu": N, "rows": T, "substeps": sub, "vehicle_steps_per_step": world * steps_per_call,
                       "l2": "no explicit flush: every step streams 24.5 GB of state + activation records (>> 126 MB L2) out to HBM in the "
                             "forward sweep and back in the reverse sweep",
                       "parallelism": f"dp{world} (shard by trajectory, one all-reduce of 418 fp64 per step)"},
            "clocks": clocks, "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu, "loss_mean": loss_mean, "forward_only": fwd_only,
        }
        print(json.dumps(out))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
