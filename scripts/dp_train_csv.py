"""Data-parallel trainThreads steps over Train3-style CSV logs through the Python host layer (run under torchrun, one rank
per GPU): the same windows, shard split, explosion filter and device RMSProp as the C++ `synth_demo --gpus G`, with the
all-reduce behind the C ABI (avs_allreduce_dev).  Prints the parameter fingerprint in synth_demo's format so the two host
languages can be compared bit for bit, and checks that every rank ends with identical parameters.

usage: torchrun --nproc-per-node G scripts/dp_train_csv.py <csv dir> <steps>"""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from auvsl_neural_ode_b200 import capi
from auvsl_neural_ode_b200.trainer import BatchTrainer, shard_range

d, steps = sys.argv[1], int(sys.argv[2])
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
torch.cuda.set_device(local)
dev = torch.device(f"cuda:{local}")

windows = []
for i in range(1, 17):  # Trainer::trainThreads: files 1..16, 200-row windows, stride 200 (Trainer.cpp:204-263)
    a = np.loadtxt(os.path.join(d, f"Train3_data{i:02d}.csv"), delimiter=",", skiprows=1)
    rows = a.shape[0] + 1  # the reference's loader pushes the last row twice when the file ends in a newline (Trainer.cpp:123)
    for j in range(0, rows - 200, 200):
        windows.append(a[j:j + 200])
N, T = len(windows), 200
ctx = capi.Context(local, capi.TIRE_NN)
z = ctx.settle()[6]
b, e = shard_range(N, rank, world)
n = e - b
x0, ctrl, gt = np.zeros((21, n)), np.zeros((T - 1, 2, n)), np.zeros((T, 3, n))
for k, w in enumerate(windows[b:e]):
    s = np.zeros(21)
    s[2], s[3] = math.sin(w[0, 6] / 2.0), math.cos(w[0, 6] / 2.0)  # libm, like the C++ host (np.sin may differ by an ulp)
    s[4:7] = (w[0, 4], w[0, 5], z)
    s[13], s[14], s[15] = w[0, 9], w[0, 10], w[0, 11]
    x0[:, k] = capi.init_state_com(s)
    ctrl[:, 0, k], ctrl[:, 1, k] = w[:-1, 2], w[:-1, 3]
    gt[:, 0, k], gt[:, 1, k], gt[:, 2, k] = w[:, 6], w[:, 4], w[:, 5]

ctx.set_nn_params(capi.nn_default_params())
ctx.reset_optimizer()
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
tr = BatchTrainer(ctx, dev, explode_thresh=100.0, explode_mode=1)
tr.init_native_comm()
tx = [torch.from_numpy(np.ascontiguousarray(v)).to(dev) for v in (x0, ctrl, gt)]
for _ in range(steps):
    red = tr.train_step(*tx)
torch.cuda.synchronize()
p = ctx.get_nn_params()
allp = [torch.zeros(416, dtype=torch.float64, device=dev) for _ in range(world)]
dist.all_gather(allp, torch.from_numpy(p).to(dev))
same = all(torch.equal(allp[0], q) for q in allp)
if rank == 0:
    w_, r_, ver = ctx.comm_info()
    ssum, wsum = 0.0, 0.0
    for i in range(416):  # same accumulation order as synth_demo.cpp
        ssum += float(p[i])
        wsum += float(i + 1) * float(p[i])
    print(f"dp_train_csv: world={world} native_comm={tr.native_comm} nccl={ver} windows={N} n_kept={int(red[417].item())} ranks_bit_identical={same}")
    print(f"params sum {ssum!r} weighted {wsum!r} p[0] {float(p[0])!r} p[103] {float(p[103])!r}")
    assert same
dist.barrier()
ctx.comm_destroy()
dist.destroy_process_group()
