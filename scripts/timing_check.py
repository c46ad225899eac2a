import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from auvsl_neural_ode_b200 import capi, synth
from auvsl_neural_ode_b200.trainer import BatchTrainer
N, T = 4096, 200
dev = torch.device("cuda:0")
ctx = capi.Context(0, capi.TIRE_NN)
stream = torch.cuda.current_stream()
ctx.set_stream(stream.cuda_stream)
x0, ctrl, gt = synth.make_batch(N, T)
d = [torch.from_numpy(a).to(dev) for a in (x0, ctrl, gt)]
tr = BatchTrainer(ctx, dev, explode_mode=1)
for _ in range(3): tr.train_step(*d)
torch.cuda.synchronize()
for K in (1, 2, 5, 10, 20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    e0.record(stream)
    for _ in range(K): tr.train_step(*d)
    e1.record(stream)
    torch.cuda.synchronize(); wall = (time.perf_counter() - t0) * 1e3
    f, b = ctx.last_kernel_ms()
    print(f"K={K:2d} events {e0.elapsed_time(e1)/K:.3f} ms/step  wall {wall/K:.3f} ms/step  last-step kernels fwd {f:.3f} + bwd {b:.3f} = {f+b:.3f}")
