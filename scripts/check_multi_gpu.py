"""One-off multi-GPU consistency check (run under torchrun, 2+ GPUs): after K data-parallel training steps the
parameter vectors of all ranks are bit-identical, and they equal a single-rank run over the concatenated batch
up to summation order (allclose 1e-12)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from auvsl_neural_ode_b200 import capi, synth
from auvsl_neural_ode_b200.trainer import BatchTrainer, shard_range

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
torch.cuda.set_device(local)
dev = torch.device(f"cuda:{local}")
N, T, K = 1024, 40, 3
x0, ctrl, gt = synth.make_batch(N, T, seed=99)

def run(ctx, parts, steps):
    ctx.set_nn_params(capi.nn_default_params()); ctx.reset_optimizer()
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    tr = BatchTrainer(ctx, dev, explode_mode=1)
    d = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in parts]
    for _ in range(steps):
        red = tr.train_step(*d)
    torch.cuda.synchronize()
    return ctx.get_nn_params(), red.cpu().numpy()

ctx = capi.Context(local, capi.TIRE_NN)
b, e = shard_range(N, rank, world)
p_dp, red_dp = run(ctx, (x0[:, b:e], ctrl[:, :, b:e], gt[:, :, b:e]), K)
allp = [torch.zeros(416, dtype=torch.float64, device=dev) for _ in range(world)]
dist.all_gather(allp, torch.from_numpy(p_dp).to(dev))
same = all(torch.equal(allp[0], q) for q in allp)
if rank == 0:
    saved_group = dist.group.WORLD
    # single-rank reference over the whole batch (no collective: world size seen by the trainer is patched to 1)
    import auvsl_neural_ode_b200.trainer as trmod
    orig = trmod.BatchTrainer._world
    trmod.BatchTrainer._world = lambda self: 1
    p_one, red_one = run(ctx, (x0, ctrl, gt), K)
    trmod.BatchTrainer._world = orig
    ok = np.allclose(p_dp, p_one, rtol=1e-12, atol=0) and red_dp[417] == N
    print(f"multi-gpu check: world={world} ranks_bit_identical={same} matches_single_rank={ok} max|dp-one|={np.max(np.abs(p_dp-p_one)):.3e} n_kept={red_dp[417]}")
    assert same and ok
dist.barrier()
dist.destroy_process_group()
