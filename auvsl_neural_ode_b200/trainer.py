"""Data-parallel training / evaluation steps with the reference Trainer's semantics.

Mirrors, for a whole batch of windows at once (paths relative to the reference repository):
  * Trainer::trainThreads + combineResults + updateParams (src/Trainer.cpp:204-263, 661-689, 428-453):
    per-sample loss/gradient, per-sample explosion filter, sum, RMSProp with the float-truncated
    gradient — `train_step`
  * Trainer::evaluateTrajectory (src/Trainer.cpp:550-589): forward rollout, loss =
    sqrt(lin_mse_final) / traj_len — `eval_losses`

One process per GPU.  The batch shards by trajectory (contiguous ranges, no data-path collective);
training adds ONE all-reduce (sum) of `reduced` = grad_sum[416] | loss_sum | n_kept per step, after
which every rank applies the identical RMSProp update on its own device copy of the parameters, so the
replicas stay bit-identical without a broadcast.  The all-reduce is the C ABI's own
(`avs_allreduce_dev`: ncclAllReduce on the context's stream, straight from the device buffer) once
`init_native_comm()` has given the context its NCCL rank; `torch.distributed` then only ships the 128-byte
NCCL unique id at start-up.  Without it (CPU / gloo tests with a stand-in backend) the step falls back
to `dist.all_reduce` on the same tensor.

`backend` is anything with the `capi.Context` device-pointer interface; tests drive the host logic on
CPU/gloo with a stand-in backend.
"""
import torch
import torch.distributed as dist

REDUCED_LEN = 418


def shard_range(n_total, rank, world):
    """Contiguous shard [begin, end) of rank; sizes differ by at most one."""
    base, rem = divmod(n_total, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


class BatchTrainer:
    def __init__(self, backend, device, lr=1e-3, l1_weight=0.0, explode_thresh=100.0, explode_mode=1, group=None):
        self.backend = backend
        self.device = device
        self.lr, self.l1_weight = lr, l1_weight
        self.explode_thresh, self.explode_mode = explode_thresh, explode_mode
        self.group = group
        self.reduced = torch.zeros(REDUCED_LEN, dtype=torch.float64, device=device)
        self._loss = None
        self.native_comm = False

    def init_native_comm(self):
        """Give the backend context its NCCL rank (include/avsim.h, avs_comm_*): rank 0 draws the unique id, torch.distributed
        broadcasts the 128 bytes, every rank joins.  Afterwards train_step never touches torch.distributed."""
        world = self._world()
        if world == 1:
            return
        from . import capi
        rank = dist.get_rank(self.group)
        box = [capi.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0, group=self.group)
        self.backend.comm_init_rank(box[0], world, rank)
        self.native_comm = True

    def _world(self):
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    def train_step(self, x0, ctrl, gt, substeps=10):
        """x0 [21][n], ctrl [T-1][2][n], gt [T][3][n]: this rank's shard, resident on `device`.
        Returns the `reduced` tensor after the all-reduce (grad_sum | loss_sum | n_kept)."""
        n = x0.shape[1]
        T = gt.shape[0]
        if self._loss is None or self._loss.numel() != n:
            self._loss = torch.empty(n, dtype=torch.float64, device=self.device)
        self.backend.rollout_grad_dev(n, T, substeps, x0.data_ptr(), ctrl.data_ptr(), gt.data_ptr(), self._loss.data_ptr(),
                                      self.reduced.data_ptr(), 0, self.explode_thresh, self.explode_mode)
        if self.native_comm:
            self.backend.allreduce_dev(self.reduced.data_ptr())  # ncclAllReduce on the context's stream, in place
        elif self._world() > 1:
            dist.all_reduce(self.reduced, op=dist.ReduceOp.SUM, group=self.group)
        self.backend.rmsprop_step_dev(self.reduced.data_ptr(), self.lr, self.l1_weight)
        return self.reduced

    @staticmethod
    def eval_losses(x_final, gt):
        """Trainer::evaluateTrajectory's metric from final states: x_final [21][n], gt [T][3][n] -> loss [n]."""
        d = gt[1:, 1:3, :] - gt[:-1, 1:3, :]
        traj_len = torch.sqrt((d * d).sum(dim=1)).sum(dim=0)
        ex = gt[-1, 1, :] - x_final[4, :]
        ey = gt[-1, 2, :] - x_final[5, :]
        return torch.sqrt(ex * ex + ey * ey) / traj_len
