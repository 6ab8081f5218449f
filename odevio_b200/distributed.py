"""Multi-GPU plumbing for the regressor path: one process per GPU, rows sharded, NCCL only for
the training gradient exchange.

* Inference: sequences are independent (per-row step control, no cross-row op in
  src/models/PoseODERNN.py:88-123), so every rank integrates its own contiguous shard of rows --
  no data-path collective.
* Training: the reference optimises exactly the ``Pose_net`` parameters (utils/utils.py:143-147)
  with a mean-reduction loss (scripts/train_model.py:72-77), clips the global norm and steps Adam
  (:83-86).  With equal shards the global-batch gradient is the mean of the per-rank gradients:
  one flat fp32 bucket (3.77 M values = 15 MB at the default shapes) is all-reduced (sum) and
  scaled by 1/world BEFORE the clip, so the clip sees the global-batch norm.

``torch.distributed`` is plumbing here; the backend is "nccl" on the GPU box and "gloo" in the
CPU tests (tests/test_distributed_cpu.py).
"""

import torch
import torch.distributed as dist


def shard_rows(n_rows, rank, world_size):
    """Contiguous, balanced [start, stop) of this rank's sequences (first `n_rows % world` ranks get one more)."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of size {world_size}")
    base, extra = divmod(n_rows, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def pose_net_params(model):
    """The parameter set the reference optimises (utils/utils.py:143-147), regressor group first."""
    return list(model.get_regressor_params()) + list(model.get_other_params())


def allreduce_pose_net_grads(model, world_size=None, group=None, shard_rows_count=None, global_rows=None):
    """Average the Pose_net gradients over ranks through ONE flat bucket.

    With unequal shards pass ``shard_rows_count`` / ``global_rows``: each rank's gradient of its
    shard-mean loss is weighted by shard/global so the result is the global-batch mean gradient.
    Returns the number of all-reduced values."""
    if world_size is None:
        world_size = dist.get_world_size(group) if dist.is_initialized() else 1
    params = [p for p in pose_net_params(model) if p.requires_grad]
    for p in params:
        if p.grad is None:
            p.grad = torch.zeros_like(p)
    if world_size == 1:
        return sum(p.numel() for p in params)
    weight = (1.0 / world_size) if shard_rows_count is None else (float(shard_rows_count) / float(global_rows))
    flat = torch.cat([p.grad.reshape(-1) for p in params]).mul_(weight)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for p in params:
        n = p.numel()
        p.grad.copy_(flat[off:off + n].view_as(p.grad))
        off += n
    return off


def pose_loss(poses, gts):
    """scripts/train_model.py:72-77."""
    angle = torch.nn.functional.mse_loss(poses[:, :, :3], gts[:, :, :3])
    trans = torch.nn.functional.mse_loss(poses[:, :, 3:], gts[:, :, 3:])
    return 100 * angle + trans


def make_optimizer(model, lr=1e-4, weight_decay=5e-5):
    """utils/utils.py:115-130 (Adam over the two Pose_net groups, [other, regressor], both at lr_warmup = 1e-4)."""
    return torch.optim.Adam([{"params": list(model.get_other_params()), "lr": lr},
                             {"params": list(model.get_regressor_params()), "lr": lr}],
                            lr=lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=weight_decay)


def train_step(model, optimizer, fv, fi, ts, gts, world_size=1, group=None, clip=5.0):
    """One optimisation step on this rank's shard (scripts/train_model.py:69-86): forward, loss,
    backward (fused kernels), gradient all-reduce, global-norm clip, Adam."""
    optimizer.zero_grad(set_to_none=True)
    poses, _ = model(fv, fi, ts, prev=None)
    loss = pose_loss(poses, gts)
    loss.backward()
    allreduce_pose_net_grads(model, world_size, group)
    if clip:
        torch.nn.utils.clip_grad_norm_(pose_net_params(model), max_norm=clip)
    optimizer.step()
    return loss.detach()
