"""Data-parallel plumbing: independent clips/windows are sharded across ranks (one process per GPU); the only
exchange of a training step is ONE all-reduce of the flat gradient buffer (NCCL on GPUs; gloo in the CPU tests).
Inference and the mel front end shard clips with no collective at all."""
import os

import torch
import torch.distributed as dist


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def shard_range(n, rank, world):
    """Contiguous shard [lo, hi) of n independent units for `rank`; sizes differ by at most one."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_sum_(flat, group=None):
    """In-place sum of the flat gradient over all ranks (the optimizer folds in 1/world via grad_scale)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat


def mean_of_means_is_global_mean(local_mean_grad, local_count, group=None):
    """Every loss term is a mean over the batch (train.py:340-421): with EQUAL per-rank batch sizes the global-batch
    gradient is the plain average of the per-rank gradients.  Asserts equal counts across ranks."""
    t = torch.tensor([float(local_count)])
    lo, hi = t.clone(), t.clone()
    if dist.is_initialized():
        dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=group)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=group)
    if lo.item() != hi.item():
        raise ValueError("data-parallel ranks must use equal per-rank batch sizes")
    g = local_mean_grad.clone()
    allreduce_sum_(g, group)
    return g / (dist.get_world_size(group) if dist.is_initialized() else 1)
