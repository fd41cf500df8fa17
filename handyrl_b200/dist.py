"""Multi-GPU plumbing of the learner: one process per GPU, batch sharded over ranks, ONE all-reduce(SUM).

The reference's multi-GPU mode is a single-process nn.DataParallel whose loss and backward run on
GPU 0 (train.py:339-340, 366).  Here every rank runs the full step on its B/N shard and the flat
gradient bucket (with the six loss sums appended) is all-reduced with SUM -- the loss is a sum over
the batch, not a mean (train.py:202-213), so SUM of shard gradients == full-batch gradient, and the
clip threshold / learning-rate schedule see global quantities (SURVEY.md hard part 7).
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise torch.distributed from RANK / WORLD_SIZE / MASTER_* (torchrun); returns (rank, world, local)."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        kw = {}
        if backend == 'nccl':
            torch.cuda.set_device(local)
            kw['device_id'] = torch.device('cuda', local)
        dist.init_process_group(backend, **kw)
    return rank, world, local


def shard_bounds(B, rank, world):
    """Contiguous split of the batch dimension; the first B % world ranks get one extra window."""
    base, extra = divmod(B, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(batch, rank, world):
    """Slice every (B, ...) tensor of a make_batch dict (nested observations included) to this rank's windows."""
    from .batch import tree_map
    B = batch['action'].shape[0]
    lo, hi = shard_bounds(B, rank, world)
    return tree_map(lambda t: t[lo:hi].contiguous(), batch)


def allreduce_sum_(flat, group=None):
    """In-place SUM all-reduce of the flat gradient bucket (gradients + appended loss sums)."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat
