"""Data-parallel glue for the training step: one process per GPU, one gradient all-reduce per step.

The reference has no distributed code; sharding the batch (or the noised copies of the inner loop,
train_nn_patch.py:278-303) across ranks is new functionality defined by north_star. Every kernel of the step is
sample-local, so the only exchange is the sum / mean of the parameter gradients. The qeb backward hands autograd views
of ONE flat gradient buffer per network, so the exchange is a single NCCL call on 31 MB (UNet) or 35 MB (CRNN).
"""
import torch
import torch.distributed as dist


def flat_grad_buffer(params):
    """The single tensor all `.grad`s of `params` are views of, or None if they do not share one base."""
    base = None
    for p in params:
        g = p.grad
        if g is None:
            continue
        b = g._base if g._base is not None else None
        if b is None:
            return None
        if base is None:
            base = b
        elif b.data_ptr() != base.data_ptr() or b.numel() != base.numel():
            return None
    return base


def allreduce_grads(params, average=True, group=None):
    """Sum (inner-loop copies: the patch trainer accumulates, train_nn_patch.py:301-303) or average (batch shards: the
    losses are means over the batch) the gradients of `params` over the ranks, in place. Returns the number of
    collectives issued."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return 0
    params = [p for p in params if p.grad is not None]
    if not params:
        return 0
    world = dist.get_world_size(group)
    flat = flat_grad_buffer(params)
    if flat is not None:
        if average and flat.is_cuda and dist.get_backend(group) == "nccl":
            dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=group)   # the division happens inside the collective
            return 1
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        if average:
            flat.div_(world)
        return 1
    bucket = torch.cat([p.grad.reshape(-1) for p in params])
    dist.all_reduce(bucket, op=dist.ReduceOp.SUM, group=group)
    if average:
        bucket.div_(world)
    off = 0
    for p in params:
        p.grad.copy_(bucket[off:off + p.numel()].view_as(p.grad))
        off += p.numel()
    return 1


def shard_batch(n, rank, world):
    """Contiguous shard [lo, hi) of a batch of n samples for `rank` (sizes differ by at most one)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)
