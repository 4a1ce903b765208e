"""Data-parallel glue for the training step: one process per GPU, one gradient all-reduce per step.

The reference has no distributed code; sharding the batch (or the noised copies of the inner loop,
train_nn_patch.py:278-303) across ranks is new functionality defined by north_star. Every kernel of the step is
sample-local, so the only exchange is the sum / mean of the parameter gradients. The qeb backward hands autograd views
of ONE flat gradient buffer per network, so the exchange is a single NCCL call on 31 MB (UNet) or 35 MB (CRNN).
"""
import torch
import torch.distributed as dist


def flat_grad_buffer(params):
    """The single tensor that covers all `.grad`s of `params`, or None if they do not live back to back in one storage.
    The qeb backward carves the gradients of a network out of one zero-filled buffer (64-float aligned slices);
    AccumulateGrad adopts them with `.detach()`, which drops the view relation (`._base`), so the test is on the storage:
    same storage, contiguous tensors, and a covered span no larger than the tensors plus their alignment padding."""
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return None
    g0 = grads[0]
    st = g0.untyped_storage()
    for g in grads:
        if g.dtype != g0.dtype or g.device != g0.device or not g.is_contiguous() or g.untyped_storage().data_ptr() != st.data_ptr():
            return None
    lo = min(g.storage_offset() for g in grads)
    hi = max(g.storage_offset() + g.numel() for g in grads)
    if hi - lo > sum(g.numel() for g in grads) + 64 * len(grads):
        return None   # a subset with other tensors in between: reducing the span would touch gradients not asked for
    return torch.empty(0, dtype=g0.dtype, device=g0.device).set_(st, lo, (hi - lo,))


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
