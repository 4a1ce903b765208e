"""Data-parallel glue for the training step: one process per GPU, one gradient all-reduce per step.

The reference has no distributed code; sharding the batch (or the noised copies of the inner loop,
train_nn_patch.py:278-303) across ranks is new functionality defined by north_star. Every kernel of the step is
sample-local, so the only exchange is the sum / mean of the parameter gradients. The qeb backward hands autograd views
of ONE flat gradient buffer per network, so the exchange is a single NCCL call on 31 MB (UNet) or 35 MB (CRNN).
"""
import numpy as np
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


class BucketedAllReduce:
    """The UNet gradient exchange of a data-parallel step, overlapped with the backward pass ("bucketed in
    reverse-layer order and overlapped with the remaining wgrads").

    Gradients become final in the order the backward pass - and, behind it, the weight-gradient stream - walks the network: the
    decoder first, then the bottleneck and encoder block 4, the full-resolution encoder blocks last; that is the flat gradient
    buffer from its END. `qeb_unet_backward_bucketed` records an event when each of the first two ranges (12.2 MB, 17.7 MB) is
    final; this object all-reduces them on a communication stream that waits for the events, i.e. while the rest of the
    backward runs, and the remaining 1.1 MB on the caller's stream afterwards:

        ar = BucketedAllReduce(prep_model)            # once; installs the events on the module
        loss.backward(); ar(); optimizer.step()       # every step (also inside a GraphedStep capture)

    Works inside a CUDA-graph capture (the event records, the stream fork / join and the NCCL calls are captured)."""

    def __init__(self, unet, average=True, group=None):
        self.unet, self.average, self.group = unet, average, group
        self.active = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        self.events = self.comm = None
        if self.active and next(unet.parameters()).is_cuda:
            self.events = (torch.cuda.Event(), torch.cuda.Event())
            for ev in self.events:
                ev.record()            # torch creates the cudaEvent lazily at the first record: the C ABI needs the handle
                if not ev.cuda_event:
                    raise RuntimeError("BucketedAllReduce: could not create the CUDA event")
            self.comm = torch.cuda.Stream()
            unet._qeb_tail_event = self.events

    def close(self):
        if getattr(self.unet, "_qeb_tail_event", None) is self.events:
            self.unet._qeb_tail_event = None

    def __call__(self):
        if not self.active:
            return 0
        params = self.unet.qeb_parameters()
        flat = flat_grad_buffer(params)
        if flat is None or self.events is None:
            return allreduce_grads(params, self.average, self.group)
        lo = min(p.grad.storage_offset() for p in params)
        cuts = [params[i].grad.storage_offset() - lo for i in self.unet.QEB_BUCKET_STARTS]      # descending offsets
        op = dist.ReduceOp.AVG if self.average else dist.ReduceOp.SUM
        cur = torch.cuda.current_stream()
        end = flat.numel()
        for ev, cut in zip(self.events, cuts):          # ranges in the order they become final
            part = flat[cut:end]
            self.comm.wait_event(ev)                    # fires in the middle of the backward pass
            with torch.cuda.stream(self.comm):
                dist.all_reduce(part, op=op, group=self.group)
            part.record_stream(self.comm)
            end = cut
        dist.all_reduce(flat[:end], op=op, group=self.group)
        cur.wait_stream(self.comm)
        return len(cuts) + 1


def shard_batch(n, rank, world):
    """Contiguous shard [lo, hi) of a batch of n samples for `rank` (sizes differ by at most one)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def global_topk(local_cers, k, group=None, local_topk=None):
    """Stable global top-k over the ranks' CER shards: the indices `argsort(all, descending, stable)[:k]`
    of the concatenation of the shards in rank order - what `TopKCERSampler.query` (selection_utils.py:144-151) and
    `pruning/methods.topk` (:5-8) compute on one process - returned on every rank as an int64 tensor of GLOBAL indices.

    Each rank reduces its shard to its own stable top-k on the device (`qeb_cer_topk_segmented`), the ranks all-gather the
    k (value, index) candidates (12 bytes each) and merge them by (value descending, global index ascending); any element of
    the global top-k is in its shard's top-k, and ties keep the lower global index exactly as the stable sort does.
    local_topk(values_f32_numpy, k) -> int64 indices: the shard reduction; defaults to the device kernel (tests on CPU
    processes pass the oracle here - the host-side merge is what they cover)."""
    vals = torch.as_tensor(local_cers, dtype=torch.float32).cpu().numpy().reshape(-1)
    if local_topk is None:
        from . import selection_utils
        local_topk = selection_utils.topk_cer_indices
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    world = dist.get_world_size(group) if multi else 1
    rank = dist.get_rank(group) if multi else 0
    k = int(k)
    if k <= 0:
        return torch.zeros(0, dtype=torch.long)
    kl = min(k, len(vals))
    idx = np.asarray(local_topk(vals, kl), dtype=np.int64).reshape(-1) if kl > 0 else np.zeros(0, dtype=np.int64)
    if not multi:
        return torch.from_numpy(idx)
    # shard sizes -> global offsets; candidates padded to k entries (count travels with them)
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    sizes = [torch.zeros(1, dtype=torch.long, device=dev) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([len(vals)], dtype=torch.long, device=dev), group=group)
    offs = np.concatenate([[0], np.cumsum([int(s) for s in sizes])])
    cand_v = torch.full((k,), float("-inf"), dtype=torch.float32)
    cand_i = torch.full((k + 1,), -1, dtype=torch.long)
    cand_v[:kl] = torch.from_numpy(vals[idx])
    cand_i[:kl] = torch.from_numpy(idx + offs[rank])
    cand_i[k] = kl
    all_v = [torch.empty(k, dtype=torch.float32, device=dev) for _ in range(world)]
    all_i = [torch.empty(k + 1, dtype=torch.long, device=dev) for _ in range(world)]
    dist.all_gather(all_v, cand_v.to(dev), group=group)
    dist.all_gather(all_i, cand_i.to(dev), group=group)
    vs, gs = [], []
    for v, i in zip(all_v, all_i):
        i = i.cpu().numpy()
        n = int(i[k])
        vs.append(v.cpu().numpy()[:n])
        gs.append(i[:n])
    v, g = np.concatenate(vs), np.concatenate(gs)
    order = np.lexsort((g, -v.astype(np.float64)))   # primary: value descending; ties: global index ascending
    return torch.from_numpy(g[order[:k]].astype(np.int64))
