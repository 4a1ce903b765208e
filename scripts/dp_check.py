#!/usr/bin/env python
"""Data-parallel gradient exchange on real GPUs (torchrun, one rank per GPU): the bucketed, overlapped all-reduce of the UNet
gradients (mirror/dist.BucketedAllReduce + qeb_unet_backward_bucketed) against the plain one-collective exchange and against
the mean of the ranks' local gradients; eagerly and inside a GraphedStep capture. The SUM exchange of the CRNN gradients
(jitter step) the same way.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 scripts/dp_check.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import bench as BB
import qeb_b200
from qeb_b200.graphs import GraphedStep
from qeb_b200.mirror import ctc as qctc, dist as qdist, train_ops
from qeb_b200.mirror.models.model_crnn import CRNN
from qeb_b200.mirror.models.model_unet import UNet
from qeb_b200.mirror.utils import set_bn_eval

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(42)
prep, crnn = UNet().to(dev), CRNN(95, False).to(dev)
x, labels = BB.synth_batch(16, 7 + rank)
x = x.to(dev)
c2i = {c: i for i, c in enumerate(BB.CHAR_SET)}
y, ys = BB.encode(labels, c2i)
tg = qctc.pack_targets(y, torch.tensor([31] * 16, dtype=torch.int32), ys, dev)
loss_fn = qctc.CTCLoss()
prep.train(); crnn.train(); crnn.apply(set_bn_eval)


def fwd_bwd():
    prep.zero_grad(set_to_none=True); crnn.zero_grad(set_to_none=True)
    img = prep(x)
    loss = loss_fn(crnn(img), tg) + train_ops.mse_to_ones(img)
    loss.backward()
    return loss


def grads():
    return [p.grad.clone() for p in prep.qeb_parameters()]


def worst(a, b):
    return max(float((u - v).abs().max() / v.abs().max().clamp_min(1e-30)) for u, v in zip(a, b))


def same_on_all_ranks(ts):
    for u in ts:
        t = [torch.empty_like(u) for _ in range(world)]
        dist.all_gather(t, u)
        assert all(torch.equal(t[0], v) for v in t)


ar = qdist.BucketedAllReduce(prep, average=True)       # installs the event before the first backward
fwd_bwd()
local = grads()
want = []
for g in local:                                   # the definition: mean over the ranks of the local gradients
    t = g.clone(); dist.all_reduce(t); want.append(t / world)
n = ar()                                          # bucketed exchange on the gradients of THIS backward
assert n == 3, n
bucketed = grads()
for p, g in zip(prep.qeb_parameters(), local):    # put the local gradients back (in place: the flat buffer stays)
    p.grad.copy_(g)
qdist.allreduce_grads(prep.qeb_parameters(), average=True)
plain = grads()
torch.cuda.synchronize()
e_plain, e_bucket = worst(plain, want), worst(bucketed, want)
assert e_plain < 1e-6 and e_bucket < 1e-6, (e_plain, e_bucket)
same_on_all_ranks(bucketed)


def step():
    loss = fwd_bwd_static()
    ar()
    return loss


def fwd_bwd_static():
    img = prep(x)
    loss = loss_fn(crnn(img), tg) + train_ops.mse_to_ones(img)
    loss.backward()
    return loss


g = GraphedStep(step, modules=[prep, crnn], warmup=2)               # NCCL calls, event and stream fork / join captured
g(); torch.cuda.synchronize()
graphed = grads()
same_on_all_ranks(graphed)
assert all(torch.isfinite(u).all() for u in graphed)
e_graph = worst(graphed, want)      # another backward pass: split-K reductions and ReLU flips make passes differ, loosely bounded
assert e_graph < 0.5, e_graph
g.close()                           # before the communicator goes (a live graph with captured NCCL calls hangs the teardown)
# SUM exchange of the CRNN gradients (inner-loop copies sharded over the ranks)
crnn.zero_grad(set_to_none=True); crnn.train()
loss_fn(crnn(x), tg).backward()
lc = [p.grad.clone() for p in crnn.parameters()]
qdist.allreduce_grads(list(crnn.parameters()), average=False)
for p, g0 in zip(crnn.parameters(), lc):
    t = g0.clone(); dist.all_reduce(t)
    assert float((p.grad - t).abs().max()) <= 1e-6 * float(t.abs().max()) + 1e-12
if rank == 0:
    print(f"dp_check ok: world {world}, worst relative deviation from the mean of local gradients: plain {e_plain:.2e}, bucketed {e_bucket:.2e}, graphed {e_graph:.2e}")
dist.destroy_process_group()
