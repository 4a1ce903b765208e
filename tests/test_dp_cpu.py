"""Host-side data-parallel logic on CPU: world_size-2 gloo processes (the N > 1 path of bench.py without GPUs)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, mode, out):
    sys.path.insert(0, ROOT)
    import qeb_b200  # noqa: F401
    from qeb_b200.mirror import dist as qdist
    from qeb_b200.mirror.models.model_crnn import _alloc_grads

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.randn(s)) for s in ((4, 3), (7,), (2, 2, 2))]
    if mode == "flat":      # the layout the qeb backward produces: views of one zero-filled buffer
        grads = _alloc_grads(params, [True] * len(params))
        for p, g in zip(params, grads):
            g.copy_(torch.full_like(p, float(rank + 1)) * p.detach())
            p.grad = g
        assert qdist.flat_grad_buffer(params) is not None
    else:                   # independent gradient tensors (foreign autograd functions)
        for p in params:
            p.grad = torch.full_like(p, float(rank + 1)) * p.detach()
        assert qdist.flat_grad_buffer(params) is None
    n = qdist.allreduce_grads(params, average=(mode != "sum"))
    expect = sum(range(1, world + 1)) / (1 if mode == "sum" else world)
    ok = n == 1 and all(torch.allclose(p.grad, expect * p.detach()) for p in params)
    lo, hi = qdist.shard_batch(67, rank, world)
    sizes = [torch.zeros(1, dtype=torch.long) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([hi - lo]))
    ok = ok and sum(int(s) for s in sizes) == 67
    out[rank] = ok
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["flat", "separate", "sum"])
def test_allreduce_grads_gloo_world2(mode):
    world = 2
    port = 29500 + (os.getpid() + hash(mode)) % 2000
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, mode, out), nprocs=world, join=True)
        assert all(out[r] for r in range(world))


def test_single_process_is_a_noop():
    sys.path.insert(0, ROOT)
    import qeb_b200  # noqa: F401
    from qeb_b200.mirror import dist as qdist

    p = torch.nn.Parameter(torch.ones(3))
    p.grad = torch.ones(3)
    assert qdist.allreduce_grads([p]) == 0
    assert qdist.shard_batch(10, 0, 1) == (0, 10)


def test_flat_buffer_is_found_behind_autograd():
    """AccumulateGrad adopts the gradient views with .detach(), which drops `._base`: the flat buffer must still be found
    (it was not in round 1: every data-parallel step silently took the cat / all-reduce / copy-back path, +0.4 ms)."""
    sys.path.insert(0, ROOT)
    import qeb_b200  # noqa: F401
    from qeb_b200.mirror import dist as qdist
    from qeb_b200.mirror.models.model_crnn import _alloc_grads

    params = [torch.nn.Parameter(torch.randn(s)) for s in ((4, 3), (70,), (2, 2, 2), (5,))]

    class Fn(torch.autograd.Function):
        @staticmethod
        def forward(ctx, *ps):
            ctx.ps = ps
            return sum(p.sum() for p in ps)

        @staticmethod
        def backward(ctx, g):
            grads = _alloc_grads(ctx.ps, [True] * len(ctx.ps))
            for i, gr in enumerate(grads):
                gr.fill_(float(i + 1))
            return tuple(grads)

    Fn.apply(*params).backward()
    assert all(p.grad._base is None for p in params)            # the situation the storage test exists for
    flat = qdist.flat_grad_buffer(params)
    assert flat is not None and flat.numel() >= sum(p.numel() for p in params)
    flat.mul_(2.0)                                               # the all-reduce acts on the gradients themselves
    assert all(torch.equal(p.grad, torch.full_like(p, 2.0 * (i + 1))) for i, p in enumerate(params))
    assert qdist.flat_grad_buffer([params[0], params[3]]) is None   # a subset with foreign tensors in between
    assert qdist.flat_grad_buffer(params[:2]) is not None           # a contiguous prefix is fine


def _topk_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import numpy as np
    import qeb_b200  # noqa: F401
    from qeb_b200.mirror import dist as qdist
    from oracle import pyoracle as po

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.RandomState(7)
    # tie-heavy values (CERs are small rationals), ragged shards, one shard shorter than k
    allv = (rng.randint(0, 12, size=1000) / 4.0).astype(np.float32)
    cuts = [0, 5, 1000]
    shard = allv[cuts[rank]:cuts[rank + 1]]
    ok = True
    for k in (1, 7, 64, 1000, 1500):
        got = qdist.global_topk(shard, k, local_topk=lambda v, kk: po.topk_query(v, kk)).numpy()
        want = np.argsort(-allv.astype(np.float64), kind="stable")[:k]
        ok = ok and np.array_equal(got, want)
    ok = ok and len(qdist.global_topk(shard, 0, local_topk=lambda v, kk: po.topk_query(v, kk))) == 0
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_global_topk_gloo_world2():
    """8(e): global top-k = per-rank stable top-k + all-gather of the candidates + merge; equals the one-process stable sort."""
    world = 2
    port = 31500 + os.getpid() % 2000
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_topk_worker, args=(world, port, out), nprocs=world, join=True)
        assert all(out[r] for r in range(world))
