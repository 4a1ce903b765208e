"""Experiment: where the extra time of the data-parallel step goes (graph replay / + all-reduce / + Adam), N ranks."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, torch.distributed as dist, bench
import qeb_b200
from qeb_b200.graphs import GraphedStep, StaticTargets
from qeb_b200.mirror import ctc as qctc, dist as qdist, train_ops
from qeb_b200.mirror.models.model_crnn import CRNN
from qeb_b200.mirror.models.model_unet import UNet
from qeb_b200.mirror.utils import set_bn_eval
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
if world > 1: dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(42)
prep, crnn = UNet().to(dev), CRNN(95, False).to(dev)
opt = train_ops.Adam(prep.parameters(), lr=5e-5)
x, labels = bench.synth_batch(64, 7 + rank); x = x.to(dev)
c2i = {c: i for i, c in enumerate(bench.CHAR_SET)}
y, ys = bench.encode(labels, c2i)
tg = StaticTargets(64, 31, dev).load(y, torch.tensor([31] * 64, dtype=torch.int32), ys)
loss_fn = qctc.CTCLoss()
prep.train(); crnn.train(); crnn.apply(set_bn_eval)
def fwd_bwd():
    img = prep(x); scores = crnn(img)
    loss = loss_fn(scores, tg) + train_ops.mse_to_ones(img); loss.backward(); return loss
gs = GraphedStep(fwd_bwd, modules=[prep, crnn])
params = list(prep.parameters())
def t(fn, n=40):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
a = t(lambda: gs())
b = t(lambda: (gs(), qdist.allreduce_grads(params, average=True)))
c = t(lambda: (gs(), qdist.allreduce_grads(params, average=True), opt.step()))
d = t(lambda: (gs(), opt.step()))
tiny = torch.zeros(1, device=dev)
e = t(lambda: (gs(), dist.all_reduce(tiny) if world > 1 else None))
# all-reduce captured inside the graph
def fwd_bwd_ar():
    l = fwd_bwd(); qdist.allreduce_grads(params, average=True); return l
try:
    gs2 = GraphedStep(fwd_bwd_ar, modules=[prep, crnn])
    params2 = list(prep.parameters())
    f = t(lambda: gs2())
    g = t(lambda: (gs2(), opt.step()))
except Exception as ex:
    f = g = float("nan"); print("capture with all-reduce failed:", repr(ex)[:300])
if rank == 0: print(f"world {world}: replay + 4-byte all-reduce {e:.3f}; all-reduce inside the graph {f:.3f}, + Adam {g:.3f}")
if rank == 0: print(f"world {world}: replay {a:.3f} ms, + all-reduce {b:.3f}, + all-reduce + Adam {c:.3f}, replay + Adam {d:.3f}")
if world > 1: dist.destroy_process_group()
