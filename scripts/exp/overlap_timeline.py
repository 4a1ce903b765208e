"""When do the bucketed all-reduces run relative to the backward pass? Eager step, CUDA events on the communication stream
and on the compute stream. torchrun, 2+ GPUs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, torch.distributed as dist
import bench as BB
import qeb_b200
from qeb_b200.mirror import ctc as qctc, dist as qdist, train_ops
from qeb_b200.mirror.models.model_crnn import CRNN
from qeb_b200.mirror.models.model_unet import UNet
from qeb_b200.mirror.utils import set_bn_eval
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(42)
prep, crnn = UNet().to(dev), CRNN(95, False).to(dev)
x, labels = BB.synth_batch(64, 7 + rank); x = x.to(dev)
c2i = {c: i for i, c in enumerate(BB.CHAR_SET)}
y, ys = BB.encode(labels, c2i)
tg = qctc.pack_targets(y, torch.tensor([31] * 64, dtype=torch.int32), ys, dev)
loss_fn = qctc.CTCLoss()
prep.train(); crnn.train(); crnn.apply(set_bn_eval)
ar = qdist.BucketedAllReduce(prep)
E = lambda: torch.cuda.Event(enable_timing=True)
def step(measure=False):
    prep.zero_grad(set_to_none=True); crnn.zero_grad(set_to_none=True)
    img = prep(x); loss = loss_fn(crnn(img), tg) + train_ops.mse_to_ones(img)
    t0 = E(); t0.record()
    loss.backward()
    t1 = E(); t1.record()                     # end of the backward on the compute stream
    ev = []
    if measure:
        for e_ in ar.events:                  # when did each bucket become final?
            m = E(); ar.comm.wait_event(e_); m.record(ar.comm); ev.append(m)
    ar()
    c_end = E(); c_end.record(ar.comm)
    t2 = E(); t2.record()
    return t0, t1, ev, c_end, t2
for _ in range(5): step()
torch.cuda.synchronize(); dist.barrier()
t0, t1, ev, c_end, t2 = step(True)
torch.cuda.synchronize()
if rank == 0:
    print(f"backward {t0.elapsed_time(t1):.3f} ms; bucket 0 final at {t0.elapsed_time(ev[0]):.3f}, bucket 1 final at {t0.elapsed_time(ev[1]):.3f}; "
          f"comm stream done at {t0.elapsed_time(c_end):.3f}; step (backward + exchange) done at {t0.elapsed_time(t2):.3f} ms")
dist.destroy_process_group()
