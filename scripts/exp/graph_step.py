"""Experiment: does replaying the step as ONE CUDA graph beat eager launches? (launch gaps / host overhead check)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, bench
import qeb_b200
from qeb_b200 import _lib
from qeb_b200.mirror import ctc as qctc, train_ops
from qeb_b200.mirror.models.model_crnn import CRNN
from qeb_b200.mirror.models.model_unet import UNet
from qeb_b200.mirror.utils import set_bn_eval
dev = "cuda"
torch.manual_seed(42)
prep, crnn = UNet().to(dev), CRNN(95, False).to(dev)
opt = train_ops.Adam(prep.parameters(), lr=5e-5)
x, labels = bench.synth_batch(64, 7); x = x.to(dev)
c2i = {c: i for i, c in enumerate(bench.CHAR_SET)}
y, ys = bench.encode(labels, c2i)
packed = qctc.pack_targets(y, torch.tensor([31] * 64, dtype=torch.int32), ys, dev)
loss_fn = qctc.CTCLoss()
prep.train(); crnn.train(); crnn.apply(set_bn_eval)
def fwd_bwd():
    img = prep(x); scores = crnn(img)
    loss = loss_fn(scores, packed) + train_ops.mse_to_ones(img)
    loss.backward()
    return loss
def step():
    prep.zero_grad(set_to_none=True); crnn.zero_grad(set_to_none=True)
    l = fwd_bwd(); opt.step(); return l
def timeit(fn, n=30):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n): fn()
    e1.record(); t_host = time.perf_counter() - t0; torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, 1e3 * t_host / n
print("eager step: %.3f ms GPU, %.3f ms host issue" % timeit(step))
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3): step()
torch.cuda.current_stream().wait_stream(s)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
prep.zero_grad(set_to_none=True); crnn.zero_grad(set_to_none=True)
try:
    with torch.cuda.graph(g):
        loss = fwd_bwd()
    def gstep():
        g.replay(); opt.step()
    print("graph step: %.3f ms GPU, %.3f ms host issue" % timeit(gstep), "loss", float(loss))
except Exception as e:
    print("capture failed:", repr(e)[:2000])
