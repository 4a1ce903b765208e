"""Experiment: forward-only vs full step time (eager, events), and inference throughput in eval mode."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, bench
import qeb_b200
from qeb_b200.mirror import ctc as qctc, train_ops
from qeb_b200.mirror.models.model_crnn import CRNN
from qeb_b200.mirror.models.model_unet import UNet
from qeb_b200.mirror.utils import set_bn_eval
dev = "cuda"
torch.manual_seed(42)
prep, crnn = UNet().to(dev), CRNN(95, False).to(dev)
x, labels = bench.synth_batch(64, 7); x = x.to(dev)
def t(fn, n=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
prep.train(); crnn.train(); crnn.apply(set_bn_eval)
with torch.no_grad():
    a = t(lambda: prep(x)); img = prep(x); b = t(lambda: crnn(img))
print(f"train-mode forward, no grad: UNet {a:.3f} ms, CRNN {b:.3f} ms")
prep.eval(); crnn.eval()
with torch.no_grad():
    a = t(lambda: prep(x)); img = prep(x); b = t(lambda: crnn(img))
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3): crnn(prep(x))
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    with torch.cuda.graph(g):
        out = crnn(prep(x))
    c = t(lambda: g.replay())
print(f"eval-mode forward: UNet {a:.3f} ms, CRNN {b:.3f} ms; both as one graph {c:.3f} ms = {64/c*1e3:.0f} patches/s inference")
