"""Experiment: per-step phase times (clock64) of the LSTM recurrence kernels, CTA 0."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import qeb_b200
from qeb_b200 import _lib
L = _lib.load()
DEV = "cuda"; T, B = 31, 64
torch.manual_seed(0)
w = torch.randn(2, 1024, 256, device=DEV) * 0.05
gates0 = torch.randn(T, B, 2, 1024, device=DEV); dy = torch.randn(T, B, 512, device=DEV)
cells = torch.empty(T, B, 2, 256, device=DEV); y = torch.empty(T, B, 512, device=DEV)
st = torch.cuda.current_stream().cuda_stream
def fwd(g): _lib.call("qeb_lstm_layer_fwd", g.data_ptr(), w[0].data_ptr(), w[1].data_ptr(), cells.data_ptr(), y.data_ptr(), T, B, st)
def bwd(g): _lib.call("qeb_lstm_layer_bwd", g.data_ptr(), cells.data_ptr(), dy.data_ptr(), w[0].data_ptr(), w[1].data_ptr(), T, B, st)
for name, fn, labels in (("fwd", fwd, ["wait peers' h chunks", "mma issue + wait", "tmem ld + act + sync", "cell + fence + sync", "push issue", "global stores"]),
                         ("bwd", bwd, ["wait peers' dh blocks", "cell + dgates + fence + sync", "mma issue + wait (+ stores, prefetch)", "tmem ld + transpose + push"])):
    g = gates0.clone()
    for _ in range(3): fn(g.clone())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    gg = [g.clone() for _ in range(10)]
    e0.record()
    for i in range(10): fn(gg[i])
    e1.record(); torch.cuda.synchronize()
    print(name, "kernel us", 1e2 * e0.elapsed_time(e1))
    buf = torch.zeros(8 * T, dtype=torch.int64, device=DEV)
    L.qeb_debug_set_timeline(buf.data_ptr()); fn(g.clone()); torch.cuda.synchronize(); L.qeb_debug_set_timeline(None)
    t = buf.cpu().numpy().reshape(T, 8)
    n = len(labels)
    d = np.diff(t[:, :n + 1], axis=1)[5:T - 1]     # skip the first steps (cold) and the last (no exchange)
    for lab, col in zip(labels, d.T):
        print(f"   {lab:28s} mean {col.mean():8.0f} cyc")
    print(f"   {'step total':28s} mean {np.diff(t[:, 0])[5:T - 1].mean():8.0f} cyc")
