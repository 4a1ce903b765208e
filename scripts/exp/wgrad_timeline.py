"""Per-CTA clock64 timeline of the weight-gradient kernel on one layer shape: python wgrad_timeline.py N H W Cin Cout."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import qeb_b200
from qeb_b200 import _lib
L = _lib.load()
N, H, W, Cin, Cout = [int(v) for v in (sys.argv[1:6] if len(sys.argv) > 5 else (64, 2, 8, 512, 512))]
KS = int(sys.argv[6]) if len(sys.argv) > 6 else 3   # 3: 3x3 pad 1 (torch gradient layout); 1: 1x1 (rows contiguous: vector reductions)
x = torch.randn(N, H, W, Cin, device="cuda"); dy = torch.randn(N, H, W, Cout, device="cuda")
dw = torch.zeros(Cout, KS * KS, Cin, device="cuda")
st = torch.cuda.current_stream().cuda_stream
F16 = os.environ.get("WG_F16", "0") == "1"   # fp16 MN-major operand shadows instead of tf32 reads
x16, dy16 = x.half(), dy.half()
def run():
    if F16:
        _lib.call("qeb_conv_wgrad_tc16", x.data_ptr(), x16.data_ptr(), Cin, Cin, H, W, dy.data_ptr(), dy16.data_ptr(), Cout, Cout, N, KS, KS,
                  KS // 2, KS // 2, None, dw.data_ptr(), st)
    else:
        _lib.call("qeb_conv_wgrad_tc", x.data_ptr(), Cin, Cin, H, W, dy.data_ptr(), Cout, Cout, N, KS, KS, KS // 2, KS // 2, dw.data_ptr(), st)
for _ in range(3): run()
torch.cuda.synchronize()
ts = []
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize(); ts.append(1e3 * e0.elapsed_time(e1))
print(f"shape N{N} {H}x{W} {Cin}->{Cout}: kernel time us (L2-warm) median {sorted(ts)[5]:.1f}")
buf = torch.zeros(16 * 16384, dtype=torch.int64, device="cuda")
L.qeb_debug_set_timeline(buf.data_ptr()); run(); torch.cuda.synchronize(); L.qeb_debug_set_timeline(None)
t = buf.cpu().numpy().reshape(-1, 16); t = t[t[:, 0] > 0]
print("CTAs", len(t), "k-steps per CTA", t[:, 3].mean(), "main-loop cycles per k-step", ((t[:, 4] - t[:, 2]) / np.maximum(t[:, 3], 1)).mean())
names = ["setup (barriers, TMEM alloc, PDL wait)", "first stage landed (from setup)", "main loop (first stage -> accumulator ready)", "epilogue (TMEM -> red.global.add)", "tail", "total CTA"]
d = [t[:, 1] - t[:, 0], t[:, 2] - t[:, 1], t[:, 4] - t[:, 2], t[:, 5] - t[:, 4], t[:, 6] - t[:, 5], t[:, 6] - t[:, 0]]
for n_, v in zip(names, d):
    print(f"{n_:48s} mean {v.mean():9.0f}  p10 {np.percentile(v,10):8.0f}  p90 {np.percentile(v,90):8.0f}")
sm = t[:, 7]
print("SMs used", len(np.unique(sm)), "max CTAs on one SM", np.bincount(sm.astype(int)).max(), "kernel span cycles", int(t[:, 6].max() - t[:, 0].min()))
