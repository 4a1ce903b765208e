"""Per-CTA timeline of the tensor-core fprop kernel on one layer shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import qeb_b200
from qeb_b200 import _lib
L = _lib.load()
N, H, W, Cin, Cout = [int(v) for v in (sys.argv[1:6] if len(sys.argv) > 5 else (64, 32, 128, 32, 32))]
x = torch.randn(N, H, W, Cin, device="cuda"); w = torch.randn(Cout, Cin, 3, 3, device="cuda")
wp = torch.empty(Cout, 9, Cin, device="cuda"); out = torch.empty(N, H, W, Cout, device="cuda")
st = torch.cuda.current_stream().cuda_stream
_lib.call("qeb_pack_weight", w.data_ptr(), wp.data_ptr(), Cout, Cin, 3, 3, 0, st)
def run():
    _lib.call("qeb_conv_fprop_tc", x.data_ptr(), N, H, W, Cin, Cin, wp.data_ptr(), Cout, 3, 3, 1, 1, None, None, 0, out.data_ptr(), Cout, 0, st)
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record(); torch.cuda.synchronize()
print("kernel time us", 1e3 * e0.elapsed_time(e1))
buf = torch.zeros(16 * 8192, dtype=torch.int64, device="cuda")
L.qeb_debug_set_timeline(buf.data_ptr()); run(); torch.cuda.synchronize(); L.qeb_debug_set_timeline(None)
t = buf.cpu().numpy().reshape(-1, 16); t = t[t[:, 0] > 0]
print("CTAs", len(t))
names = ["setup", "first stage landed (from setup)", "first tile: accum ready (from first stage)", "first tile: epilogue",
         "  tmem ld of chunk 0", "  scale/bias/relu + transposition stores", "  row reads, mask, global stores (chunk 0)",
         "total CTA", "tiles per CTA", "cycles per tile"]
d = [t[:, 1] - t[:, 0], t[:, 2] - t[:, 1], t[:, 4] - t[:, 2], t[:, 5] - t[:, 4], t[:, 8] - t[:, 4], t[:, 9] - t[:, 8], t[:, 10] - t[:, 9], t[:, 6] - t[:, 0], t[:, 3], (t[:, 6] - t[:, 0]) / np.maximum(t[:, 3], 1)]
for n_, v in zip(names, d):
    print(f"{n_:44s} mean {v.mean():9.0f}  p10 {np.percentile(v,10):8.0f}  p90 {np.percentile(v,90):8.0f}")
# per-SM concurrency: CTAs per SM and the span of the kernel in cycles on one SM
sm = t[:, 7]; s0 = sm == sm[0]
print("CTAs on SM", int(sm[0]), ":", int(s0.sum()), "span cycles", int(t[s0, 6].max() - t[s0, 0].min()))
