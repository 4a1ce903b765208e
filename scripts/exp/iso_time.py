"""Isolated (no side-stream contention, L2 flushed) timings of the up-convolution input gradient and the weight-gradient kernel
at the step's shapes. The weight-gradient call goes through the C ABI, i.e. into torch's (Cout, Cin, kh, kw) layout, whose
reduction epilogue is 3.4x slower than the packed layout the engines use (profiles/r1_notes.md 18)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import qeb_b200
from qeb_b200 import _lib
class q: lib = _lib
dev = "cuda"
st = lambda: torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def timeit(fn, n=20):
    for _ in range(3): fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]

B = 64
print("convT dgrad (dy cstride 2C as in the concat gradient)")
for (h, w, C) in [(16, 64, 32), (8, 32, 64), (4, 16, 128), (2, 8, 256)]:   # dx geometry, dx.c = 2C, dy.c = C at 2h x 2w
    dy = torch.randn(B, 2 * h, 2 * w, 2 * C, device=dev)
    wd = torch.randn(2 * C, 4 * C, device=dev)
    dx = torch.empty(B, h, w, 2 * C, device=dev)
    f = lambda: q.lib.call("qeb_convT2x2_dgrad_tc", dy.data_ptr(), B, h, w, C, 2 * C, wd.data_ptr(), 2 * C, dx.data_ptr(), 2 * C, st())
    t = timeit(f)
    gf = 2.0 * B * h * w * 2 * C * 4 * C / 1e9
    mb = 4.0 * (B * 4 * h * w * C + B * h * w * 2 * C) / 1e6
    print(f"  dx {h}x{w}x{2*C}: {t:7.1f} us  {gf/t*1e3:7.1f} TF/s  {mb/t*1e3:7.0f} GB/s")
print("conv wgrad 3x3")
for (h, w, ci, co) in [(32, 128, 32, 32), (32, 128, 64, 32), (16, 64, 32, 64), (16, 64, 64, 64), (16, 64, 128, 64), (8, 32, 64, 128), (8, 32, 128, 128),
                       (8, 32, 256, 128), (4, 16, 256, 256), (2, 8, 512, 512), (16, 64, 64, 128), (8, 32, 128, 256), (8, 32, 256, 256), (4, 32, 256, 512), (4, 32, 512, 512)]:
    x = torch.randn(B, h, w, ci, device=dev)
    dy = torch.randn(B, h, w, co, device=dev)
    dw = torch.zeros(co, 9, ci, device=dev)
    f = lambda: q.lib.call("qeb_conv_wgrad_tc", x.data_ptr(), ci, ci, h, w, dy.data_ptr(), co, co, B, 3, 3, 1, 1, dw.data_ptr(), st())
    t = timeit(f)
    gf = 2.0 * B * h * w * ci * co * 9 / 1e9
    mb = 4.0 * B * h * w * (ci + co) / 1e6
    print(f"  {h}x{w} {ci}->{co}: {t:7.1f} us  {gf/t*1e3:7.1f} TF/s  {mb/t*1e3:7.0f} GB/s (operands once)")
