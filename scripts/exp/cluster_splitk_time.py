"""Same-process timing of the deep-layer contractions (few pixel tiles, long K) of conv_fprop_tc_kernel. Run twice:
QEB_TC_CLUSTER=1 (cluster split-K, DSMEM reduce-scatter) and QEB_TC_CLUSTER=0 (global-reduction split-K for plain epilogues,
narrowed N tiles otherwise). L2-warm, median of 20."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from qeb_b200 import _lib

DEV = "cuda"
st = lambda: torch.cuda.current_stream().cuda_stream
shapes = [("256->256 @4x16", 64, 4, 16, 256, 256, 3, 1), ("128->256 @4x16", 64, 4, 16, 128, 256, 3, 1), ("512->256 @4x16", 64, 4, 16, 512, 256, 3, 1),
          ("256->512 @2x8", 64, 2, 8, 256, 512, 3, 1), ("512->512 @2x8", 64, 2, 8, 512, 512, 3, 1), ("conv7 512->512 k2 @2x32", 64, 2, 32, 512, 512, 2, 0)]
for name, N, H, W, Cin, Cout, k, p in shapes:
    x = torch.randn(N, H, W, Cin, device=DEV)
    x16 = x.half()
    w16 = (torch.randn(Cout, k * k, Cin, device=DEV) / (Cin * k * k) ** 0.5).half()
    Ho, Wo = H + 2 * p - k + 1, W + 2 * p - k + 1
    out = torch.empty(N, Ho, Wo, Cout, device=DEV)
    bias = torch.randn(Cout, device=DEV)
    res = []
    for b in (None, bias):
        def run():
            _lib.call("qeb_conv_fprop_tc16", x16.data_ptr(), N, H, W, Cin, Cin, w16.data_ptr(), Cout, k, k, p, p, None if b is None else b.data_ptr(), None,
                      0 if b is None else 1, out.data_ptr(), Cout, None, st())
        for _ in range(3):
            run()
        ts = []
        for _ in range(20):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        res.append(sorted(ts)[len(ts) // 2])
    fl = 2.0 * N * Ho * Wo * Cin * Cout * k * k
    print(f"{name:26s} plain epilogue {res[0]:6.1f} us ({fl / res[0] / 1e6:5.0f} TF/s)   bias+ReLU epilogue {res[1]:6.1f} us ({fl / res[1] / 1e6:5.0f} TF/s)")
