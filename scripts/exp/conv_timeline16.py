"""Per-CTA clock64 timeline of the tensor-core conv kernel with fp16 operands + bias + ReLU + fp16 shadow (the forward
configuration of the networks) and with tf32 operands, plain epilogue (the input-gradient configuration), per layer shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import qeb_b200
from qeb_b200 import _lib
L = _lib.load()
SHAPES = [(64, 32, 128, 32, 32), (64, 32, 128, 64, 32), (64, 16, 64, 64, 64), (64, 8, 32, 128, 128), (64, 4, 16, 256, 256),
          (64, 2, 8, 512, 512), (64, 16, 64, 64, 128), (64, 8, 32, 128, 256), (64, 4, 32, 256, 512), (64, 4, 32, 512, 512)]
if len(sys.argv) > 5:
    SHAPES = [tuple(int(v) for v in sys.argv[1:6])]
st = torch.cuda.current_stream().cuda_stream
for (N, H, W, Cin, Cout) in SHAPES:
    x = torch.randn(N, H, W, Cin, device="cuda"); w = torch.randn(Cout, Cin, 3, 3, device="cuda") * 0.05
    wp = torch.empty(Cout, 9, Cin, device="cuda"); out = torch.empty(N, H, W, Cout, device="cuda")
    _lib.call("qeb_pack_weight", w.data_ptr(), wp.data_ptr(), Cout, Cin, 3, 3, 0, st)
    x16, w16, o16 = x.half(), wp.half(), torch.empty(N, H, W, Cout, device="cuda", dtype=torch.half)
    bias = torch.randn(Cout, device="cuda")
    def run16():
        _lib.call("qeb_conv_fprop_tc16", x16.data_ptr(), N, H, W, Cin, Cin, w16.data_ptr(), Cout, 3, 3, 1, 1, bias.data_ptr(), None, 1,
                  out.data_ptr(), Cout, o16.data_ptr(), st)
    def run32():
        _lib.call("qeb_conv_fprop_tc", x.data_ptr(), N, H, W, Cin, Cin, wp.data_ptr(), Cout, 3, 3, 1, 1, None, None, 0, out.data_ptr(), Cout, 0, st)
    for name, run in (("f16+bias+relu+shadow", run16), ("tf32 plain", run32)):
        for _ in range(3): run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): run()
        e1.record(); torch.cuda.synchronize()
        us = 1e2 * e0.elapsed_time(e1)
        gf = 2.0 * N * H * W * Cout * 9 * Cin / 1e9
        buf = torch.zeros(16 * 8192, dtype=torch.int64, device="cuda")
        L.qeb_debug_set_timeline(buf.data_ptr()); run(); torch.cuda.synchronize(); L.qeb_debug_set_timeline(None)
        t = buf.cpu().numpy().reshape(-1, 16); t = t[t[:, 0] > 0]
        tot = t[:, 6] - t[:, 0]
        print(f"{N}x{H}x{W} {Cin}->{Cout} {name:22s} {us:7.1f} us {gf/us*1e-3*1e3:7.1f} TF/s | CTAs {len(t):4d} tiles/CTA {t[:,3].mean():5.1f} "
              f"setup {np.mean(t[:,1]-t[:,0]):6.0f} first-stage {np.mean(t[:,2]-t[:,1]):6.0f} tile0: mainloop {np.mean(t[:,4]-t[:,2]):6.0f} epilogue {np.mean(t[:,5]-t[:,4]):6.0f} "
              f"(ld {np.mean(t[:,8]-t[:,4]):5.0f} st.sh {np.mean(t[:,9]-t[:,8]):5.0f} out {np.mean(t[:,10]-t[:,9]):5.0f}) CTA total {tot.mean():7.0f} cyc/tile {np.mean(tot/np.maximum(t[:,3],1)):6.0f} | per CTA: epi-wait {t[:,11].mean():7.0f} epi-body {t[:,12].mean():7.0f} mma-wait-window {t[:,13].mean():7.0f} mma-wait-acc {t[:,14].mean():7.0f}")
