"""Same-box timing of the weight-gradient kernel with tf32 reads against fp16 MN-major operand shadows, per layer shape of
the phase-B step (64 patches). L2 is flushed between repetitions by the operands themselves being re-read after a 256 MB write."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from qeb_b200 import _lib

DEV = "cuda"
st = lambda: torch.cuda.current_stream().cuda_stream
shapes = [  # name, N, H, W, Cin, Cout, k, p
    ("unet 32->32 @32x128", 64, 32, 128, 32, 32, 3, 1), ("unet 64->32 @32x128", 64, 32, 128, 64, 32, 3, 1),
    ("unet 64->64 @16x64", 64, 16, 64, 64, 64, 3, 1), ("unet 128->64 @16x64", 64, 16, 64, 128, 64, 3, 1),
    ("unet 128->128 @8x32", 64, 8, 32, 128, 128, 3, 1), ("unet 256->256 @4x16", 64, 4, 16, 256, 256, 3, 1),
    ("unet 512->512 @2x8", 64, 2, 8, 512, 512, 3, 1),
    ("crnn conv2 64->128 @16x64", 64, 16, 64, 64, 128, 3, 1), ("crnn conv3 128->256 @8x32", 64, 8, 32, 128, 256, 3, 1),
    ("crnn conv4 256->256 @8x32", 64, 8, 32, 256, 256, 3, 1), ("crnn conv5 256->512 @4x32", 64, 4, 32, 256, 512, 3, 1),
    ("crnn conv6 512->512 @4x32", 64, 4, 32, 512, 512, 3, 1), ("lstm W_ih 512->1024", 1, 1, 1984, 512, 1024, 1, 0),
]
flush = torch.empty(64 << 20, device=DEV)
for name, N, H, W, Cin, Cout, k, p in shapes:
    x = torch.randn(N, H, W, Cin, device=DEV)
    Ho, Wo = H + 2 * p - k + 1, W + 2 * p - k + 1
    dy = torch.randn(N, Ho, Wo, Cout, device=DEV)
    x16, dy16 = x.half(), dy.half()
    dw = torch.zeros(Cout, Cin, k, k, device=DEV)
    res = []
    for f16 in (0, 1):
        def run():
            if f16:
                _lib.call("qeb_conv_wgrad_tc16", x.data_ptr(), x16.data_ptr(), Cin, Cin, H, W, dy.data_ptr(), dy16.data_ptr(), Cout,
                          Cout, N, k, k, p, p, None, dw.data_ptr(), st())
            else:
                _lib.call("qeb_conv_wgrad_tc", x.data_ptr(), Cin, Cin, H, W, dy.data_ptr(), Cout, Cout, N, k, k, p, p, dw.data_ptr(), st())
        for _ in range(3):
            run()
        ts = []
        for _ in range(10):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        res.append(sorted(ts)[len(ts) // 2])
    fl = 2.0 * N * Ho * Wo * Cin * Cout * k * k
    print(f"{name:30s} tf32 {res[0]:7.1f} us ({fl / res[0] / 1e6:6.0f} TF/s)   fp16 {res[1]:7.1f} us ({fl / res[1] / 1e6:6.0f} TF/s)")
