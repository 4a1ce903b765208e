"""Dev check of the tcgen05 conv kernels against torch fp32 (run on the GPU box)."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
import qeb_b200
from qeb_b200 import _lib
L = _lib.load()
P, I = ctypes.c_void_p, ctypes.c_int
L.qeb_pack_weight.argtypes = [P, P, I, I, I, I, I, P]
L.qeb_conv_fprop_tc.argtypes = [P, I, I, I, I, I, P, I, I, I, I, I, I, I, P, I, P, I, I, I, I, I, P]
L.qeb_conv_wgrad_tc.argtypes = [P, I, I, I, I, P, I, I, I, I, I, I, I, I, I, P, I, I, P]
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
st = torch.cuda.current_stream().cuda_stream

def chk(rc, what):
    if rc != 0:
        print("FAIL", what, rc, L.qeb_last_error().decode()); return False
    return True

def rel(a, b):
    return float((a - b).norm() / b.norm().clamp_min(1e-30))

def fprop(N, H, W, Cin, Cout, kh, kw, ph, pw, relu=1, bn=0):
    x = torch.randn(N, Cin, H, W, device="cuda")
    w = torch.randn(Cout, Cin, kh, kw, device="cuda") / (Cin * kh * kw) ** 0.5
    b = torch.randn(Cout, device="cuda")
    ref = F.conv2d(x, w, b, padding=(ph, pw))
    if relu: ref = ref.relu()
    Ho, Wo = ref.shape[2], ref.shape[3]
    xn = x.permute(0, 2, 3, 1).contiguous()
    wp = torch.empty(Cout, kh * kw, Cin, device="cuda")
    chk(L.qeb_pack_weight(w.data_ptr(), wp.data_ptr(), Cout, Cin, kh, kw, 0, st), "pack")
    out = torch.full((N, Ho, Wo, Cout), float("nan"), device="cuda")
    ok = chk(L.qeb_conv_fprop_tc(xn.data_ptr(), N, H, W, Cin, Cin, wp.data_ptr(), Cout, kh, kw, ph, pw, Ho, Wo, b.data_ptr(), relu,
                            out.data_ptr(), Cout, 0, 0, 0, bn, st), "fprop")
    torch.cuda.synchronize()
    e = rel(out.permute(0, 3, 1, 2), ref)
    print(f"fprop N{N} {H}x{W} {Cin}->{Cout} k{kh}x{kw} bn{bn}: rel {e:.2e} nan {int(torch.isnan(out).sum())}", "OK" if ok and e < 3e-3 else "BAD")
    return xn, w, out

def wgrad(N, H, W, Cin, Cout, kh, kw, ph, pw, bn=0):
    x = torch.randn(N, Cin, H, W, device="cuda")
    w = torch.zeros(Cout, Cin, kh, kw, device="cuda", requires_grad=True)
    y = F.conv2d(x, w, None, padding=(ph, pw))
    dy = torch.randn_like(y)
    y.backward(dy)
    Ho, Wo = y.shape[2], y.shape[3]
    xn = x.permute(0, 2, 3, 1).contiguous()
    dyn = dy.permute(0, 2, 3, 1).contiguous()
    dw = torch.zeros(Cout, Cin, kh, kw, device="cuda")
    ok = chk(L.qeb_conv_wgrad_tc(xn.data_ptr(), Cin, Cin, H, W, dyn.data_ptr(), Cout, Cout, N, Ho, Wo, kh, kw, ph, pw, dw.data_ptr(), 0, bn, st), "wgrad")
    torch.cuda.synchronize()
    e = rel(dw, w.grad)
    print(f"wgrad N{N} {H}x{W} {Cin}->{Cout} k{kh}x{kw} bn{bn}: rel {e:.2e}", "OK" if ok and e < 3e-3 else "BAD")

def dgrad(N, H, W, Cin, Cout):
    x = torch.randn(N, Cin, H, W, device="cuda", requires_grad=True)
    w = torch.randn(Cout, Cin, 3, 3, device="cuda") / (Cin * 9) ** 0.5
    y = F.conv2d(x, w, None, padding=1)
    dy = torch.randn_like(y)
    y.backward(dy)
    dyn = dy.permute(0, 2, 3, 1).contiguous()
    wp = torch.empty(Cin, 9, Cout, device="cuda")
    chk(L.qeb_pack_weight(w.data_ptr(), wp.data_ptr(), Cout, Cin, 3, 3, 1, st), "pack")
    dx = torch.empty(N, H, W, Cin, device="cuda")
    ok = chk(L.qeb_conv_fprop_tc(dyn.data_ptr(), N, H, W, Cout, Cout, wp.data_ptr(), Cin, 3, 3, 1, 1, H, W, None, 0, dx.data_ptr(), Cin, 0, 0, 0, 0, st), "dgrad")
    torch.cuda.synchronize()
    e = rel(dx.permute(0, 3, 1, 2), x.grad)
    print(f"dgrad N{N} {H}x{W} {Cin}<-{Cout}: rel {e:.2e}", "OK" if ok and e < 3e-3 else "BAD")

def convT(N, H, W, Cin, Cout):
    x = torch.randn(N, Cin, H, W, device="cuda")
    w = torch.randn(Cin, Cout, 2, 2, device="cuda") / Cin ** 0.5
    b = torch.randn(Cout, device="cuda")
    ref = F.conv_transpose2d(x, w, b, stride=2)
    xn = x.permute(0, 2, 3, 1).contiguous()
    wp = torch.empty(4 * Cout, Cin, device="cuda")
    chk(L.qeb_pack_weight(w.data_ptr(), wp.data_ptr(), Cin, Cout, 2, 2, 2, st), "pack")
    b4 = b.repeat(4).contiguous()
    out = torch.full((N, 2 * H, 2 * W, Cout), float("nan"), device="cuda")
    ok = chk(L.qeb_conv_fprop_tc(xn.data_ptr(), N, H, W, Cin, Cin, wp.data_ptr(), 4 * Cout, 1, 1, 0, 0, H, W, b4.data_ptr(), 0, out.data_ptr(), Cout, 1, Cout, 0, 0, st), "convT")
    torch.cuda.synchronize()
    e = rel(out.permute(0, 3, 1, 2), ref)
    print(f"convT N{N} {H}x{W} {Cin}->{Cout}: rel {e:.2e} nan {int(torch.isnan(out).sum())}", "OK" if ok and e < 3e-3 else "BAD")

which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "fprop"):
    fprop(2, 8, 32, 32, 32, 1, 1, 0, 0, relu=0, bn=32)     # plain small GEMM-like
    fprop(4, 8, 32, 128, 256, 3, 3, 1, 1)                   # CRNN conv3
    fprop(8, 4, 32, 512, 512, 3, 3, 1, 1)                   # CRNN conv6
    fprop(8, 2, 32, 512, 512, 2, 2, 0, 0, relu=0)           # CRNN conv7
    fprop(2, 32, 128, 32, 32, 3, 3, 1, 1)                   # UNet enc1conv2
    fprop(1, 400, 512, 32, 32, 3, 3, 1, 1, bn=32)           # UNet document
    fprop(1, 1, 1984, 512, 2048, 1, 1, 0, 0, relu=0)        # LSTM input GEMM
    fprop(1, 1, 1984, 512, 95, 1, 1, 0, 0, relu=0)          # Linear (N=95) -- out stride 95 unaligned
if which in ("all", "dgrad"):
    dgrad(4, 8, 32, 128, 256)
    dgrad(2, 32, 128, 32, 64)
if which in ("all", "convT"):
    convT(4, 4, 16, 256, 128)
    convT(2, 16, 64, 64, 32)
if which in ("all", "wgrad"):
    wgrad(4, 8, 32, 128, 256, 3, 3, 1, 1)
    wgrad(8, 4, 32, 512, 512, 3, 3, 1, 1)
    wgrad(2, 32, 128, 32, 32, 3, 3, 1, 1)
    wgrad(8, 2, 32, 512, 512, 2, 2, 0, 0)
    wgrad(1, 1, 1984, 512, 2048, 1, 1, 0, 0)
print("launches", L.qeb_launch_count())
