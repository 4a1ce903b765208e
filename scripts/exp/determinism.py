"""Experiment: run-to-run reproducibility of the forward pieces (bitwise where no atomics are involved)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import qeb_b200
from qeb_b200 import _lib
from qeb_b200.mirror.models.model_crnn import CRNN
from qeb_b200.mirror.models.model_unet import UNet
DEV = "cuda"
st = torch.cuda.current_stream().cuda_stream
def conv(N, H, W, Cin, Cout, k=3, p=1):
    g = torch.Generator(device=DEV).manual_seed(1)
    x = torch.randn(N, H, W, Cin, device=DEV, generator=g); w = torch.randn(Cout, Cin, k, k, device=DEV, generator=g) * 0.05
    wp = torch.empty(Cout, k * k, Cin, device=DEV)
    _lib.call("qeb_pack_weight", w.data_ptr(), wp.data_ptr(), Cout, Cin, k, k, 0, st)
    outs = []
    for _ in range(3):
        out = torch.zeros(N, H, W, Cout, device=DEV)
        _lib.call("qeb_conv_fprop_tc", x.data_ptr(), N, H, W, Cin, Cin, wp.data_ptr(), Cout, k, k, p, p, None, None, 0, out.data_ptr(), Cout, 0, st)
        torch.cuda.synchronize(); outs.append(out)
    ref = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2).double(), w.double(), padding=p).permute(0, 2, 3, 1)
    print(f"conv {N}x{H}x{W} {Cin}->{Cout}: run-to-run max|d| {float((outs[0]-outs[1]).abs().max()):.2e} {float((outs[1]-outs[2]).abs().max()):.2e}"
          f"  vs fp64 rel {float((outs[0].double()-ref).norm()/ref.norm()):.2e}  out scale {float(ref.abs().mean()):.2e}")
for shp in [(16, 32, 128, 32, 32), (16, 16, 64, 64, 64), (16, 8, 32, 128, 128), (16, 4, 16, 256, 256), (16, 2, 8, 256, 512), (16, 2, 8, 512, 512), (64, 2, 8, 512, 512)]:
    conv(*shp)
torch.manual_seed(3)
x = torch.rand(16, 1, 32, 128, device=DEV)
for mode in ("eval", "train"):
    m = UNet().to(DEV); getattr(m, mode)()
    with torch.no_grad():
        ys = [m(x).clone() for _ in range(3)]
    torch.cuda.synchronize()
    print("unet", mode, "run-to-run max|d|", float((ys[0] - ys[1]).abs().max()), float((ys[1] - ys[2]).abs().max()))
c = CRNN(95, False).to(DEV).train()
with torch.no_grad():
    ys = [c(x).clone() for _ in range(3)]
print("crnn train run-to-run max|d|", float((ys[0] - ys[1]).abs().max()), float((ys[1] - ys[2]).abs().max()))
