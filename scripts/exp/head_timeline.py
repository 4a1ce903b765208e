"""clock64 timeline of the CRNN head GEMM (the last fprop launch of a forward), fused log-softmax epilogue vs plain."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import qeb_b200
from qeb_b200 import _lib
import qeb_b200.mirror.models.model_crnn as mc
from qeb_b200.mirror import utils as qutils
L = _lib.load()
torch.manual_seed(0)
m = mc.CRNN(95, False).cuda(); m.train(); m.apply(qutils.set_bn_eval)
x = torch.rand(64, 1, 32, 128, device="cuda")
for fused in (0, 1):
    mc._FUSED_HEAD = bool(fused)
    with torch.no_grad():
        for _ in range(3): m(x)
        buf = torch.zeros(16 * 8192, dtype=torch.int64, device="cuda")
        torch.cuda.synchronize()
        L.qeb_debug_set_timeline(buf.data_ptr()); m(x); torch.cuda.synchronize(); L.qeb_debug_set_timeline(None)
    n = 16 if fused else 48
    t = buf.cpu().numpy().reshape(-1, 16)[:n]
    print("fused", fused, "CTAs", n)
    for name, v in (("setup", t[:, 1] - t[:, 0]), ("first stage landed", t[:, 2] - t[:, 1]), ("accum ready", t[:, 4] - t[:, 2]),
                    ("epilogue", t[:, 5] - t[:, 4]), ("exit after epilogue", t[:, 6] - t[:, 5]), ("total", t[:, 6] - t[:, 0]),
                    ("kernel span", np.full(n, t[:, 6].max() - t[:, 0].min()))):
        print(f"  {name:22s} mean {v.mean():9.0f} min {v.min():9.0f} max {v.max():9.0f}")
