"""Per-(kernel, size) time table of one training step (profiling detail mode)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
import qeb_b200
from qeb_b200 import _lib
from qeb_b200.mirror import ctc as qctc, train_ops
from qeb_b200.mirror.models.model_crnn import CRNN
from qeb_b200.mirror.models.model_unet import UNet
from qeb_b200.mirror.utils import set_bn_eval
dev = "cuda"
torch.manual_seed(42)
prep, crnn = UNet().to(dev), CRNN(95, False).to(dev)
opt = train_ops.Adam(prep.parameters(), lr=5e-5)
x, labels = bench.synth_batch(64, 7); x = x.to(dev)
c2i = {c: i for i, c in enumerate(bench.CHAR_SET)}
y, ys = bench.encode(labels, c2i)
packed = qctc.pack_targets(y, torch.tensor([31] * 64, dtype=torch.int32), ys, dev)
loss_fn = qctc.CTCLoss()
def step():
    prep.train(); crnn.train(); crnn.apply(set_bn_eval)
    prep.zero_grad(set_to_none=True); crnn.zero_grad(set_to_none=True)
    img = prep(x); scores = crnn(img)
    loss = loss_fn(scores, packed) + train_ops.mse_to_ones(img)
    loss.backward(); opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
_lib.load().qeb_prof_enable(2)
_lib.prof_report()
N = 5
for _ in range(N): step()
torch.cuda.synchronize()
rep = _lib.prof_report()
tot = sum(v["ms"] for v in rep.values())
filt = sys.argv[1] if len(sys.argv) > 1 else ""
for k, v in sorted(rep.items(), key=lambda kv: -kv[1]["ms"]):
    if filt in k:
        print(f"{k:48s} {v['launches']/N:5.1f}x {1e3*v['ms']/v['launches']:8.1f} us/launch {100*v['ms']/tot:5.1f}%  {v['flops']/v['ms']/1e9 if v['ms'] else 0:7.1f} TF/s {v['bytes']/v['ms']/1e6 if v['ms'] else 0:7.0f} GB/s")
