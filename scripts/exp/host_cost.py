"""Experiment: host-side cost of issuing one step (time inside each C entry point vs Python/autograd around them)."""
import os, sys, time, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
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
acc = collections.defaultdict(float)
orig = _lib.call
def timed_call(name, *a):
    t = time.perf_counter(); orig(name, *a); acc[name] += time.perf_counter() - t
for m in (sys.modules[n] for n in list(sys.modules) if n.startswith("qeb_b200")):
    if getattr(m, "_lib", None) is _lib: pass
_lib.call = timed_call
def step():
    prep.train(); crnn.train(); crnn.apply(set_bn_eval)
    prep.zero_grad(set_to_none=True); crnn.zero_grad(set_to_none=True)
    img = prep(x); scores = crnn(img)
    loss = loss_fn(scores, packed) + train_ops.mse_to_ones(img)
    loss.backward(); opt.step()
for _ in range(3): step()
torch.cuda.synchronize(); acc.clear()
N = 20
t0 = time.perf_counter()
for _ in range(N):
    step()
    torch.cuda.synchronize()      # GPU idle at every step start: pure issue cost, no back-pressure
tot = time.perf_counter() - t0
print(f"step wall (sync each) {1e3*tot/N:.3f} ms")
for k, v in sorted(acc.items(), key=lambda kv: -kv[1]):
    print(f"  {k:28s} {1e3*v/N:7.3f} ms/step")
print(f"  sum in C calls               {1e3*sum(acc.values())/N:7.3f} ms/step")
