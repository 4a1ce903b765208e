"""Experiment: repeat one CRNN forward+CTC+backward on fixed data; report the spread of every gradient across repeats.
A rare corruption (race) shows up as an outlier far above the typical run-to-run spread."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import qeb_b200
from qeb_b200.mirror import ctc as qctc, train_ops
from qeb_b200.mirror.models.model_crnn import CRNN
DEV = "cuda"
torch.manual_seed(0)
m = CRNN(95, False).to(DEV).train()
m.register_backward_hook(m.backward_hook)
opt = train_ops.Adam(m.parameters(), lr=5e-4)
loss_fn = qctc.CTCLoss()
B = 64
g = torch.Generator().manual_seed(1)
x = torch.rand(B, 1, 32, 128, generator=g).to(DEV)
tl = torch.randint(1, 9, (B,), generator=g, dtype=torch.int32)
y = torch.randint(1, 95, (int(tl.sum()),), generator=g, dtype=torch.int32)
il = torch.full((B,), 31, dtype=torch.int32)
def grads():
    m.zero_grad(set_to_none=True)
    loss = loss_fn(m(x), y, il, tl); loss.backward()
    return float(loss), torch.cat([p.grad.reshape(-1) for p in m.parameters()])
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 0):   # optional pre-training steps
    grads(); opt.step()
l0, g0 = grads()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 500
devs, losses = [], []
for i in range(N):
    l, gi = grads()
    devs.append(float((gi - g0).norm() / g0.norm())); losses.append(l)
d = torch.tensor(devs)
print(f"loss {l0:.6f} (spread {max(losses)-min(losses):.2e}); rel grad deviation vs first run: median {float(d.median()):.2e} p99 {float(d.quantile(0.99)):.2e} max {float(d.max()):.2e}; non-finite {int((~torch.isfinite(d)).sum())}")
