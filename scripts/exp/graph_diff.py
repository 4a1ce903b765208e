"""Experiment: run-to-run spread of the eager gradients vs the graph/eager difference, per parameter."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import qeb_b200
from qeb_b200.graphs import GraphedStep, StaticTargets
from qeb_b200.mirror import ctc as qctc, train_ops
from qeb_b200.mirror.models.model_crnn import CRNN
from qeb_b200.mirror.models.model_unet import UNet
from qeb_b200.mirror.utils import set_bn_eval
DEV = "cuda"
torch.manual_seed(3)
B, V = 16, 95
prep, crnn = UNet().to(DEV), CRNN(V, False).to(DEV)
prep.train(); crnn.train(); crnn.apply(set_bn_eval)
loss_fn = qctc.CTCLoss()
il = torch.full((B,), 31, dtype=torch.int32)
g = torch.Generator().manual_seed(11)
x = torch.rand(B, 1, 32, 128, generator=g).to(DEV)
tl = torch.randint(1, 12, (B,), generator=g, dtype=torch.int32)
y = torch.randint(1, V, (int(tl.sum()),), generator=g, dtype=torch.int32)
names = [n for n, _ in prep.named_parameters()] + ["crnn." + n for n, _ in crnn.named_parameters()]
params = list(prep.parameters()) + list(crnn.parameters())
def eager():
    prep.zero_grad(set_to_none=True); crnn.zero_grad(set_to_none=True)
    img = prep(x)
    loss = loss_fn(crnn(img), y, il, tl) + train_ops.mse_to_ones(img)
    loss.backward()
    torch.cuda.synchronize()
    return float(loss), [p.grad.clone() for p in params], img.detach().clone()
def diff(a, b, tag):
    worst = sorted(((float((u - v).norm() / v.norm().clamp_min(1e-30)), n) for u, v, n in zip(a, b, names)), reverse=True)[:6]
    print(tag, " ".join(f"{n}:{d:.1e}" for d, n in worst))
l1, g1, i1 = eager(); l2, g2, i2 = eager(); l3, g3, i3 = eager()
print("losses", l1, l2, l3, "img diff", float((i1 - i2).abs().max()))
diff(g1, g2, "eager1 vs eager2:"); diff(g2, g3, "eager2 vs eager3:")
tg = StaticTargets(B, 31, DEV).load(y, il, tl)
def fwd_bwd():
    img = prep(x)
    loss = loss_fn(crnn(img), tg) + train_ops.mse_to_ones(img)
    loss.backward()
    return loss
gs = GraphedStep(fwd_bwd, modules=[prep, crnn], warmup=2)
for i in range(3):
    l = gs(); torch.cuda.synchronize()
    gg = [p.grad.clone() for p in params]
    print("graph loss", float(l)); diff(gg, g1, f"graph{i} vs eager1:")
