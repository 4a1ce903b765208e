import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import qeb_b200
from qeb_b200.mirror.models.model_crnn import CRNN
from qeb_b200.mirror.models.model_unet import UNet
dev = "cuda"
torch.manual_seed(1)
which = sys.argv[1]
if which == "crnn":
    m = CRNN(95, False).to(dev); m.train()
    x = torch.rand(3, 1, 32, 64, device=dev, requires_grad=True)
    lp = m(x); lp.backward(torch.randn_like(lp))
    m.convo.batchnorm1.eval(); m.convo.batchnorm2.eval(); m.zero_grad()
    lp = m(x); lp.backward(torch.randn_like(lp))
else:
    m = UNet().to(dev); m.train()
    x = torch.rand(2, 1, 32, 64, device=dev, requires_grad=True)
    y = m(x); y.backward(torch.randn_like(y))
torch.cuda.synchronize()
print("done", float(sum(p.grad.abs().sum() for p in m.parameters())))
