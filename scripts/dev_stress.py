import os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import qeb_b200
from qeb_b200.mirror.models.model_crnn import CRNN
dev = "cuda"
torch.manual_seed(1); random.seed(1)
m = CRNN(95, False).to(dev); m.train()
names = [n for n, _ in m.named_parameters()]
fails = 0
for it in range(300):
    # poison the allocator's free lists with huge values
    junk = [torch.full((random.randint(1, 4000000),), 3e38, device=dev) for _ in range(random.randint(1, 6))]
    del junk
    B, W = random.choice([(3, 64), (5, 32), (2, 128), (7, 64)])
    x = torch.rand(B, 1, 32, W, device=dev, requires_grad=True)
    m.zero_grad(set_to_none=True)
    lp = m(x)
    lp.backward(torch.randn_like(lp))
    bad = []
    for n, p in m.named_parameters():
        big = (p.grad.abs() > 1e4) | ~torch.isfinite(p.grad)
        if big.any():
            idx = big.reshape(-1).nonzero().reshape(-1)
            bad.append((n, tuple(p.shape), int(big.sum()), idx[:8].tolist(), idx[-3:].tolist(), p.grad.reshape(-1)[idx[:4]].tolist()))
    if (x.grad.abs() > 1e4).any() or not torch.isfinite(x.grad).all(): bad.append(("dx",))
    if (lp.abs() > 1e4).any() or not torch.isfinite(lp).all(): bad.append(("lp",))
    if bad:
        fails += 1
        print("iter", it, "B,W", B, W, "BAD:", bad[:4])
        if fails >= 5: break
print("stress done, failures:", fails)
