"""Train the surrogate with the qeb path on procedural glyph patches, then compare greedy decodes with the fp32 oracle."""
import copy, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import qeb_b200
from qeb_b200.mirror import ctc as qctc, train_ops, utils
from qeb_b200.mirror.models.model_crnn import CRNN
from oracle import nn_oracle
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
dev = "cuda"

def glyph_bank(n_cls, seed=0):
    g = torch.Generator().manual_seed(seed)
    pat = (torch.rand(n_cls, 10, 6, generator=g) < 0.45).float()
    return torch.nn.functional.interpolate(pat[:, None], scale_factor=2, mode="nearest")[:, 0]   # (n_cls, 20, 12)

def render(bank, n, gen, max_len=8):
    x = torch.ones(n, 1, 32, 128)
    lens = torch.randint(1, max_len + 1, (n,), generator=gen)
    labels = []
    for i in range(n):
        cls = torch.randint(1, bank.shape[0] + 1, (int(lens[i]),), generator=gen)
        for j, c in enumerate(cls.tolist()):
            x0 = 4 + 14 * j + int(torch.randint(0, 2, (1,), generator=gen))
            x[i, 0, 6:26, x0:x0 + 12] -= 0.9 * bank[c - 1]
        labels.append(cls.to(torch.int32))
    x = (x + 0.03 * torch.randn(x.shape, generator=gen)).clamp_(0, 1)
    return x, labels

bank = glyph_bank(94)
gen = torch.Generator().manual_seed(1)
torch.manual_seed(0)
m = CRNN(95, False).to(dev); m.train()
m.register_backward_hook(m.backward_hook)
opt = train_ops.Adam(m.parameters(), lr=float(sys.argv[2]) if len(sys.argv) > 2 else 5e-4)
loss_fn = qctc.CTCLoss()
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 600
t0 = time.time()
for it in range(steps):
    x, labels = render(bank, 64, gen)
    y = torch.cat(labels); ylen = torch.tensor([len(l) for l in labels], dtype=torch.int32)
    il = torch.full((64,), 31, dtype=torch.int32)
    m.zero_grad(set_to_none=True)
    loss = loss_fn(m(x.to(dev)), y, il, ylen)
    loss.backward(); opt.step()
    hist = globals().setdefault("hist", [])
    hist.append(loss.detach())
    if it % 100 == 0 or it == steps - 1:
        h = torch.stack(hist[-100:]).float()
        print(f"step {it} loss {float(loss):.4f} last-100 mean {float(h.mean()):.4f} max {float(h.max()):.4f} ({time.time()-t0:.1f}s)", flush=True)
m.eval(); mr = copy.deepcopy(m); mr.eval()
same = correct = total = 0
with torch.no_grad():
    for chunk in range(16):
        x, labels = render(bank, 256, gen)
        lp = m(x.to(dev)); lpr = nn_oracle.crnn_forward(mr, x.to(dev))
        ca, la = utils.decode_batch(lp); cb, lb = utils.decode_batch(lpr.contiguous())
        eq = ((ca == cb).all(dim=1) & (la == lb)).cpu()
        same += int(eq.sum()); total += 256
        ca, la = ca.cpu(), la.cpu()
        for i, l in enumerate(labels):
            correct += int(int(la[i]) == len(l) and torch.equal(ca[i, :len(l)], l))
print(f"trained-surrogate decode parity: {same}/{total} identical strings; accuracy vs labels {correct}/{total}; max |dlogp| {float((lp-lpr).abs().max()):.2e}")
