"""north_star gate: greedy-decoded strings identical to the fp32 reference graph on >= 99.9 % of patches.

The gate is ill-conditioned on random-init weights (near-flat log-probs), so the surrogate is first
TRAINED - through the qeb path itself (jitter-free phase-A steps: CRNN train mode -> CTC -> backward -> Adam, all
sm_100a kernels) - on procedurally rendered glyph strips until it reads them, and the comparison is made on 4096 fresh
patches in eval mode against oracle/nn_oracle.py (same weights, torch fp32, TF32 disabled). This also is the
end-to-end proof that the backward pass trains: accuracy against the true labels goes from 0 to > 90 %.
"""
import copy

import pytest
import torch

from oracle import nn_oracle

pytestmark = pytest.mark.gpu
DEV = "cuda"


def glyph_bank(n_cls, seed=0):
    g = torch.Generator().manual_seed(seed)
    pat = (torch.rand(n_cls, 10, 6, generator=g) < 0.45).float()
    return torch.nn.functional.interpolate(pat[:, None], scale_factor=2, mode="nearest")[:, 0]   # (n_cls, 20, 12)


def render(bank, n, gen, max_len=8):
    """White 32x128 strips with 1..8 dark glyphs at a 14-pixel pitch (3.5 CTC steps per glyph) + sensor noise."""
    x = torch.ones(n, 1, 32, 128)
    lens = torch.randint(1, max_len + 1, (n,), generator=gen)
    labels = []
    for i in range(n):
        cls = torch.randint(1, bank.shape[0] + 1, (int(lens[i]),), generator=gen)
        for j, c in enumerate(cls.tolist()):
            x0 = 4 + 14 * j + int(torch.randint(0, 2, (1,), generator=gen))
            x[i, 0, 6:26, x0:x0 + 12] -= 0.9 * bank[c - 1]
        labels.append(cls.to(torch.int32))
    return (x + 0.03 * torch.randn(x.shape, generator=gen)).clamp_(0, 1), labels


@pytest.mark.timeout(300)
def test_trained_surrogate_decode_parity():
    import qeb_b200  # noqa: F401
    from qeb_b200.mirror import ctc as qctc, train_ops, utils
    from qeb_b200.mirror.models.model_crnn import CRNN

    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    bank = glyph_bank(94)
    gen = torch.Generator().manual_seed(1)
    torch.manual_seed(0)
    m = CRNN(95, False).to(DEV)
    m.train()
    m.register_backward_hook(m.backward_hook)
    opt = train_ops.Adam(m.parameters(), lr=5e-4)
    loss_fn = qctc.CTCLoss()
    il = torch.full((64,), 31, dtype=torch.int32)
    first = last = None
    for it in range(2600):
        if it == 2200:   # settle: at lr 5e-4 the loss still spikes and a spike at the last step decides the eval accuracy
            for grp in opt.param_groups:
                grp["lr"] = 1e-4
        x, labels = render(bank, 64, gen)
        y = torch.cat(labels)
        ylen = torch.tensor([len(l) for l in labels], dtype=torch.int32)
        m.zero_grad(set_to_none=True)
        loss = loss_fn(m(x.to(DEV)), y, il, ylen)
        loss.backward()
        opt.step()
        if it == 0:
            first = float(loss)
        last = float(loss) if it >= 2590 else last
    assert last < 0.2 * first
    m.eval()
    mr = copy.deepcopy(m)
    same = correct = total = 0
    with torch.no_grad():
        for _ in range(16):
            x, labels = render(bank, 256, gen)
            lp, lpr = m(x.to(DEV)), nn_oracle.crnn_forward(mr, x.to(DEV))
            ca, la = utils.decode_batch(lp)
            cb, lb = utils.decode_batch(lpr.contiguous())
            same += int(((ca == cb).all(dim=1) & (la == lb)).sum())
            total += 256
            ca, la = ca.cpu(), la.cpu()
            correct += sum(int(int(la[i]) == len(l) and torch.equal(ca[i, :len(l)], l)) for i, l in enumerate(labels))
    print(f"trained surrogate: {same}/{total} strings identical to the fp32 graph, {correct}/{total} read correctly")
    assert correct >= 0.9 * total
    assert same >= 0.999 * total        # north_star: >= 99.9 % identical greedy decodes
