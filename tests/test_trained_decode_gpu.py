"""north_star gate: greedy-decoded strings identical to the fp32 reference graph on >= 99.9 % of patches.

The gate is ill-conditioned on random-init weights (near-flat log-probs), so the surrogate is first
TRAINED on procedurally rendered glyph strips until it reads them, and the comparison is made on fresh patches in eval
mode against oracle/nn_oracle.py (same weights, torch fp32, TF32 disabled):
  - `reference`: trained by the REFERENCE path (the oracle graph on torch eager / cuDNN fp32, torch CTCLoss and Adam -
    none of the qeb kernels), gate on 10,240 patches (SURVEY.md H1 asks for >= 10 k);
  - `qeb`: trained through the qeb path itself (CRNN train mode -> CTC -> backward -> Adam, all sm_100a kernels), gate on
    4,096 patches - also the end-to-end proof that the backward pass trains (accuracy goes from 0 to > 90 %).
The 35 MB of trained weights are not committed as a fixture (the recipe is seeded: same weights on every run); the observed
counts are written to gpurun_out/decode_gate.json (copied to profiles/ and quoted by bench.py's `parity` block).
"""
import copy
import json
import os

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


@pytest.mark.timeout(900)
@pytest.mark.parametrize("trainer,n_eval", [("reference", 10240), ("qeb", 4096)])
def test_trained_surrogate_decode_parity(trainer, n_eval):
    """The REFERENCE-path training (torch eager + cuDNN, not under test here) is not bit-reproducible run to run, and at these
    learning rates one run in a few ends in a loss spike: such a run is repeated once with the next seed before the gate is
    applied. The gate itself (identical greedy decodes on >= 99.9 % of the patches) is never relaxed."""
    try:
        _decode_parity(trainer, n_eval, 0)
    except _NotConverged:
        if trainer != "reference":
            raise
        _decode_parity(trainer, n_eval, 1)


class _NotConverged(AssertionError):
    pass


def _decode_parity(trainer, n_eval, attempt):
    import qeb_b200  # noqa: F401
    from qeb_b200.mirror import ctc as qctc, train_ops, utils
    from qeb_b200.mirror.models.model_crnn import CRNN

    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    bank = glyph_bank(94)
    gen = torch.Generator().manual_seed(1 + attempt)
    torch.manual_seed(attempt)
    m = CRNN(95, False).to(DEV)
    m.train()
    if trainer == "qeb":
        m.register_backward_hook(m.backward_hook)
        opt = train_ops.Adam(m.parameters(), lr=5e-4)
        loss_fn = qctc.CTCLoss()
        forward = m
    else:   # the reference's graph and library ops only
        opt = torch.optim.Adam(m.parameters(), lr=5e-4)
        loss_fn = torch.nn.CTCLoss()
        forward = lambda x: nn_oracle.crnn_forward(m, x)
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
        if trainer == "qeb":
            loss = loss_fn(forward(x.to(DEV)), y, il, ylen)
        else:
            loss = loss_fn(forward(x.to(DEV)), y.to(DEV), il.to(DEV), ylen.to(DEV))
        loss.backward()
        opt.step()
        if it == 0:
            first = float(loss)
        last = float(loss) if it >= 2590 else last
    if not last < 0.2 * first:
        raise _NotConverged(f"training by the {trainer} path did not converge: loss {first:.3f} -> {last:.3f}")
    m.eval()
    mr = copy.deepcopy(m)
    same = correct = total = 0
    with torch.no_grad():
        for _ in range(n_eval // 256):
            x, labels = render(bank, 256, gen)
            lp, lpr = m(x.to(DEV)), nn_oracle.crnn_forward(mr, x.to(DEV))
            ca, la = utils.decode_batch(lp)                     # rides the head's arg-max path
            cb, lb = utils.decode_batch(lpr.contiguous())
            same += int(((ca == cb).all(dim=1) & (la == lb)).sum())
            total += 256
            ca, la = ca.cpu(), la.cpu()
            correct += sum(int(int(la[i]) == len(l) and torch.equal(ca[i, :len(l)], l)) for i, l in enumerate(labels))
    print(f"surrogate trained by the {trainer} path: {same}/{total} strings identical to the fp32 graph, {correct}/{total} read correctly")
    try:   # observed counts for profiles/ (never fails the test)
        out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "decode_gate.json")
        rec = json.load(open(out)) if os.path.exists(out) else {}
        rec[trainer] = {"patches": total, "identical_to_fp32_graph": same, "read_correctly": correct, "required_identical": 0.999}
        os.makedirs(os.path.dirname(out), exist_ok=True)
        json.dump(rec, open(out, "w"), indent=1)
    except OSError:
        pass
    if not correct >= 0.9 * total:
        raise _NotConverged(f"the surrogate trained by the {trainer} path reads {correct}/{total} patches")
    assert same >= 0.999 * total        # north_star: >= 99.9 % identical greedy decodes
