"""The two fusions north_star names, against the unfused kernels and the oracle:
  - log-softmax (+ the per-frame arg-max of pred_to_string) in the epilogue of the CRNN head's Linear GEMM
    (models/model_crnn.py:20, utils.py:78-89), with the log-softmax autograd node - what CRNN.backward_hook sees - kept;
  - the Gaussian jitter of AddGaussianNoice (transform_helper.py:33-45) inside conv1's input load, the noisy batch still
    materialised for the OCR hand-off."""
import copy

import numpy as np
import pytest
import torch

from conftest import CHAR_SET
from oracle import nn_oracle, pyoracle as po

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def q():
    import qeb_b200  # noqa: F401
    from qeb_b200.mirror import ctc, transform_helper, utils
    from qeb_b200.mirror.models.model_crnn import CRNN

    class Q:
        pass

    Q.ctc, Q.th, Q.utils, Q.CRNN = ctc, transform_helper, utils, CRNN
    return Q


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


@pytest.mark.parametrize("B,W,V", [(64, 128, 95), (3, 64, 95), (5, 256, 63), (130, 32, 96)])
def test_fused_head_matches_separate_log_softmax_and_decode(q, B, W, V):
    torch.manual_seed(B)
    m = q.CRNN(V, False).to(DEV)
    m.train(); m.apply(q.utils.set_bn_eval)                  # frozen BatchNorm: the forward has no atomics, runs are bit-equal
    x = torch.rand(B, 1, 32, W, device=DEV)
    lp = m(x)                                                # fused head
    lp_ref = q.ctc.log_softmax(m.forward_logits(x))          # logits from the GEMM + the separate log-softmax kernel
    assert lp.shape == (W // 4 - 1, B, V)
    assert float((lp - lp_ref).abs().max()) < 2e-5
    assert float((lp.exp().sum(2) - 1).abs().max()) < 1e-5
    path, ver = lp._qeb_path
    assert ver == lp._version and path.dtype == torch.int32 and path.shape == lp.shape[:2]
    assert torch.equal(path.long(), lp.detach().argmax(2))   # first maximal index of the STORED log-probs
    # decode riding the head: collapse of the path == full greedy decode of the scores == the oracle
    codes, lens = q.utils.decode_batch(lp)
    plain = lp.detach().clone()                              # a copy carries no path: the (T,B,V) kernel runs
    assert not hasattr(plain, "_qeb_path")
    codes2, lens2 = q.utils.decode_batch(plain)
    assert torch.equal(codes, codes2) and torch.equal(lens, lens2)
    oc, ol = po.greedy_decode(lp.detach().cpu().numpy())
    assert np.array_equal(codes.cpu().numpy(), oc) and np.array_equal(lens.cpu().numpy(), ol)
    if V == 95:
        i2c = {i: c for i, c in enumerate(CHAR_SET)}
        assert q.utils.pred_to_string(lp, None, i2c) == po.pred_to_string(lp.detach().cpu().numpy(), i2c)
    # an in-place edit invalidates the attached path (the scores are decoded again)
    lp2 = m(x).detach()
    lp2._qeb_path = (path, lp2._version)
    lp2[0, 0, :] = -50.0
    lp2[0, 0, 7] = 0.0
    c3, _ = q.utils.decode_batch(lp2)
    assert int(c3[0, 0]) == 7


def test_fused_head_backward_and_nan_hook(q):
    torch.manual_seed(1)
    B, T = 16, 31
    m = q.CRNN(95, False).to(DEV)
    m.register_backward_hook(m.backward_hook)
    m.train(); m.apply(q.utils.set_bn_eval)
    x = torch.rand(B, 1, 32, 128, device=DEV)
    ylen = torch.randint(1, 10, (B,), dtype=torch.int32)
    ylen[3] = 40                                             # longer than T: infeasible, loss inf, NaN gradient rows
    y = torch.randint(1, 95, (int(ylen.sum()),), dtype=torch.int32)
    il = torch.full((B,), T, dtype=torch.int32)
    loss = q.ctc.CTCLoss()(m(x), y, il, ylen)
    assert torch.isinf(loss)
    loss.backward()
    g_fused = [p.grad.clone() for p in m.parameters()]
    assert all(torch.isfinite(g).all() for g in g_fused)     # the hook zeroed the NaNs at the logits
    m.zero_grad(set_to_none=True)
    logits = m.forward_logits(x)                             # outside m(...): the module hook is not on this graph,
    logits.register_hook(lambda g: torch.where(g != g, torch.zeros_like(g), g))   # so restate it on the logits' gradient
    loss2 = q.ctc.CTCLoss()(q.ctc.log_softmax(logits), y, il, ylen)
    loss2.backward()
    for a, p in zip(g_fused, m.parameters()):
        assert rel(a, p.grad) < 2e-4 or float((a - p.grad).abs().max()) < 1e-6


@pytest.mark.parametrize("bn_train", [False, True])
def test_jitter_fused_into_conv1_matches_standalone_kernel(q, bn_train):
    torch.manual_seed(5)
    B = 64
    m = q.CRNN(95, False).to(DEV)
    m.train()
    if not bn_train:
        m.apply(q.utils.set_bn_eval)
    mr = copy.deepcopy(m)
    x = torch.rand(B, 1, 32, 128, device=DEV)
    sig = (torch.randint(0, 6, (B,)).double() / 100 + 1e-13).float()
    lp, noisy, noise = m.forward_jittered(x, sig, mean=0.01, noise_coef=1, seed=77, return_noise=True)
    want, want_noise = q.th.jitter_batch(x, sig, mean=0.01, noise_coef=1, seed=77, return_noise=True)
    assert torch.equal(noisy, want) and torch.equal(noise, want_noise)          # same Philox stream, same arithmetic
    assert torch.equal(noisy, torch.clamp(x - noise, 0, 1))                      # transform_helper.py:40-41
    lp_ref = mr(want)
    if bn_train:
        assert float((lp - lp_ref).abs().max()) < 1e-4                           # batch statistics are summed with atomics
    else:
        assert torch.equal(lp, lp_ref)                                           # conv1 out of shared memory: same FMA order
    ylen = torch.randint(1, 12, (B,), dtype=torch.int32)
    y = torch.randint(1, 95, (int(ylen.sum()),), dtype=torch.int32)
    il = torch.full((B,), 31, dtype=torch.int32)
    q.ctc.CTCLoss()(lp, y, il, ylen).backward()
    q.ctc.CTCLoss()(lp_ref, y, il, ylen).backward()
    for (n, a), b in zip(m.named_parameters(), mr.parameters()):
        if bn_train and n in ("convo.conv5.bias", "convo.conv6.bias"):
            continue                                                             # analytically zero: rounding noise only
        assert rel(a.grad, b.grad) < 5e-3, (n, rel(a.grad, b.grad))
    # oracle: the reference graph on the materialised noisy batch
    lpo = nn_oracle.crnn_forward(copy.deepcopy(mr), noisy)
    assert float((lp.detach() - lpo).abs().max()) < 2e-3
    # the helper with the reference's noiser object; x needs no gradient and gets none
    xg = x.clone().requires_grad_(True)
    out = q.th.crnn_on_noised(m, xg, q.th.AddGaussianNoice(std=5, is_stochastic=True, return_noise=True))
    assert len(out) == 3 and out[1].shape == x.shape and float(out[1].min()) >= 0 and float(out[1].max()) <= 1
    out[0].sum().backward()
    assert xg.grad is None


def test_jitter_fused_device_seed_draws_fresh_noise_per_launch(q):
    torch.manual_seed(6)
    m = q.CRNN(95, False).to(DEV)
    m.train(); m.apply(q.utils.set_bn_eval)
    x = torch.rand(8, 1, 32, 64, device=DEV)
    sig = torch.full((8,), 0.05, device=DEV)
    key = torch.tensor([10], dtype=torch.int64, device=DEV)
    buf = torch.empty_like(x)
    with torch.no_grad():
        _, a = m.forward_jittered(x, sig, seed_dev=key, out=buf)
        a = a.clone()
        key.add_(1)
        _, b = m.forward_jittered(x, sig, seed_dev=key, out=buf)
    assert b.data_ptr() == buf.data_ptr() and not torch.equal(a, b)
    assert torch.equal(a, q.th.jitter_batch(x, sig, seed=10)) and torch.equal(b, q.th.jitter_batch(x, sig, seed=11))
