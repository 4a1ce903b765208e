"""Parity of the tensor-core kernels, the LSTM recurrence and the two network engines (all through the C ABI / the
mirror modules) against (a) the golden fixtures written from the UNMODIFIED reference (tests/golden/crnn.npz, unet.npz,
oracle/gen_golden.py) and (b) the oracle port (oracle/nn_oracle.py: the reference graph on torch fp32, TF32 disabled).

Tolerances. The dense contractions run with TF32 operands and fp32 accumulation (10-bit mantissa, truncated by the
tensor core), so single kernels are checked at 3e-3 relative L2 (measured ~2-3e-4), network outputs at 2e-3 absolute on
log-probs / sigmoid outputs, and network gradients by relative L2 + cosine similarity with the bounds cuDNN's own TF32
path shows against fp32 on the same problems (measured side by side in scripts/dev_net.py). The LSTM
recurrence, CTC, BatchNorm, pooling and the direct convs are fp32 and are checked at 1e-5..1e-4.
"""
import copy
import ctypes
import json
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import CHAR_SET, GOLDEN, load_golden
from oracle import nn_oracle

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def q():
    import qeb_b200
    from qeb_b200 import _lib
    from qeb_b200.mirror import ctc, train_ops, utils
    from qeb_b200.mirror.models.model_crnn import CRNN
    from qeb_b200.mirror.models.model_unet import UNet

    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    assert _lib.load().qeb_check_device() == 0

    class Q:
        pass

    Q.lib, Q.ctc, Q.train_ops, Q.utils, Q.CRNN, Q.UNet = _lib, ctc, train_ops, utils, CRNN, UNet
    return Q


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def cos(a, b):
    a, b = a.detach().double().cpu().reshape(-1), b.detach().double().cpu().reshape(-1)
    return float((a @ b) / (a.norm() * b.norm()).clamp_min(1e-30))


def st():
    return torch.cuda.current_stream().cuda_stream


# ------------------------------------------------------------------------------------------------ tensor-core kernels
@pytest.mark.parametrize("N,H,W,Cin,Cout,k,p,relu", [
    (4, 8, 32, 128, 256, 3, 1, 1),      # CRNN conv3
    (8, 4, 32, 512, 512, 3, 1, 1),      # CRNN conv6
    (8, 2, 32, 512, 512, 2, 0, 0),      # CRNN conv7
    (2, 32, 128, 32, 32, 3, 1, 0),      # UNet enc1conv2
    (1, 48, 80, 64, 32, 3, 1, 1),       # UNet dec1conv1 on a non-power-of-two image
    (1, 1, 1984, 512, 2048, 1, 0, 0),   # LSTM input projection as a GEMM
    (1, 1, 1984, 512, 95, 1, 0, 0),     # Linear: N = 95 (tile overhang, unaligned rows)
    (3, 5, 7, 32, 64, 3, 1, 1),         # ragged: tiles larger than the image
    (16, 32, 128, 32, 32, 3, 1, 1),     # UNet level 1 at batch 16
    (12, 32, 128, 64, 32, 3, 1, 0),     # two channel slices, 65-wide N overhang none
    (40, 16, 64, 32, 64, 3, 1, 1),      # two N tiles, W = 64
    (70, 10, 40, 64, 64, 3, 1, 0),      # odd geometry (last tile of every image is partial)
    (64, 32, 128, 32, 32, 3, 1, 1),     # UNet level 1 at the bench batch: the WINDOW variant (>= 7 tiles of 8 x 16 per SM)
    (2, 200, 520, 64, 32, 3, 1, 0),     # window variant, partial tiles at the bottom edge, two channel-slice weight sets
])
def test_conv_fprop_dgrad_wgrad_tc(q, N, H, W, Cin, Cout, k, p, relu):
    g = torch.Generator(device=DEV).manual_seed(N * 1000 + Cin + Cout)
    x = torch.randn(N, Cin, H, W, device=DEV, generator=g, requires_grad=True)
    w = (torch.randn(Cout, Cin, k, k, device=DEV, generator=g) / (Cin * k * k) ** 0.5).requires_grad_(True)
    b = torch.randn(Cout, device=DEV, generator=g)
    ref = F.conv2d(x, w, b, padding=p)
    if relu:
        ref = ref.relu()
    Ho, Wo = ref.shape[2:]
    dy = torch.randn(ref.shape, device=DEV, generator=g)
    # ReLU is not part of the wgrad/dgrad kernels: take the gradients of the linear part
    lin = F.conv2d(x, w, None, padding=p)
    lin.backward(dy)
    xn = x.detach().permute(0, 2, 3, 1).contiguous()
    wp = torch.empty(Cout, k * k, Cin, device=DEV)
    q.lib.call("qeb_pack_weight", w.data_ptr(), wp.data_ptr(), Cout, Cin, k, k, 0, st())
    cs = Cout if Cout % 4 == 0 else Cout  # dense rows, also when unaligned (Cout = 95)
    out = torch.full((N, Ho, Wo, cs), float("nan"), device=DEV)
    q.lib.call("qeb_conv_fprop_tc", xn.data_ptr(), N, H, W, Cin, Cin, wp.data_ptr(), Cout, k, k, p, p, b.data_ptr(), None, relu,
               out.data_ptr(), cs, 0, st())
    assert not torch.isnan(out).any()
    assert rel(out.permute(0, 3, 1, 2), ref) < 3e-3
    if Cout % 32 == 0:
        # input gradient = the same kernel on dy with flipped, transposed weights
        dyn = dy.permute(0, 2, 3, 1).contiguous()
        wd = torch.empty(Cin, k * k, Cout, device=DEV)
        q.lib.call("qeb_pack_weight", w.data_ptr(), wd.data_ptr(), Cout, Cin, k, k, 1, st())
        dx = torch.empty(N, H, W, Cin, device=DEV)
        q.lib.call("qeb_conv_fprop_tc", dyn.data_ptr(), N, Ho, Wo, Cout, Cout, wd.data_ptr(), Cin, k, k, k - 1 - p, k - 1 - p, None,
                   None, 0, dx.data_ptr(), Cin, 0, st())
        assert rel(dx.permute(0, 3, 1, 2), x.grad) < 3e-3
        dw = torch.zeros_like(w)
        q.lib.call("qeb_conv_wgrad_tc", xn.data_ptr(), Cin, Cin, H, W, dyn.data_ptr(), Cout, Cout, N, k, k, p, p, dw.data_ptr(), st())
        assert rel(dw, w.grad) < 3e-3
        q.lib.call("qeb_conv_wgrad_tc", xn.data_ptr(), Cin, Cin, H, W, dyn.data_ptr(), Cout, Cout, N, k, k, p, p, dw.data_ptr(), st())
        assert rel(dw, 2 * w.grad) < 3e-3  # accumulates


def test_conv_tc_channel_slices_scale_and_accumulate(q):
    """Concat buffers are read / written in place (channel stride > channels); per-channel scale; out += result."""
    g = torch.Generator(device=DEV).manual_seed(5)
    N, H, W = 2, 16, 24
    buf = torch.randn(N, H, W, 96, device=DEV, generator=g)            # x = channels [32, 96)
    w = torch.randn(32, 64, 3, 3, device=DEV, generator=g) / 24
    sc, sh = torch.rand(32, device=DEV, generator=g) + 0.5, torch.randn(32, device=DEV, generator=g)
    x = buf[..., 32:].permute(0, 3, 1, 2)
    ref = (F.conv2d(x, w, None, padding=1) * sc.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1)).relu()
    wp = torch.empty(32, 9, 64, device=DEV)
    q.lib.call("qeb_pack_weight", w.data_ptr(), wp.data_ptr(), 32, 64, 3, 3, 0, st())
    out = torch.zeros(N, H, W, 64, device=DEV)                           # result goes to channels [32, 64)
    out[..., 32:] = 1.0
    q.lib.call("qeb_conv_fprop_tc", buf.data_ptr() + 32 * 4, N, H, W, 64, 96, wp.data_ptr(), 32, 3, 3, 1, 1, sh.data_ptr(),
               sc.data_ptr(), 1, out.data_ptr() + 32 * 4, 64, 1, st())
    assert float(out[..., :32].abs().max()) == 0.0
    assert rel(out[..., 32:].permute(0, 3, 1, 2), ref + 1.0) < 3e-3


@pytest.mark.parametrize("N,H,W,Cin,Cout", [(4, 4, 16, 256, 128), (2, 16, 64, 64, 32), (1, 2, 8, 512, 256)])
def test_conv_transpose_tc(q, N, H, W, Cin, Cout):
    g = torch.Generator(device=DEV).manual_seed(Cin)
    x = torch.randn(N, Cin, H, W, device=DEV, generator=g, requires_grad=True)
    w = (torch.randn(Cin, Cout, 2, 2, device=DEV, generator=g) / Cin ** 0.5).requires_grad_(True)
    b = torch.randn(Cout, device=DEV, generator=g)
    ref = F.conv_transpose2d(x, w, b, stride=2)
    dy = torch.randn(ref.shape, device=DEV, generator=g)
    ref.backward(dy)
    xn = x.detach().permute(0, 2, 3, 1).contiguous()
    dyn = dy.permute(0, 2, 3, 1).contiguous()
    wp = torch.empty(4 * Cout, Cin, device=DEV)
    q.lib.call("qeb_pack_weight", w.data_ptr(), wp.data_ptr(), Cin, Cout, 2, 2, 2, st())
    out = torch.full((N, 2 * H, 2 * W, Cout), float("nan"), device=DEV)
    q.lib.call("qeb_convT2x2_fprop_tc", xn.data_ptr(), N, H, W, Cin, Cin, wp.data_ptr(), b.data_ptr(), Cout, out.data_ptr(), Cout, st())
    assert rel(out.permute(0, 3, 1, 2), ref) < 3e-3
    wd = torch.empty(Cin, 4 * Cout, device=DEV)
    q.lib.call("qeb_pack_weight", w.data_ptr(), wd.data_ptr(), Cin, Cout, 2, 2, 0, st())
    dx = torch.empty(N, H, W, Cin, device=DEV)
    q.lib.call("qeb_convT2x2_dgrad_tc", dyn.data_ptr(), N, H, W, Cout, Cout, wd.data_ptr(), Cin, dx.data_ptr(), Cin, st())
    assert rel(dx.permute(0, 3, 1, 2), x.grad) < 3e-3
    dw = torch.zeros_like(w)
    q.lib.call("qeb_convT2x2_wgrad_tc", xn.data_ptr(), Cin, Cin, H, W, dyn.data_ptr(), Cout, Cout, N, dw.data_ptr(), st())
    assert rel(dw, w.grad) < 3e-3


# ------------------------------------------------------------------------------------------------ LSTM recurrence
@pytest.mark.parametrize("T,B", [(31, 64), (31, 20), (15, 3), (63, 9)])
def test_lstm_layer_vs_torch(q, T, B):
    torch.manual_seed(T + B)
    lstm = torch.nn.LSTM(512, 256, 1, bidirectional=True).to(DEV)
    x = torch.randn(T, B, 512, device=DEV)
    y_ref, _ = lstm(x)
    dy = torch.randn_like(y_ref)
    y_ref.backward(dy)
    gates = torch.empty(T, B, 2, 1024, device=DEV)
    with torch.no_grad():
        gates[:, :, 0] = x @ lstm.weight_ih_l0.T + lstm.bias_ih_l0 + lstm.bias_hh_l0
        gates[:, :, 1] = x @ lstm.weight_ih_l0_reverse.T + lstm.bias_ih_l0_reverse + lstm.bias_hh_l0_reverse
    cells = torch.empty(T, B, 2, 256, device=DEV)
    y = torch.empty(T, B, 512, device=DEV)
    q.lib.call("qeb_lstm_layer_fwd", gates.data_ptr(), lstm.weight_hh_l0.data_ptr(), lstm.weight_hh_l0_reverse.data_ptr(),
               cells.data_ptr(), y.data_ptr(), T, B, st())
    # the recurrent operands (W_hh, h, d gates) are rounded to tf32: 2^-12 relative per element
    assert rel(y, y_ref) < 2e-4
    q.lib.call("qeb_lstm_layer_bwd", gates.data_ptr(), cells.data_ptr(), dy.contiguous().data_ptr(), lstm.weight_hh_l0.data_ptr(),
               lstm.weight_hh_l0_reverse.data_ptr(), T, B, st())
    dg = gates.reshape(T * B, 2, 1024)
    xf = x.reshape(T * B, 512)
    assert rel(dg[:, 0].T @ xf, lstm.weight_ih_l0.grad) < 5e-4
    assert rel(dg[:, 1].T @ xf, lstm.weight_ih_l0_reverse.grad) < 5e-4
    assert rel(dg[:, 1].sum(0), lstm.bias_hh_l0_reverse.grad) < 5e-4
    hprev = torch.zeros(T, B, 256, device=DEV)
    hprev[1:] = y[:-1, :, :256]
    assert rel(dg[:, 0].T @ hprev.reshape(T * B, 256), lstm.weight_hh_l0.grad) < 5e-4


# ------------------------------------------------------------------------------------------------ networks vs golden
def _encode(labels):
    c2i = {c: i for i, c in enumerate(CHAR_SET)}
    y = torch.tensor([c2i[c] for l in labels for c in l], dtype=torch.int32)
    return y, torch.tensor([len(l) for l in labels], dtype=torch.int32)


def test_crnn_golden_from_reference(q):
    """The reference's CRNN step of train_crnn.py:157-162 on the fixture inputs (seed-42 init, one infeasible label),
    then the phase-B mode (train() + set_bn_eval, gradient w.r.t. the input image, train_nn_area.py:277-286)."""
    g = load_golden("crnn.npz")
    dig = json.load(open(os.path.join(GOLDEN, "crnn_digest.json")))
    torch.manual_seed(42)
    m = q.CRNN(len(CHAR_SET), False).to(DEV)
    m.register_backward_hook(m.backward_hook)
    m.train()
    x = torch.from_numpy(g["x"]).to(DEV)
    B = x.shape[0]
    scores = m(x)
    assert scores.shape == g["scores_train"].shape
    np.testing.assert_allclose(scores.detach().cpu().numpy(), g["scores_train"], atol=2e-3)
    y, ylen = torch.from_numpy(g["targets"]), torch.from_numpy(g["target_lengths"])
    il = torch.tensor([scores.shape[0]] * B, dtype=torch.int32)
    loss = q.ctc.CTCLoss()(scores, y, il, ylen)
    assert torch.isinf(loss) and np.isinf(g["loss_train"])          # infeasible sample: loss stays inf ...
    loss.backward()
    for p in m.parameters():                                         # ... and the hook keeps every gradient finite
        assert torch.isfinite(p.grad).all()
    np.testing.assert_allclose(m.convo.batchnorm1.running_mean.cpu().numpy(), g["bn1_mean"], rtol=2e-3, atol=2e-5)
    np.testing.assert_allclose(m.convo.batchnorm1.running_var.cpu().numpy(), g["bn1_var"], rtol=2e-3, atol=2e-5)
    assert int(m.convo.batchnorm1.num_batches_tracked) == 1
    for name, ref in (("linear.bias", g["grad_linear_b"]), ("convo.conv1.weight", g["grad_conv1_w"])):
        got = dict(m.named_parameters())[name].grad
        assert cos(got, torch.from_numpy(ref)) > 0.995 and rel(got, torch.from_numpy(ref)) < 0.1, name
    assert cos(m.lstm.weight_ih_l0.grad[:16], torch.from_numpy(g["grad_lstm_w_ih_l0"])) > 0.995
    assert cos(m.convo.conv7.weight.grad[:2], torch.from_numpy(g["grad_conv7_w"])) > 0.995
    for name, d in dig["grad_digest_train"].items():                 # every parameter: gradient norm as the reference's
        got = dict(m.named_parameters())[name].grad
        if name in ("convo.conv5.bias", "convo.conv6.bias"):         # conv bias before a train-mode BN: analytically
            # zero; the reference holds fp32 rounding noise, here the noise of summing dz rounded to tf32 (a dgrad operand)
            assert float(got.abs().max()) < 2e-3 and d["norm"] < 1e-4, name
        else:
            assert abs(float(got.double().norm()) - d["norm"]) < 0.1 * d["norm"], name
    # phase-B mode
    m.zero_grad()
    m.train()
    m.apply(q.utils.set_bn_eval)
    xg = x.clone().requires_grad_(True)
    scores2 = m(xg)
    np.testing.assert_allclose(scores2.detach().cpu().numpy(), g["scores_bneval"], atol=2e-3)
    y2, y2len = torch.from_numpy(g["targets_b"]), torch.from_numpy(g["target_lengths_b"])
    loss2 = q.ctc.CTCLoss()(scores2, y2, il, y2len)
    np.testing.assert_allclose(float(loss2), float(g["loss_bneval"]), rtol=1e-3)   # north_star: CTC loss within 1e-3
    loss2.backward()
    gx = torch.from_numpy(g["grad_x_bneval"])
    assert cos(xg.grad, gx) > 0.995 and rel(xg.grad, gx) < 0.1
    for name, d in dig["grad_digest_bneval"].items():
        got = dict(m.named_parameters())[name].grad
        assert abs(float(got.double().norm()) - d["norm"]) < 0.1 * d["norm"] + 1e-7, name


def test_unet_golden_from_reference(q):
    g = load_golden("unet.npz")
    dig = json.load(open(os.path.join(GOLDEN, "unet_digest.json")))
    torch.manual_seed(42)
    m = q.UNet().to(DEV)
    m.train()
    x = torch.from_numpy(g["x"]).to(DEV)
    y = m(x)
    np.testing.assert_allclose(y.detach().cpu().numpy(), g["y_train"], atol=2e-3)
    loss = q.train_ops.mse_to_ones(y)
    np.testing.assert_allclose(float(loss), float(g["loss_train"]), rtol=2e-3)
    loss.backward()
    np.testing.assert_allclose(m.encoder1.enc1norm1.running_mean.cpu().numpy(), g["enc1norm1_mean"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(m.encoder1.enc1norm1.running_var.cpu().numpy(), g["enc1norm1_var"], rtol=1e-4, atol=1e-6)
    for name, ref in (("encoder1.enc1conv1.weight", g["grad_enc1conv1_w"]), ("upconv4.bias", g["grad_upconv4_b"]),
                      ("conv.weight", g["grad_conv_w"])):
        got = dict(m.named_parameters())[name].grad
        assert cos(got, torch.from_numpy(ref)) > 0.99 and rel(got, torch.from_numpy(ref)) < 0.15, name
    for name, d in dig["grad_digest_train"].items():
        got = dict(m.named_parameters())[name].grad
        assert abs(float(got.double().norm()) - d["norm"]) < 0.15 * d["norm"] + 1e-9, name
    m.eval()
    with torch.no_grad():
        np.testing.assert_allclose(m(x).cpu().numpy(), g["y_eval"], atol=2e-3)


# ------------------------------------------------------------------------------------------------ networks vs oracle
def _fp16_backward_launches(rep):
    return sum(v["launches"] for k, v in rep.items() if k in ("tc_conv_wgrad.f16",)), rep


@pytest.mark.parametrize("B,W,mode", [(64, 128, "train"), (64, 128, "bneval"), (5, 256, "train"), (3, 64, "bneval")])
def test_crnn_vs_oracle(q, B, W, mode):
    _crnn_vs_oracle(q, B, W, mode, False)


@pytest.mark.parametrize("B,W,mode", [(64, 128, "train"), (64, 128, "bneval"), (5, 256, "bneval")])
def test_crnn_vs_oracle_fp16_backward(q, B, W, mode):
    """The same comparison for the SECOND backward call of a network: the first records the gradient maxima (tf32 operands), from
    then on the conv stack's dgrad / wgrad contractions read scaled fp16 shadows (nn.cuh GradShadow)."""
    _crnn_vs_oracle(q, B, W, mode, True)


def _crnn_vs_oracle(q, B, W, mode, prime):
    torch.manual_seed(B + W)
    m = q.CRNN(95, False).to(DEV)
    with torch.no_grad():
        for bn in (m.convo.batchnorm1, m.convo.batchnorm2):
            bn.running_mean.normal_(0, 0.1); bn.running_var.uniform_(0.5, 1.5); bn.weight.uniform_(0.5, 1.5); bn.bias.normal_(0, 0.1)
    mr = copy.deepcopy(m)
    for mm in (m, mr):
        mm.train()
        if mode == "bneval":
            mm.apply(q.utils.set_bn_eval)
    x = torch.rand(B, 1, 32, W, device=DEV)
    if prime:   # a backward call on another batch (a third of the loss: the scales must not depend on matching magnitudes)
        bufs = copy.deepcopy(m.state_dict())
        lp0 = m(torch.rand(B, 1, 32, W, device=DEV).requires_grad_(True))
        yl0 = torch.randint(1, 9, (B,), dtype=torch.int32)
        (q.ctc.CTCLoss()(lp0, torch.randint(1, 95, (int(yl0.sum()),), dtype=torch.int32), torch.full((B,), lp0.shape[0], dtype=torch.int32), yl0) / 3).backward()
        m.zero_grad()
        m.load_state_dict(bufs)   # BatchNorm running statistics as before the priming call
        q.lib.prof_enable(True)
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    lp, lpr = m(xa), nn_oracle.crnn_forward(mr, xb)
    assert lp.shape == (W // 4 - 1, B, 95)
    assert float((lp - lpr).abs().max()) < 2e-3
    # gradients of a CTC loss (the real upstream gradient)
    ylen = torch.randint(1, 9, (B,), dtype=torch.int32)
    y = torch.randint(1, 95, (int(ylen.sum()),), dtype=torch.int32)
    il = torch.full((B,), lp.shape[0], dtype=torch.int32)
    la = q.ctc.CTCLoss()(lp, y, il, ylen)
    lb = torch.nn.CTCLoss()(lpr, y.to(DEV), il.to(DEV), ylen.to(DEV))
    assert abs(float(la) - float(lb)) < 1e-3 * abs(float(lb))
    la.backward(); lb.backward()
    if prime:
        rep = q.lib.prof_report()
        q.lib.prof_enable(False)
        assert rep.get("tc_conv_wgrad.f16", {}).get("launches", 0) >= 6, rep.keys()   # conv2..conv7 weight gradients
    assert cos(xa.grad, xb.grad) > 0.99
    for (n, p), (_, r) in zip(m.named_parameters(), mr.named_parameters()):
        if mode == "train" and n in ("convo.conv5.bias", "convo.conv6.bias"):   # analytically zero (train-mode BN follows)
            assert float(p.grad.abs().max()) < 4e-3 and float(r.grad.abs().max()) < 1e-4, n
        else:
            assert cos(p.grad, r.grad) > 0.99, (n, cos(p.grad, r.grad))
    if mode == "train":
        assert rel(m.convo.batchnorm2.running_var, mr.convo.batchnorm2.running_var) < 1e-4


@pytest.mark.parametrize("B,H,W,mode", [(64, 32, 128, "train"), (1, 400, 512, "train"), (2, 48, 80, "eval")])
def test_unet_vs_oracle(q, B, H, W, mode):
    _unet_vs_oracle(q, B, H, W, mode, False)


@pytest.mark.parametrize("B,H,W,mode", [(64, 32, 128, "train"), (2, 48, 80, "train"), (2, 48, 80, "eval")])
def test_unet_vs_oracle_fp16_backward(q, B, H, W, mode):
    """Second backward call of the network: scaled fp16 operands in every tensor-core dgrad / wgrad (see the CRNN twin)."""
    _unet_vs_oracle(q, B, H, W, mode, True)


def _unet_vs_oracle(q, B, H, W, mode, prime):
    torch.manual_seed(H + W)
    m = q.UNet().to(DEV)
    with torch.no_grad():
        for mod in m.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.running_mean.normal_(0, 0.1); mod.running_var.uniform_(0.5, 1.5); mod.weight.uniform_(0.5, 1.5); mod.bias.normal_(0, 0.1)
    mr = copy.deepcopy(m)
    for mm in (m, mr):
        mm.train() if mode == "train" else mm.eval()
    x = torch.rand(B, 1, H, W, device=DEV)
    if prime:
        bufs = copy.deepcopy(m.state_dict())
        (q.train_ops.mse_to_ones(m(torch.rand(B, 1, H, W, device=DEV).requires_grad_(True))) / 3).backward()
        m.zero_grad()
        m.load_state_dict(bufs)
        q.lib.prof_enable(True)
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    y, yr = m(xa), nn_oracle.unet_forward(mr, xb)
    assert float((y - yr).abs().max()) < 3e-3
    (q.train_ops.mse_to_ones(y)).backward()
    torch.nn.MSELoss()(yr, torch.ones_like(yr)).backward()
    if prime:
        rep = q.lib.prof_report()
        q.lib.prof_enable(False)
        assert rep.get("tc_conv_wgrad.f16", {}).get("launches", 0) >= 17, rep.keys()   # 17 conv units + 4 up-convolutions
    assert cos(xa.grad, xb.grad) > 0.98
    for (n, p), (_, r) in zip(m.named_parameters(), mr.named_parameters()):
        assert cos(p.grad, r.grad) > 0.98, (n, cos(p.grad, r.grad))
    if mode == "train":
        assert rel(m.decoder1.dec1norm2.running_var, mr.decoder1.dec1norm2.running_var) < 1e-4
        assert int(m.bottleneck.bottlenecknorm1.num_batches_tracked) == 1


def test_greedy_decode_parity_margin_aware(q):
    """Decoded strings against the fp32 graph on 1024 random-init patches. Random-init log-probs are nearly flat, so
    the arg-max is decided by margins at the TF32 noise level for part of the batch: strings whose
    reference top-1/top-2 margin exceeds 10x the measured log-prob error must ALL agree, and the overall rate is
    reported."""
    torch.manual_seed(3)
    m = q.CRNN(95, False).to(DEV)
    mr = copy.deepcopy(m)
    m.eval(); mr.eval()
    agree = safe = safe_agree = total = 0
    with torch.no_grad():
        for chunk in range(4):
            x = torch.rand(256, 1, 32, 128, device=DEV)
            lp, lpr = m(x), nn_oracle.crnn_forward(mr, x)
            err = float((lp - lpr).abs().max())
            top2 = lpr.topk(2, dim=2).values
            margin = (top2[..., 0] - top2[..., 1]).min(dim=0).values        # (B)
            ca, la = q.utils.decode_batch(lp)
            cb, lb = q.utils.decode_batch(lpr.contiguous())
            same = ((ca == cb).all(dim=1) & (la == lb))
            ok = margin > 10 * err
            agree += int(same.sum()); total += same.numel(); safe += int(ok.sum()); safe_agree += int((same & ok).sum())
    assert safe_agree == safe
    assert agree >= 0.9 * total
    print(f"decode parity: {agree}/{total} identical; {safe_agree}/{safe} of the margin-safe strings")


# ------------------------------------------------------------------------------------------------ training-step pieces
def test_mse_and_adam_match_torch(q):
    torch.manual_seed(0)
    x = torch.rand(3, 1, 32, 128, device=DEV, requires_grad=True)
    xr = x.detach().clone().requires_grad_(True)
    a = q.train_ops.mse_to_ones(x); b = torch.nn.MSELoss()(xr, torch.ones_like(xr))
    (3 * a).backward(); (3 * b).backward()
    assert abs(float(a) - float(b)) < 1e-6 and rel(x.grad, xr.grad) < 1e-6
    ps = [torch.randn(s, device=DEV).requires_grad_(True) for s in ((64, 1, 3, 3), (64,), (1000, 37), (5,))]
    pr = [p.detach().clone().requires_grad_(True) for p in ps]
    oa = q.train_ops.Adam(ps, lr=1e-3, weight_decay=5e-4)
    ob = torch.optim.Adam(pr, lr=1e-3, weight_decay=5e-4)
    for it in range(5):
        for p, r in zip(ps, pr):
            gr = torch.randn_like(p)
            p.grad = gr.clone(); r.grad = gr.clone()
        oa.step(); ob.step()
    for p, r in zip(ps, pr):
        assert rel(p, r) < 1e-6
    sd = oa.state_dict()                                   # same state layout as torch.optim.Adam
    assert set(sd["state"][0].keys()) == set(ob.state_dict()["state"][0].keys())
    ob.load_state_dict(sd)


def test_full_step_trains(q):
    """Phase B of train_nn_area.py:277-287 through the mirror modules: the loss goes down and only the UNet moves."""
    torch.manual_seed(1)
    prep, crnn = q.UNet().to(DEV), q.CRNN(95, False).to(DEV)
    crnn.register_backward_hook(crnn.backward_hook)
    opt = q.train_ops.Adam(prep.parameters(), lr=2e-3)
    x = torch.rand(16, 1, 32, 128, device=DEV)
    ylen = torch.randint(1, 7, (16,), dtype=torch.int32)
    y = torch.randint(1, 95, (int(ylen.sum()),), dtype=torch.int32)
    il = torch.full((16,), 31, dtype=torch.int32)
    w0 = crnn.linear.weight.detach().clone()
    losses = []
    for it in range(12):
        prep.train(); crnn.train(); crnn.apply(q.utils.set_bn_eval)
        prep.zero_grad(); crnn.zero_grad()
        img = prep(x)
        loss = q.ctc.CTCLoss()(crnn(img), y, il, ylen) + 1.0 * q.train_ops.mse_to_ones(img)
        loss.backward()
        opt.step()
        losses.append(float(loss))
    assert losses[-1] < losses[0] and all(np.isfinite(losses))
    assert torch.equal(w0, crnn.linear.weight)
    assert crnn.linear.weight.grad is not None   # the reference also computes (and discards) the surrogate's gradients


def test_phase_a_jitter_step_vs_oracle(q):
    """BASELINE configs[2] shape on one GPU: one inner iteration of the black-box approximation phase
    (train_nn_area.py:245-275 / train_nn_patch.py:278-309): jitter the preprocessor output, run the surrogate in
    train mode (batch statistics), CTC against the OCR strings, backward, Adam on the surrogate. The oracle replays the
    same noise (the kernel returns it) through the reference arithmetic (transform_helper.py:40-41)."""
    from qeb_b200.mirror import transform_helper as th
    torch.manual_seed(11)
    B = 64
    m = q.CRNN(95, False).to(DEV)
    m.register_backward_hook(m.backward_hook)
    mr = copy.deepcopy(m)
    m.train(); mr.train()
    opt = q.train_ops.Adam(m.parameters(), lr=1e-4, weight_decay=5e-4)          # train_nn_patch.py:146-148
    optr = torch.optim.Adam(mr.parameters(), lr=1e-4, weight_decay=5e-4)
    imgs = torch.rand(B, 1, 32, 128, device=DEV)
    noiser = th.AddGaussianNoice(std=5, is_stochastic=True, return_noise=True)
    noisy, noise = th.add_noise(imgs, noiser)
    assert torch.equal(noisy, torch.clamp(imgs - noise, 0, 1))                    # bit-exact jitter arithmetic
    sd = noise.view(B, -1).std(dim=1)
    assert float(sd.max()) < 0.06 and float(noisy.min()) >= 0 and float(noisy.max()) <= 1
    ylen = torch.randint(1, 12, (B,), dtype=torch.int32)
    y = torch.randint(1, 95, (int(ylen.sum()),), dtype=torch.int32)
    il = torch.full((B,), 31, dtype=torch.int32)
    la = q.ctc.CTCLoss()(m(noisy), y, il, ylen)
    lb = torch.nn.CTCLoss()(nn_oracle.crnn_forward(mr, noisy), y.to(DEV), il.to(DEV), ylen.to(DEV))
    assert abs(float(la) - float(lb)) < 1e-3 * abs(float(lb))
    la.backward(); lb.backward()
    for (n, p), (_, r) in zip(m.named_parameters(), mr.named_parameters()):
        if n in ("convo.conv5.bias", "convo.conv6.bias"):
            continue
        assert cos(p.grad, r.grad) > 0.99, (n, cos(p.grad, r.grad))
    opt.step(); optr.step()
    for (n, p), (_, r) in zip(m.named_parameters(), mr.named_parameters()):
        # one Adam step moves every weight by ~lr in the direction of sign(grad): compare the updates
        assert float((p - r).abs().max()) <= 2.1e-4, n
    assert rel(m.convo.batchnorm1.running_mean, mr.convo.batchnorm1.running_mean) < 1e-3


@pytest.mark.parametrize("B,W", [(1, 128), (2, 8), (130, 32)])
def test_crnn_edge_shapes(q, B, W):
    """Smallest legal width (W = 8 -> T = 1), a single patch, a batch that is not a multiple of any tile."""
    torch.manual_seed(B)
    m = q.CRNN(95, False).to(DEV)
    mr = copy.deepcopy(m)
    for mm in (m, mr):               # phase-B mode (cuDNN's RNN backward, which the oracle runs on, needs train())
        mm.train(); mm.apply(q.utils.set_bn_eval)
    x = torch.rand(B, 1, 32, W, device=DEV, requires_grad=True)
    xr = x.detach().clone().requires_grad_(True)
    lp, lpr = m(x), nn_oracle.crnn_forward(mr, xr)
    assert lp.shape == lpr.shape == (W // 4 - 1, B, 95)
    assert float((lp - lpr).abs().max()) < 2e-3
    g = torch.randn_like(lp)
    lp.backward(g); lpr.backward(g)
    assert cos(x.grad, xr.grad) > 0.98
    with pytest.raises(q.lib.QebError):
        m(torch.rand(1, 1, 16, 128, device=DEV))     # the reference's map_to_sequence needs H == 32 too
    with pytest.raises(q.lib.QebError):
        m(torch.rand(1, 1, 32, 128))                 # CPU tensor: no fallback


def test_graphed_step_matches_eager(q):
    """qeb_b200.graphs: one captured forward+loss+backward replays to the same loss and gradients as the eager modules, and
    follows new inputs / labels copied into the static buffers."""
    from qeb_b200.graphs import GraphedStep, StaticTargets
    from qeb_b200.mirror import ctc as qctc, train_ops
    from qeb_b200.mirror.models.model_crnn import CRNN
    from qeb_b200.mirror.models.model_unet import UNet
    from qeb_b200.mirror.utils import set_bn_eval
    torch.manual_seed(3)
    B, V = 16, 95
    prep, crnn = UNet().to(DEV), CRNN(V, False).to(DEV)
    crnn.register_backward_hook(crnn.backward_hook)
    prep.train(); crnn.train(); crnn.apply(set_bn_eval)
    loss_fn = qctc.CTCLoss()
    il = torch.full((B,), 31, dtype=torch.int32)

    def batch(seed):
        g = torch.Generator().manual_seed(seed)
        x = torch.rand(B, 1, 32, 128, generator=g)
        tl = torch.randint(1, 12, (B,), generator=g, dtype=torch.int32)
        y = torch.randint(1, V, (int(tl.sum()),), generator=g, dtype=torch.int32)
        return x, y, tl

    def eager(x, y, tl):
        prep.zero_grad(set_to_none=True); crnn.zero_grad(set_to_none=True)
        img = prep(x.to(DEV))
        loss = loss_fn(crnn(img), y, il, tl) + train_ops.mse_to_ones(img)
        loss.backward()
        return float(loss), [p.grad.clone() for p in prep.parameters()], [p.grad.clone() for p in crnn.parameters()]

    # BatchNorm running statistics advance at every call: snapshot them so that both arms start from the same state
    state = {k: v.clone() for k, v in prep.state_dict().items()}
    ref = [eager(*batch(s)) for s in (11, 12)]
    prep.load_state_dict(state)

    xs = torch.empty(B, 1, 32, 128, device=DEV)
    tg = StaticTargets(B, 31, DEV)
    x0, y0, tl0 = batch(11)
    xs.copy_(x0); tg.load(y0, il, tl0)

    def fwd_bwd():
        img = prep(xs)
        loss = loss_fn(crnn(img), tg) + train_ops.mse_to_ones(img)
        loss.backward()
        return loss

    gs = GraphedStep(fwd_bwd, modules=[prep, crnn], warmup=2)
    assert gs.launches > 100
    for (x, y, tl), (l_ref, gp_ref, gc_ref) in zip((batch(11), batch(12)), ref):
        xs.copy_(x); tg.load(y, il, tl)
        loss = gs()
        torch.cuda.synchronize()
        assert abs(float(loss) - l_ref) <= 1e-4 * abs(l_ref)
        # Gradients are compared at the bar of the other network tests, not bit for bit: split-K partial sums are combined
        # with fp32 atomics, a last-bit difference that crosses a TF32 truncation boundary in the next layer becomes a
        # 2^-10 relative jump of that operand, and the backward pass of the random-init networks amplifies it - two EAGER
        # runs on the same input differ by up to 6 % per parameter tensor (scripts/exp/graph_diff.py, profiles/r1_notes.md).
        for p_, g_ in zip(list(prep.parameters()) + list(crnn.parameters()), gp_ref + gc_ref):
            assert cos(p_.grad, g_) > 0.99
            assert 0.85 < float(p_.grad.norm() / g_.norm().clamp_min(1e-30)) < 1.15
    with pytest.raises(Exception):
        tg.load(torch.ones(B * 40, dtype=torch.int32), il, torch.full((B,), 40, dtype=torch.int32))


def test_validation_batch_matches_reference_loop(q):
    """mirror/eval_ops.validation_batch (8(f).2) against the statements of the reference's validation loop
    (train_nn_area.py:327-341) evaluated with the oracle: same strings, same counts, bit-identical CER sums, loss within
    the CTC tolerance."""
    from oracle import pyoracle as po
    from qeb_b200.mirror import eval_ops
    torch.manual_seed(5)
    B = 24
    prep, crnn = q.UNet().to(DEV).eval(), q.CRNN(95, False).to(DEV).eval()
    c2i = {c: i for i, c in enumerate(CHAR_SET)}
    i2c = {i: c for i, c in enumerate(CHAR_SET)}
    x = torch.rand(B, 1, 32, 128, device=DEV)
    import random
    rng = random.Random(1)
    labels = ["".join(rng.choice(CHAR_SET[1:]) for _ in range(rng.randint(1, 9))) for _ in range(B)]
    fake_ocr = lambda imgs: [l[:-1] + "x" if i % 3 == 0 else l for i, l in enumerate(labels)]   # noqa: E731
    out = eval_ops.validation_batch(prep, crnn, x, labels, c2i, i2c, ocr=fake_ocr)
    # the reference's statements on the oracle
    with torch.no_grad():
        img_r = nn_oracle.unet_forward(copy.deepcopy(prep).cpu(), x.cpu())
        sc_r = nn_oracle.crnn_forward(copy.deepcopy(crnn).cpu(), img_r)
        y, ys = eval_ops.encode_labels(labels, c2i)
        loss_r = float(torch.nn.CTCLoss()(sc_r, y, torch.full((B,), 31, dtype=torch.int32), ys) +
                       torch.nn.MSELoss()(img_r, torch.ones_like(img_r)))
    assert abs(out["loss"] - loss_r) <= 2e-3 * abs(loss_r)
    # decode + scoring are integer work: judged on the GPU's own scores so that tf32 noise cannot flip an argmax
    scores = crnn(prep(x))
    preds_r = po.pred_to_string(scores.detach().cpu().numpy(), i2c)
    assert out["preds"] == preds_r
    ocr_labels = fake_ocr(None)
    assert (out["crt"], out["cer"]) == po.compare_labels(preds_r, labels)
    assert (out["ocr_crt"], out["ocr_cer"]) == po.compare_labels(ocr_labels, labels)
    assert (out["matching_crt"], out["matching_cer"]) == po.compare_labels(preds_r, ocr_labels)
    assert out["ocr_crt"] == B - len(range(0, B, 3))


@pytest.mark.parametrize("N,H,W,Cin,Cout,k,p,relu", [
    (8, 32, 128, 32, 32, 3, 1, 0),      # 32 input channels: 64-byte operand rows (SWIZZLE_64B)
    (8, 16, 64, 64, 64, 3, 1, 1),       # 64-element K blocks
    (4, 8, 32, 128, 256, 3, 1, 0),
    (64, 2, 8, 512, 512, 3, 1, 0),      # split-K with fp16 operands
    (2, 2, 31, 512, 512, 2, 0, 0),      # conv7 geometry (2x2, no padding)
    (1, 1, 1984, 512, 2048, 1, 0, 0),   # LSTM input projection
    (3, 5, 7, 96, 64, 3, 1, 1),         # 96 channels: not a multiple of 64 -> 32-element K blocks; ragged tiles
    (64, 32, 128, 32, 32, 3, 1, 1),     # window variant, 64-byte operand rows (SWIZZLE_64B window)
    (64, 32, 128, 64, 32, 3, 1, 0),     # window variant, 128-byte operand rows
    (1, 400, 512, 32, 64, 3, 1, 1),     # window variant on a document-sized image, two N tiles... one CTA walks both
])
def test_conv_fprop_fp16_operands(q, N, H, W, Cin, Cout, k, p, relu):
    """kind::f16 operand path of the fprop kernel (forward pass): exact on fp16-representable inputs up to fp32 accumulation
    order, and within fp16 rounding (2^-11 relative per operand) of the fp32 convolution; fp16 output shadow = RN(output)."""
    g = torch.Generator(device=DEV).manual_seed(N * 1000 + Cin + Cout)
    x = torch.randn(N, H, W, Cin, device=DEV, generator=g)
    w = torch.randn(Cout, Cin, k, k, device=DEV, generator=g) / (Cin * k * k) ** 0.5
    bias = torch.randn(Cout, device=DEV, generator=g)
    x16 = x.half()
    wp16 = w.permute(0, 2, 3, 1).reshape(Cout, k * k * Cin).contiguous().half()
    Ho, Wo = H + 2 * p - k + 1, W + 2 * p - k + 1
    out = torch.zeros(N, Ho, Wo, Cout, device=DEV)
    out16 = torch.zeros(N, Ho, Wo, Cout, device=DEV, dtype=torch.float16)
    q.lib.call("qeb_conv_fprop_tc16", x16.data_ptr(), N, H, W, Cin, Cin, wp16.data_ptr(), Cout, k, k, p, p, bias.data_ptr(), None, relu,
               out.data_ptr(), Cout, out16.data_ptr(), st())
    ref16 = F.conv2d(x16.double().permute(0, 3, 1, 2), w.half().double(), bias.double(), padding=p)
    ref32 = F.conv2d(x.double().permute(0, 3, 1, 2), w.double(), bias.double(), padding=p)
    if relu:
        ref16, ref32 = ref16.relu(), ref32.relu()
    ref16, ref32 = ref16.permute(0, 2, 3, 1), ref32.permute(0, 2, 3, 1)
    assert rel(out, ref16) < 5e-6          # the products of fp16 operands are exact in fp32; only the summation order differs
    assert rel(out, ref32) < 1e-3          # operand rounding: 2^-11 relative each
    assert torch.equal(out16, out.half())


def test_fp16_shadow_saturates(q):
    """An activation beyond fp16's range must not turn into inf in the operand copy (the fp32 tensor keeps the value)."""
    N, H, W, C = 1, 4, 32, 32
    x16 = torch.full((N, H, W, C), 60000.0, device=DEV).half()
    wp16 = torch.ones(C, C, device=DEV).half()          # 1x1 conv: every output = 32 * 60000 = 1.92e6
    out = torch.zeros(N, H, W, C, device=DEV)
    out16 = torch.zeros(N, H, W, C, device=DEV, dtype=torch.float16)
    q.lib.call("qeb_conv_fprop_tc16", x16.data_ptr(), N, H, W, C, C, wp16.data_ptr(), C, 1, 1, 0, 0, None, None, 0,
               out.data_ptr(), C, out16.data_ptr(), st())
    assert torch.allclose(out, torch.full_like(out, 1.92e6))
    assert torch.isfinite(out16).all() and float(out16.max()) == 65504.0


def test_window_variant_on_every_conv_shape():
    """The window variant of the conv kernel is chosen only for images with many 8 x 16 tiles; QEB_WIN=2 drops that
    condition, so that every 3x3 pad-1 shape of the parametrised conv tests above (small batches, ragged tiles, N tile
    changes inside a CTA, channel slices of wider buffers) runs through it. The switch is read once per process."""
    import subprocess
    import sys
    env = dict(os.environ, QEB_WIN="2")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-q", "-x", "-m", "gpu", "-k",
                        "conv_fprop or channel_slices or unet_vs_oracle or crnn_vs_oracle"], env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


# ------------------------------------------------------------------------------------------------ fp16 backward operands
@pytest.mark.parametrize("N,H,W,Cin,Cout,k,p", [
    (8, 4, 32, 512, 512, 3, 1),      # CRNN conv6: 64-channel rows (SWIZZLE_128B), 256-wide N tiles
    (4, 8, 32, 128, 256, 3, 1),      # CRNN conv3
    (8, 2, 32, 512, 512, 2, 0),      # CRNN conv7 (2 x 2 taps, no padding)
    (16, 32, 128, 32, 32, 3, 1),     # UNet level 1: 32-channel rows (SWIZZLE_64B)
    (12, 32, 128, 64, 32, 3, 1),     # mixed: 64 input channels, 32 output channels -> 32-channel rows
    (40, 16, 64, 32, 64, 3, 1),
    (70, 10, 40, 64, 64, 3, 1),      # odd geometry (partial pixel boxes)
    (1, 1, 1984, 512, 2048, 1, 0),   # LSTM W_ih gradient as a GEMM
    (1, 1, 1984, 512, 96, 1, 0),     # Linear (padded to 96 classes): 32-channel rows, 3 boxes on the N side
    (3, 5, 7, 64, 128, 3, 1),        # boxes larger than the image
])
def test_conv_wgrad_tc16(q, N, H, W, Cin, Cout, k, p):
    """Weight gradient with fp16 MN-major operand shadows (x16; dy16 = dy * 2^s, alpha = 2^-s) against torch fp32."""
    g = torch.Generator(device=DEV).manual_seed(N * 77 + Cin + Cout)
    x = torch.randn(N, Cin, H, W, device=DEV, generator=g)
    w = (torch.randn(Cout, Cin, k, k, device=DEV, generator=g) / (Cin * k * k) ** 0.5).requires_grad_(True)
    lin = F.conv2d(x, w, None, padding=p)
    dy = torch.randn(lin.shape, device=DEV, generator=g) * 1e-4     # gradient-sized values: need the scale
    lin.backward(dy)
    xn = x.permute(0, 2, 3, 1).contiguous()
    dyn = dy.permute(0, 2, 3, 1).contiguous()
    S = 2.0 ** 16
    x16, dy16 = xn.half(), (dyn * S).half()
    alpha = torch.tensor([1.0 / S], device=DEV)
    dw = torch.zeros_like(w)
    q.lib.call("qeb_conv_wgrad_tc16", xn.data_ptr(), x16.data_ptr(), Cin, Cin, H, W, dyn.data_ptr(), dy16.data_ptr(), Cout, Cout, N,
               k, k, p, p, alpha.data_ptr(), dw.data_ptr(), st())
    assert rel(dw, w.grad) < 3e-3
    q.lib.call("qeb_conv_wgrad_tc16", xn.data_ptr(), x16.data_ptr(), Cin, Cin, H, W, dyn.data_ptr(), dy16.data_ptr(), Cout, Cout, N,
               k, k, p, p, alpha.data_ptr(), dw.data_ptr(), st())
    assert rel(dw, 2 * w.grad) < 3e-3  # accumulates


@pytest.mark.parametrize("factor", [256.0, 1.0 / 256.0, 3000.0])
def test_unet_fp16_backward_scale_jump(q, factor):
    """Delayed gradient scales (csrc/nn.cuh GradShadow): the operand copies of a backward call are scaled by the PREVIOUS call's
    maxima with 2^12 of headroom above and 2^18 of normal range below. A loss that jumps by 256x / 3000x up or 256x down between
    two calls must still give gradients that match the fp32 graph (direction and norm), and the call after it is scaled afresh."""
    torch.manual_seed(11)
    m = q.UNet().to(DEV)
    mr = copy.deepcopy(m)
    m.train(); mr.train()
    x = torch.rand(8, 1, 32, 128, device=DEV)
    bufs = copy.deepcopy(m.state_dict())
    q.train_ops.mse_to_ones(m(x)).backward()                 # first call: tf32 reads, records the maxima at loss scale 1
    m.zero_grad(); m.load_state_dict(bufs)
    (factor * q.train_ops.mse_to_ones(m(x))).backward()      # second call: fp16 operands scaled for loss scale 1, gradients x factor
    yr = nn_oracle.unet_forward(mr, x)
    (factor * torch.nn.MSELoss()(yr, torch.ones_like(yr))).backward()
    for (n, p), (_, r) in zip(m.named_parameters(), mr.named_parameters()):
        assert torch.isfinite(p.grad).all(), n
        assert cos(p.grad, r.grad) > 0.98, (n, cos(p.grad, r.grad))
        assert abs(float(p.grad.double().norm()) / float(r.grad.double().norm()) - 1.0) < 0.2, n
    g2 = {n: p.grad.clone() for n, p in m.named_parameters()}
    m.zero_grad(); m.load_state_dict(bufs)
    (factor * q.train_ops.mse_to_ones(m(x))).backward()      # third call: scales from the second call's maxima
    for n, p in m.named_parameters():
        assert cos(p.grad, g2[n]) > 0.995, n


def test_batch_stager_feeds_the_graph_one_batch_ahead(q):
    """qeb_b200.graphs.BatchStager: batches staged on the copy stream one step ahead reach the captured step in order - the
    replayed losses equal those of the serial arrangement (copy, replay, read back on one stream)."""
    from qeb_b200.graphs import BatchStager, GraphedStep, StaticTargets
    from qeb_b200.mirror import ctc as qctc
    torch.manual_seed(7)
    B, V = 8, 95
    crnn = q.CRNN(V, False).to(DEV)
    crnn.train()
    crnn.apply(q.utils.set_bn_eval)
    il = torch.full((B,), 31, dtype=torch.int32)
    xs = torch.empty(B, 1, 32, 128, device=DEV)
    tg = StaticTargets(B, 31, DEV)

    def batch(seed):
        g = torch.Generator().manual_seed(seed)
        tl = torch.randint(1, 12, (B,), generator=g, dtype=torch.int32)
        return (torch.rand(B, 1, 32, 128, generator=g).pin_memory(), torch.randint(1, V, (int(tl.sum()),), generator=g, dtype=torch.int32), tl)

    batches = [batch(s) for s in range(5)]
    xs.copy_(batches[0][0]); tg.load(batches[0][1], il, batches[0][2])
    loss_fn = qctc.CTCLoss()

    def fwd_bwd():
        loss = loss_fn(crnn(xs), tg)
        loss.backward()
        return loss

    gs = GraphedStep(fwd_bwd, modules=[crnn], warmup=2)
    serial = []
    for x, y, tl in batches:
        xs.copy_(x, non_blocking=True); tg.load(y, il, tl)
        serial.append(gs().item())
    st = BatchStager(xs, tg)
    st.stage(batches[0][0], batches[0][1], il, batches[0][2])
    staged = []
    for i in range(len(batches)):
        st.commit()
        loss = gs()
        if i + 1 < len(batches):
            st.stage(batches[i + 1][0], batches[i + 1][1], il, batches[i + 1][2])
        staged.append(loss.item())
    assert len(set(serial)) == len(serial)          # five different batches, five different losses
    assert staged == serial
    with pytest.raises(Exception):
        st.commit()                                   # nothing staged
