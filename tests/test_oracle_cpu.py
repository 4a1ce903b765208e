"""The CPU oracle against the golden fixtures generated from the real reference, known answers and properties."""
import json
import os

import numpy as np
import pytest
import torch
from hypothesis import given, settings, strategies as st

from conftest import CHAR_SET, GOLDEN, load_golden
from oracle import pyoracle as po
from oracle import refload


def test_levenshtein_known_answers():
    # textbook values (the reference pins none: python-Levenshtein is not vendored)
    assert po.levenshtein("kitten", "sitting") == 3
    assert po.levenshtein("flaw", "lawn") == 2
    assert po.levenshtein("", "abc") == 3
    assert po.levenshtein("abc", "") == 3
    assert po.levenshtein("same", "same") == 0
    assert po.levenshtein("€10", "10€") == 2
    assert po.levenshtein("intention", "execution") == 5


words = st.text(alphabet=CHAR_SET[1:], max_size=24)


@settings(max_examples=300, deadline=None)
@given(words, words, words)
def test_levenshtein_properties(a, b, c):
    d = po.levenshtein(a, b)
    assert d == po.levenshtein(b, a) == po.levenshtein_py(a, b)
    assert abs(len(a) - len(b)) <= d <= max(len(a), len(b))
    assert d <= po.levenshtein(a, c) + po.levenshtein(c, b)
    assert (d == 0) == (a == b)


def test_compare_labels_golden():
    g = json.load(open(os.path.join(GOLDEN, "cer.json")))
    correct, total, dist, cer = po.compare_labels(g["preds"], g["labels"], return_all=True)
    assert correct == g["correct"]
    assert total == g["total_cer"]  # bit-exact fp64, same accumulation order
    for i, (c, t) in enumerate(g["per_pair"]):
        assert (int(dist[i] == 0), cer[i]) == (c, t)


def test_decode_golden():
    g = load_golden("decode.npz")
    idx2c = {i: c for i, c in enumerate(CHAR_SET)}
    assert po.pred_to_string(g["scores"], idx2c) == list(g["strings"])


def test_select_golden():
    g = load_golden("select.npz")
    for s in range(int(g["n_seg"])):
        vals, k = g[f"topk_{s}_vals"], int(g[f"topk_{s}_k"])
        idx = po.topk_query(vals, k)
        assert np.array_equal(idx, g[f"topk_{s}_idx_stable"])
        assert np.array_equal(np.sort(vals[idx]), np.sort(g[f"topk_{s}_ref_values"]))
        ridx = po.range_query(vals, g[f"range_{s}_rands"])
        assert np.array_equal(ridx, g[f"range_{s}_idx"]), s


def test_jitter_golden():
    g = load_golden("jitter.npz")
    for i in range(3):
        out = po.jitter_apply(g["imgs"][i], g[f"noise_{i}"], float(g[f"coef_{i}"]))
        assert np.array_equal(out, g[f"out_{i}"])


def test_crop_golden():
    g = load_golden("crop.npz")
    out = po.crop_pad(g["img"][0], g["boxes"], 32, 128)
    assert np.array_equal(out, g["out"][:, 0])
    grad = po.crop_pad_backward((g["w"][:, 0]).astype(np.float64), g["boxes"], *g["img"].shape[1:])
    np.testing.assert_allclose(grad, g["grad"][0], rtol=1e-6, atol=1e-7)


def test_ctc_golden_is_torch_cpu():
    # the CTC oracle *is* ATen's CPU implementation; the fixture pins the reference call shapes and edge cases
    g = load_golden("ctc.npz")
    T, B, V = g["log_probs"].shape
    loss, grad = po.ctc_loss(g["log_probs"], g["targets"], [T] * B, g["target_lengths"], "mean")
    assert np.isinf(loss) and np.isinf(g["loss_mean"])
    assert np.array_equal(np.isnan(grad), np.isnan(g["grad_mean"]))
    np.testing.assert_allclose(np.nan_to_num(grad), np.nan_to_num(g["grad_mean"]), rtol=1e-5, atol=1e-7)
    assert np.isnan(g["grad_mean"][:, 2]).all() and not np.isnan(g["grad_mean"][:, 3]).any()


def test_philox_known_answer():
    # Random123 known-answer vectors for philox4x32-10
    assert po.philox4x32_10((0, 0, 0, 0), (0, 0)) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert po.philox4x32_10((0xffffffff,) * 4, (0xffffffff, 0xffffffff)) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert po.philox4x32_10((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


@pytest.mark.skipif(not refload.available(), reason="reference not mounted (GPU box)")
def test_fixture_cer_consistency():
    """Every CER in the reference's own fixture is distance/len(label) for an integer distance."""
    cers = json.load(open(os.path.join(refload.REFERENCE_ROOT, "cer_data_utils", "pos_dataset_cers.json")))
    bad = 0
    for k, v in cers.items():
        label = k.split("_")[1]
        x = v * max(1, len(label))
        bad += abs(x - round(x)) > 1e-9
    assert bad == 0 and len(cers) == 73424


@pytest.mark.skipif(not refload.available(), reason="reference not mounted (GPU box)")
def test_oracle_matches_live_reference():
    """Randomised cross-check of the restatements against the imported reference functions."""
    ref = refload.load()
    idx2c = {i: c for i, c in enumerate(CHAR_SET)}
    g = torch.Generator().manual_seed(0)
    for _ in range(5):
        scores = torch.randn(31, 8, 95, generator=g)
        scores[torch.rand(31, 8, generator=g) < 0.5, 0] += 3
        assert po.pred_to_string(scores.numpy(), idx2c) == ref.utils.pred_to_string(scores, [""] * 8, idx2c)
    su = ref.selection_utils
    for trial in range(20):
        n = int(torch.randint(1, 130, (1,), generator=g))
        k = int(torch.randint(1, n + 5, (1,), generator=g))
        vals = (torch.randint(0, 12, (n,), generator=g).double() / torch.randint(1, 9, (n,), generator=g).double()).tolist()
        d = {f"n{i}": v for i, v in enumerate(vals)}
        names = list(d)
        imgs = torch.arange(n).float().reshape(n, 1)
        torch.manual_seed(trial)
        _, _, ridx = su.CerRangeSampler(d).query(imgs, names, k, names)
        torch.manual_seed(trial)
        rands = torch.rand(k)
        v32 = torch.tensor(vals).numpy()
        assert np.array_equal(po.range_query(v32, rands.numpy()), ridx.numpy())
        _, _, tidx = su.TopKCERSampler(d).query(imgs, names, min(k, n), names)
        mine = po.topk_query(v32, min(k, n))
        assert np.array_equal(np.sort(v32[mine]), np.sort(v32[tidx.numpy()]))
        assert np.array_equal(mine, torch.argsort(torch.tensor(vals), descending=True, stable=True)[:min(k, n)].numpy())


# ------------------------------------------------------------------------------------------------ network oracle
def _encode(labels):
    c2i = {c: i for i, c in enumerate(CHAR_SET)}
    y = torch.tensor([c2i[c] for l in labels for c in l], dtype=torch.int32)
    return y, torch.tensor([len(l) for l in labels], dtype=torch.int32)


def test_nn_oracle_crnn_matches_golden():
    """The mirror CRNN's parameter containers initialise exactly like the reference under seed 42 (same state_dict
    keys, same values), and the torch restatement of the graph reproduces the reference's scores, loss and
    gradients (fixture written by oracle/gen_golden.py from the unmodified models/model_crnn.py)."""
    from oracle import nn_oracle
    from qeb_b200.mirror.models.model_crnn import CRNN
    g = load_golden("crnn.npz")
    dig = json.load(open(os.path.join(GOLDEN, "crnn_digest.json")))
    torch.manual_seed(42)
    m = CRNN(len(CHAR_SET), False)
    sd = m.state_dict()
    assert list(sd.keys()) == list(dig["param_digest_seed42"].keys())
    for k, (s, a) in dig["param_digest_seed42"].items():
        v = sd[k].double()
        assert abs(float(v.sum()) - s) <= 1e-9 * max(1, abs(s)) and abs(float(v.abs().sum()) - a) <= 1e-9 * max(1, a), k
    m.register_backward_hook(m.backward_hook)
    m.train()
    x = torch.from_numpy(g["x"])
    scores = nn_oracle.crnn_forward(m, x)
    np.testing.assert_allclose(scores.detach().numpy(), g["scores_train"], atol=2e-5)
    y, ylen = torch.from_numpy(g["targets"]), torch.from_numpy(g["target_lengths"])
    il = torch.tensor([scores.shape[0]] * x.shape[0], dtype=torch.int)
    loss = torch.nn.CTCLoss()(scores, y, il, ylen)
    assert torch.isinf(loss) and np.isinf(g["loss_train"])
    # the reference registers its NaN-scrubbing hook on the module; the restatement is a free function, so scrub here
    scores.register_hook(lambda gr: torch.nan_to_num(gr, nan=0.0))
    loss.backward()
    np.testing.assert_allclose(m.convo.conv1.weight.grad.numpy(), g["grad_conv1_w"], rtol=1e-3, atol=1e-6)
    np.testing.assert_allclose(m.linear.bias.grad.numpy(), g["grad_linear_b"], rtol=1e-3, atol=1e-6)
    np.testing.assert_allclose(m.convo.batchnorm1.running_mean.numpy(), g["bn1_mean"], atol=1e-6)


def test_nn_oracle_unet_matches_golden():
    from oracle import nn_oracle
    from qeb_b200.mirror.models.model_unet import UNet
    g = load_golden("unet.npz")
    dig = json.load(open(os.path.join(GOLDEN, "unet_digest.json")))
    torch.manual_seed(42)
    m = UNet()
    sd = m.state_dict()
    assert list(sd.keys()) == list(dig["param_digest_seed42"].keys())
    for k, (s, a) in dig["param_digest_seed42"].items():
        v = sd[k].double()
        assert abs(float(v.sum()) - s) <= 1e-9 * max(1, abs(s)) and abs(float(v.abs().sum()) - a) <= 1e-9 * max(1, a), k
    m.train()
    x = torch.from_numpy(g["x"])
    y = nn_oracle.unet_forward(m, x)
    np.testing.assert_allclose(y.detach().numpy(), g["y_train"], atol=2e-6)
    loss = torch.nn.MSELoss()(y, torch.ones_like(y))
    np.testing.assert_allclose(float(loss), float(g["loss_train"]), rtol=1e-5)
    loss.backward()
    np.testing.assert_allclose(m.encoder1.enc1conv1.weight.grad.numpy(), g["grad_enc1conv1_w"], rtol=2e-3, atol=1e-7)
    np.testing.assert_allclose(m.conv.weight.grad.numpy(), g["grad_conv_w"], rtol=2e-3, atol=1e-7)
    m.eval()
    with torch.no_grad():
        np.testing.assert_allclose(nn_oracle.unet_forward(m, x).numpy(), g["y_eval"], atol=2e-6)


def test_topk_oracle_against_reference_artifacts():
    """The one result the reference itself holds for the selection path: pruning/cer_artifacts/
    cers_pos_topk_*.json = methods.topk = stable descending sort. The oracle's top-k restatement reproduces the kept
    names in order (fixture copied from the reference's data files by oracle/gen_golden.py gen_pruning)."""
    g = json.load(open(os.path.join(GOLDEN, "pruning.json")))
    names = list(g["cers"].keys())
    v32 = np.array(list(g["cers"].values()), dtype=np.float32)
    for pct in (10, 50):
        want = g[f"topk_{pct}_names"]
        idx = po.topk_query(v32, len(want))
        assert [names[i] for i in idx] == want
