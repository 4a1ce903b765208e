"""Parity of the integer / elementwise / CTC kernels (through the C ABI) against the CPU oracle and the golden
fixtures generated from the real reference. Bit-exact for integer and index work; CTC within 1e-3 relative
(north_star), measured here at <= 1e-4."""
import json
import os
import random

import numpy as np
import pytest
import torch

from conftest import CHAR_SET, GOLDEN, load_golden
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def m():
    import qeb_b200  # noqa: F401
    from qeb_b200.mirror import ctc, selection_utils, transform_helper, utils

    class M:
        pass

    M.ctc, M.sel, M.th, M.utils = ctc, selection_utils, transform_helper, utils
    assert qeb_b200._lib.load().qeb_check_device() == 0
    return M


def rand_words(rng, n, lo=0, hi=16, alphabet=CHAR_SET[1:]):
    return ["".join(rng.choice(alphabet) for _ in range(rng.randint(lo, hi))) for _ in range(n)]


def perturb(rng, words, p_same=0.55, alphabet=CHAR_SET[1:]):
    out = []
    for w in words:
        s = list(w)
        if rng.random() >= p_same:
            for _ in range(rng.randint(1, 4)):
                op = rng.random()
                if op < 0.33 and s:
                    del s[rng.randrange(len(s))]
                elif op < 0.66:
                    s.insert(rng.randrange(len(s) + 1), rng.choice(alphabet))
                elif s:
                    s[rng.randrange(len(s))] = rng.choice(alphabet)
        out.append("".join(s))
    return out


# ------------------------------------------------------------------------------------------------ Levenshtein
def test_compare_labels_golden(m):
    g = json.load(open(os.path.join(GOLDEN, "cer.json")))
    correct, total = m.utils.compare_labels(g["preds"], g["labels"])
    assert (correct, total) == (g["correct"], g["total_cer"])
    for p, l, (c, t) in list(zip(g["preds"], g["labels"], g["per_pair"]))[-12:]:
        assert m.utils.compare_labels([p], [l]) == (c, t)


def test_levenshtein_vs_oracle_100k(m):
    rng = random.Random(0)
    labels = rand_words(rng, 100_000)
    preds = perturb(rng, labels)
    dist, cer = m.utils.levenshtein_strings(preds, labels)
    _, total, odist, ocer = po.compare_labels(preds, labels, return_all=True)
    assert np.array_equal(dist, odist)
    assert np.array_equal(cer, ocer)  # fp64 bit-exact


def test_levenshtein_edge_cases(m):
    labels = ["", "", "a", "€uro", "x" * 100, "y" * 128, "z" * 129 + "q", "w" * 300, "ab" * 70]
    preds = ["", "abc", "", "euro€", "x" * 99 + "y", "y" * 127, "z" * 140, "w" * 200 + "v" * 150, "ba" * 70]
    dist, cer = m.utils.levenshtein_strings(preds, labels)
    _, _, odist, ocer = po.compare_labels(preds, labels, return_all=True)
    assert np.array_equal(dist, odist) and np.array_equal(cer, ocer)
    assert m.utils.compare_labels("abc", "abd") == po.compare_labels("abc", "abd")  # non-list label is wrapped


def test_levenshtein_uint8_and_padded_rows(m):
    rng = random.Random(3)
    n, T = 1000, 31
    lab = [[rng.randrange(1, 95) for _ in range(rng.randint(0, 20))] for _ in range(n)]
    prd = [[rng.randrange(1, 95) for _ in range(rng.randint(0, T))] for _ in range(n)]
    odist = np.array([po.levenshtein("".join(map(chr, a)), "".join(map(chr, b))) for a, b in zip(lab, prd)])
    for dt in (torch.uint8, torch.int32):
        a = torch.tensor([c for r in lab for c in r], dtype=dt, device=DEV)
        aoff = torch.tensor(np.concatenate([[0], np.cumsum([len(r) for r in lab])]), dtype=torch.int32, device=DEV)
        b = torch.full((n, T), 0, dtype=dt, device=DEV)
        for i, r in enumerate(prd):
            b[i, : len(r)] = torch.tensor(r, dtype=dt)
        boff = torch.arange(0, n * T, T, dtype=torch.int32, device=DEV)
        blen = torch.tensor([len(r) for r in prd], dtype=torch.int32, device=DEV)
        dist, cer = m.utils.cer_batch(a, aoff, None, b.reshape(-1), boff, blen, n, T)
        assert np.array_equal(dist.cpu().numpy(), odist)
        assert np.array_equal(cer.cpu().numpy(), odist / np.maximum(1, [len(r) for r in lab]))


def test_levenshtein_full_size_properties(m):
    """BASELINE config 5 size (1M pairs): size-independent properties instead of the slow oracle."""
    rng = np.random.default_rng(0)
    n = 1_000_000
    la = rng.integers(0, 17, n)
    a = rng.integers(1, 95, int(la.sum()), dtype=np.int32)
    aoff = np.concatenate([[0], np.cumsum(la)]).astype(np.int32)
    # b = a with the first symbol dropped on even pairs (distance exactly 1 when len>0), identical on odd pairs
    keep = np.ones(len(a), dtype=bool)
    drop = (np.arange(n) % 2 == 0) & (la > 0)
    keep[aoff[:-1][drop]] = False
    b = a[keep]
    lb = la - drop
    boff = np.concatenate([[0], np.cumsum(lb)]).astype(np.int32)
    ta, tao, tb, tbo = (torch.from_numpy(x).to(DEV) for x in (a, aoff, b, boff))
    d_ab, cer = m.utils.cer_batch(ta, tao, None, tb, tbo, None, n, 16)
    d_ba, _ = m.utils.cer_batch(tb, tbo, None, ta, tao, None, n, 16)
    d_aa, _ = m.utils.cer_batch(ta, tao, None, ta, tao, None, n, 16)
    d_ab = d_ab.cpu().numpy()
    assert np.array_equal(d_ab, drop.astype(np.int32))          # known construction
    assert np.array_equal(d_ab, d_ba.cpu().numpy())             # symmetry
    assert not d_aa.any()                                       # identity
    assert np.array_equal(cer.cpu().numpy(), d_ab / np.maximum(1, la))
    # checksum of a strided oracle sample
    idx = np.arange(0, n, 997)
    for i in idx[:200]:
        sa, sb = a[aoff[i]:aoff[i + 1]], b[boff[i]:boff[i + 1]]
        assert d_ab[i] == po.levenshtein("".join(map(chr, sa)), "".join(map(chr, sb)))


# ------------------------------------------------------------------------------------------------ greedy decode
def test_decode_golden(m):
    g = load_golden("decode.npz")
    idx2c = {i: c for i, c in enumerate(CHAR_SET)}
    scores = torch.from_numpy(g["scores"]).to(DEV)
    assert m.utils.pred_to_string(scores, [""] * scores.shape[1], idx2c) == list(g["strings"])


def test_decode_vs_oracle_random(m):
    g = torch.Generator().manual_seed(1)
    for (T, B, V) in ((31, 64, 95), (63, 130, 95), (1, 3, 5), (31, 512, 63)):
        s = torch.randn(T, B, V, generator=g)
        s[torch.rand(T, B, generator=g) < 0.5, 0] += 4
        s = torch.round(s * 4) / 4  # many exact ties
        codes, lens = m.utils.decode_batch(s.to(DEV))
        oc, ol = po.greedy_decode(s.numpy())
        assert np.array_equal(lens.cpu().numpy(), ol)
        assert np.array_equal(codes.cpu().numpy(), oc)
    # strided input (a permuted view) and NaN ordering
    s = torch.randn(64, 31, 95, generator=g)
    s[3, 4, 7] = float("nan")
    sv = s.to(DEV).permute(1, 0, 2)
    codes, lens = m.utils.decode_batch(sv)
    oc, ol = po.greedy_decode(s.permute(1, 0, 2).contiguous().numpy())
    assert np.array_equal(codes.cpu().numpy(), oc) and np.array_equal(lens.cpu().numpy(), ol)


def test_decode_and_cer_matches_string_path(m):
    g = torch.Generator().manual_seed(2)
    rng = random.Random(2)
    T, B, V = 31, 64, 95
    c2i = {c: i for i, c in enumerate(CHAR_SET)}
    idx2c = {i: c for i, c in enumerate(CHAR_SET)}
    labels = rand_words(rng, B, 0, 16)
    scores = torch.randn(T, B, V, generator=g)
    scores[torch.rand(T, B, generator=g) < 0.6, 0] += 4
    y = torch.tensor([c2i[c] for c in "".join(labels)], dtype=torch.int32, device=DEV)
    ylen = torch.tensor([len(l) for l in labels], dtype=torch.int32)
    yoff = (torch.cumsum(ylen, 0) - ylen).to(torch.int32).to(DEV)
    codes, lens, dist, cer = m.utils.decode_and_cer(scores.to(DEV), y, yoff, ylen.to(DEV), 16)
    preds = po.pred_to_string(scores.numpy(), idx2c)
    _, _, odist, ocer = po.compare_labels(preds, labels, return_all=True)
    assert np.array_equal(dist.cpu().numpy(), odist) and np.array_equal(cer.cpu().numpy(), ocer)


# ------------------------------------------------------------------------------------------------ selection
def test_select_golden(m):
    g = load_golden("select.npz")
    for s in range(int(g["n_seg"])):
        vals, k = g[f"topk_{s}_vals"], int(g[f"topk_{s}_k"])
        idx = m.sel.topk_cer_indices(vals, k).numpy()
        assert np.array_equal(idx, g[f"topk_{s}_idx_stable"])
        ridx = m.sel.range_cer_indices(vals, torch.from_numpy(g[f"range_{s}_rands"])).numpy()
        assert np.array_equal(ridx, g[f"range_{s}_idx"]), s


def test_sampler_query_api(m):
    rng = random.Random(4)
    names = [f"n{i}" for i in range(64)]
    cers = {n: rng.randrange(0, 9) / rng.randrange(1, 7) for n in names}
    imgs = torch.rand(64, 1, 32, 128, device=DEV)
    labels = [f"l{i}" for i in range(64)]
    sampler = m.sel.datasampler_factory("topKCER")(dict(cers))
    sub, lab, idx = sampler.query(imgs, labels, 8, names)
    v32 = np.array([cers[n] for n in names], dtype=np.float64).astype(np.float32)
    assert np.array_equal(idx.numpy(), po.topk_query(v32, 8))
    assert idx.dtype == torch.long and torch.equal(sub, imgs[idx]) and lab == [labels[i] for i in idx]
    sampler = m.sel.datasampler_factory("rangeCER")(dict(cers))
    torch.manual_seed(9)
    sub, lab, idx = sampler.query(imgs, labels, 8, names)
    torch.manual_seed(9)
    assert np.array_equal(idx.numpy(), po.range_query(v32, torch.rand(8).numpy()))
    sampler.update_cer([0.5, 0.25], names[:2])
    assert sampler.cers[names[0]] == 0.5 and sampler.all_cers[names[1]] == [0.25]
    # names missing from the dict shift indices silently, empty dict -> empty selection (reference behaviour)
    s2 = m.sel.TopKCERSampler({})
    sub, lab, idx = s2.query(imgs, labels, 8, names)
    assert idx.numel() == 0 and sub.shape[0] == 0


def test_topk_segmented_full_size(m):
    """config 5: 1M CERs in 15,625 minibatches of 64, tie-heavy rational values; one launch; vs numpy stable sort."""
    rng = np.random.default_rng(5)
    n_seg, n = 15625, 64
    d = rng.integers(0, 18, (n_seg, n)) * (rng.random((n_seg, n)) > 0.55)
    vals = (d / rng.integers(1, 17, (n_seg, n))).astype(np.float32)
    for prop in (0.5, 0.95):
        k = max(1, int(np.ceil(n * (1 - prop))))
        out = m.sel._segmented(list(vals), [k] * n_seg, torch.device(DEV))
        got = torch.stack(out).numpy()
        want = np.argsort(-vals, axis=1, kind="stable")[:, :k]
        assert np.array_equal(got, want)


# ------------------------------------------------------------------------------------------------ jitter / crop
def test_jitter_golden_exact_arithmetic(m):
    g = load_golden("jitter.npz")
    for i in range(3):
        out = m.th.apply_noise(torch.from_numpy(g["imgs"][i]).to(DEV), torch.from_numpy(g[f"noise_{i}"]), float(g[f"coef_{i}"]))
        assert np.array_equal(out.cpu().numpy(), g[f"out_{i}"])


def test_jitter_philox_stream(m):
    torch.manual_seed(0)
    imgs = torch.rand(3, 1, 8, 16, device=DEV)
    sig = torch.tensor([0.05, 1e-13, 0.02])
    out, noise = m.th.jitter_batch(imgs, sig, mean=0.0, noise_coef=1, seed=1234567891011, return_noise=True)
    want = po.philox_noise(3, 128, sig.numpy().astype(np.float64), 0.0, 1234567891011).reshape(3, 1, 8, 16)
    np.testing.assert_allclose(noise.cpu().numpy(), want, rtol=2e-5, atol=1e-7)
    assert np.array_equal(out.cpu().numpy(), po.jitter_apply(imgs.cpu().numpy(), noise.cpu().numpy(), 1.0))
    # reproducible under the host seed, different across calls
    noiser = m.th.AddGaussianNoice(std=5, is_stochastic=True, return_noise=True)
    torch.manual_seed(3); a, za = noiser.batch(imgs)
    torch.manual_seed(3); b, zb = noiser.batch(imgs)
    c, zc = noiser.batch(imgs)
    assert torch.equal(a, b) and torch.equal(za, zb) and not torch.equal(za, zc)
    o1 = noiser(imgs[0])
    assert o1[0].shape == imgs[0].shape and float(o1[0].min()) >= 0 and float(o1[0].max()) <= 1


def test_jitter_statistics_full_size(m):
    """512 patches (config 3 size): per-image std matches sigma, mean 0, clamp respected."""
    imgs = torch.full((512, 1, 32, 128), 0.5, device=DEV)
    sig = (torch.arange(512) % 6).float() / 100 + 1e-13
    out, noise = m.th.jitter_batch(imgs, sig, seed=7, return_noise=True)
    sd = noise.view(512, -1).std(dim=1).cpu()
    mu = noise.view(512, -1).mean(dim=1).cpu()
    assert torch.allclose(sd, sig, rtol=0.05, atol=1e-6) and mu.abs().max() < 0.004
    assert float(out.min()) >= 0 and float(out.max()) <= 1
    sel = sig > 0.01  # normalise per image: a mixture of different sigmas is not Gaussian (kurtosis > 3)
    z = (noise[sel].view(int(sel.sum()), -1) / sig[sel].to(noise.device)[:, None]).flatten().cpu()
    assert abs(float((z ** 3).mean())) < 0.02 and abs(float((z ** 4).mean()) - 3) < 0.05


def test_crop_pad_golden_and_oracle(m):
    g = load_golden("crop.npz")
    img = torch.from_numpy(g["img"]).to(DEV).requires_grad_(True)
    labels = [{"label": f"l{i}", "x_min": int(b[0]), "y_min": int(b[1]), "x_max": int(b[2]), "y_max": int(b[3])}
              for i, b in enumerate(g["boxes"])]
    out, labs = m.utils.get_text_stack(img, labels, (32, 128))
    assert labs == [f"l{i}" for i in range(len(labels))]
    assert np.array_equal(out.detach().cpu().numpy(), g["out"])
    (out * torch.from_numpy(g["w"]).to(DEV)).sum().backward()
    np.testing.assert_allclose(img.grad.cpu().numpy(), g["grad"], rtol=1e-5, atol=1e-6)
    # document-sized random case incl. oversize boxes (negative pads crop)
    rng = np.random.default_rng(1)
    H, W, n = 400, 512, 124
    im = rng.random((H, W), dtype=np.float32)
    x0 = rng.integers(0, W - 10, n); y0 = rng.integers(0, H - 5, n)
    boxes = np.stack([x0, y0, x0 + rng.integers(0, 160, n), y0 + rng.integers(0, 40, n)], 1).astype(np.int32)
    t = torch.from_numpy(im).to(DEV).reshape(1, H, W).requires_grad_(True)
    lab = [{"label": "", "x_min": int(b[0]), "y_min": int(b[1]), "x_max": int(b[2]), "y_max": int(b[3])} for b in boxes]
    out, _ = m.utils.get_text_stack(t, lab, (32, 128))
    assert np.array_equal(out.detach().cpu().numpy()[:, 0], po.crop_pad(im, boxes, 32, 128))
    w = rng.random((n, 32, 128), dtype=np.float32)
    (out[:, 0] * torch.from_numpy(w).to(DEV)).sum().backward()
    np.testing.assert_allclose(t.grad.cpu().numpy()[0], po.crop_pad_backward(w.astype(np.float64), boxes, H, W), rtol=1e-5, atol=1e-5)


# ------------------------------------------------------------------------------------------------ CTC
def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


def test_ctc_golden(m):
    g = load_golden("ctc.npz")
    lp0 = torch.from_numpy(g["log_probs"]).to(DEV)
    T, B, V = lp0.shape
    y, ylen = torch.from_numpy(g["targets"]), torch.from_numpy(g["target_lengths"])
    il = torch.tensor([T] * B, dtype=torch.int)
    # mean: loss is inf (infeasible sample 2), grads NaN exactly on that sample's column
    lp = lp0.clone().requires_grad_(True)
    loss = m.ctc.CTCLoss()(lp, y, il, ylen)
    loss.backward()
    assert torch.isinf(loss) and np.isinf(g["loss_mean"])
    gr = lp.grad.cpu().numpy()
    assert np.array_equal(np.isnan(gr), np.isnan(g["grad_mean"]))
    assert rel_err(np.nan_to_num(gr), np.nan_to_num(g["grad_mean"])) < 1e-4
    # none, weighted per-sample upstream gradient
    lp = lp0.clone().requires_grad_(True)
    loss = m.ctc.CTCLoss(reduction="none")(lp, y, il, ylen)
    (loss * torch.arange(1, B + 1, device=DEV).float()).sum().backward()
    ln = loss.detach().cpu().numpy()
    assert np.array_equal(np.isinf(ln), np.isinf(g["loss_none"]))
    fin = np.isfinite(ln)
    np.testing.assert_allclose(ln[fin], g["loss_none"][fin], rtol=1e-5)
    assert rel_err(np.nan_to_num(lp.grad.cpu().numpy()), np.nan_to_num(g["grad_none"])) < 1e-4
    # zero_infinity + ragged input lengths
    lp = lp0.clone().requires_grad_(True)
    loss = m.ctc.CTCLoss(zero_infinity=True)(lp, y, torch.from_numpy(g["input_lengths_ragged"]), ylen)
    loss.backward()
    np.testing.assert_allclose(float(loss), float(g["loss_zero_inf_ragged"]), rtol=1e-5)
    assert not torch.isnan(lp.grad).any()
    assert rel_err(lp.grad.cpu().numpy(), g["grad_zero_inf_ragged"]) < 1e-4
    # subset form of weighted_ctc_loss, with and without the gather copy
    idx = g["subset_idx"].tolist()
    ys, yslen = torch.from_numpy(g["subset_targets"]), torch.from_numpy(g["subset_target_lengths"])
    for mode in ("gather", "index"):
        lp = lp0.clone().requires_grad_(True)
        if mode == "gather":
            loss = m.ctc.CTCLoss()(lp[:, idx, :], ys, il[idx], yslen)
        else:
            loss = m.ctc.ctc_loss(lp, ys, il[idx], yslen, batch_index=idx)
        loss.backward()
        np.testing.assert_allclose(float(loss), float(g["subset_loss"]), rtol=1e-5)
        assert rel_err(lp.grad.cpu().numpy(), g["subset_grad"]) < 1e-4


@pytest.mark.parametrize("T,B,V", [(31, 64, 95), (31, 512, 95), (63, 96, 95), (200, 5, 30), (600, 3, 95)])  # the last two: panels too large for shared memory (warp-per-sequence kernels)
def test_ctc_vs_torch_cpu(m, T, B, V):
    g = torch.Generator().manual_seed(T + B)
    logits = torch.randn(T, B, V, generator=g) * 3
    lp0 = torch.log_softmax(logits, 2)
    maxL = min(T // 2, 100)
    ylen = torch.randint(0, maxL + 1, (B,), generator=g, dtype=torch.int)
    y = torch.randint(1, V, (int(ylen.sum()),), generator=g, dtype=torch.int)
    rep = torch.rand(len(y), generator=g) < 0.3  # plenty of repeated characters
    y[1:][rep[1:]] = y[:-1][rep[1:]]
    il = torch.full((B,), T, dtype=torch.int)
    for red in ("mean", "none", "sum"):
        ol, og = po.ctc_loss(lp0.numpy(), y.numpy(), il.numpy(), ylen.numpy(), red)
        lp = lp0.to(DEV).requires_grad_(True)
        loss = m.ctc.CTCLoss(reduction=red)(lp, y, il, ylen)
        (loss.sum() if red == "none" else loss).backward()
        ln = loss.detach().cpu().numpy()
        assert np.array_equal(np.isinf(ln), np.isinf(ol))
        np.testing.assert_allclose(np.nan_to_num(ln, posinf=0), np.nan_to_num(ol, posinf=0), rtol=1e-4)
        gg = lp.grad.cpu().numpy()
        assert np.array_equal(np.isnan(gg), np.isnan(og))
        assert rel_err(np.nan_to_num(gg), np.nan_to_num(og)) < 1e-3  # north_star tolerance; typically ~1e-6


def test_ctc_2d_targets_and_cuda_int_args(m):
    g = torch.Generator().manual_seed(5)
    T, B, V, S = 31, 16, 95, 12
    lp0 = torch.log_softmax(torch.randn(T, B, V, generator=g), 2)
    y2 = torch.randint(1, V, (B, S), generator=g, dtype=torch.int)
    ylen = torch.randint(1, S + 1, (B,), generator=g, dtype=torch.int)
    il = torch.full((B,), T, dtype=torch.int)
    want = torch.nn.functional.ctc_loss(lp0, y2, il, ylen)
    got = m.ctc.CTCLoss()(lp0.to(DEV), y2.to(DEV), il.to(DEV), ylen.to(DEV))
    np.testing.assert_allclose(float(got), float(want), rtol=1e-5)


def test_log_softmax(m):
    x = torch.randn(31, 64, 95, device=DEV) * 5
    x.requires_grad_(True)
    y = m.ctc.log_softmax(x)
    w = torch.randn_like(y)
    (y * w).sum().backward()
    xr = x.detach().cpu().clone().requires_grad_(True)
    yr = torch.log_softmax(xr, 2)
    (yr * w.cpu()).sum().backward()
    assert torch.allclose(y.cpu(), yr, atol=1e-5) and torch.allclose(x.grad.cpu(), xr.grad, atol=1e-4)


def test_pruning_topk_matches_reference_artifacts(m):
    """pruning/methods.topk (8(f).4) on the reference's own artifacts: the kept image names, in order, at 10 % and 50 %
    pruning of the 3,676-image POS CER table (golden vectors held by the reference, pruning/cer_artifacts/)."""
    from qeb_b200.mirror.pruning import methods
    g = json.load(open(os.path.join(GOLDEN, "pruning.json")))
    for pct in (10, 50):
        want = g[f"topk_{pct}_names"]
        got = methods.topk(g["cers"], len(want))
        assert list(got.keys()) == want
        assert all(got[n] == g["cers"][n] for n in want)
    assert methods.topk(g["cers"], 0) == {}
    assert len(methods.topk(g["cers"], 10 ** 6)) == len(g["cers"])
    with pytest.raises(Exception):
        methods.topk({"a": 0.1, "b": 0.1 + 1e-12}, 1)


def test_ocr_handoff_uint8_pixels_and_ring(m):
    """8(f).3: the device conversion gives the pixels ToPILImage(float tensor) gives (pic.mul(255).byte(), tess_helper.py:22),
    bit for bit, and the pinned ring returns them per ticket."""
    from qeb_b200.mirror import ocr_handoff as oh
    g = torch.Generator().manual_seed(0)
    x = torch.rand(37, 1, 32, 128, generator=g)
    x[0, 0, 0, :8] = torch.tensor([0.0, 1.0, 0.5, 1 / 255, 254.999 / 255, 0.999999, 1e-9, 0.003921569])
    want = x.mul(255).byte().numpy().reshape(37, 32, 128)
    xd = x.to(DEV)
    assert np.array_equal(oh.to_uint8(xd).cpu().numpy().reshape(37, 32, 128), want)
    odd = torch.rand(3, 1, 5, 7, generator=g)                    # 105 elements: scalar tail path
    assert np.array_equal(oh.to_uint8(odd.to(DEV)).cpu().numpy(), odd.mul(255).byte().numpy())
    # more pixels than one capped grid covers in a single pass (1184 blocks x 256 threads x 16 pixels = 4.85 M), and a count
    # that is not a multiple of 16: every output byte must be written
    big = torch.rand(5_300_003, generator=g)
    got = oh.to_uint8(big.to(DEV).view(1, -1)).cpu().numpy().reshape(-1)
    assert np.array_equal(got, big.mul(255).byte().numpy())
    ring = oh.OcrHandoff(depth=2)
    t0 = ring.submit(xd)
    t1 = ring.submit(1.0 - xd)
    assert np.array_equal(ring.fetch(t0), want)
    assert np.array_equal(ring.fetch(t1), (1.0 - x).mul(255).byte().numpy().reshape(37, 32, 128))
    seen = []
    labels = ring.labels(t1, lambda u8: seen.append(u8.shape) or ["x"] * u8.shape[0])
    assert labels == ["x"] * 37 and seen == [(37, 32, 128)]
    ring.submit(xd)
    with pytest.raises(Exception):
        ring.fetch(t0)                                           # overwritten: two submits later

    class FakeHelper:                                            # the reference helper's contract: CPU float tensor in
        def get_labels(self, imgs):
            self.pixels = imgs.mul(255).byte().numpy()[:, 0]
            return ["ok"] * imgs.shape[0]
    h = FakeHelper()
    assert oh.OcrFromUint8(h)(want) == ["ok"] * 37 and np.array_equal(h.pixels, want)


def test_global_topk_single_process_uses_device_kernel(m):
    """mirror/dist.global_topk without a process group: the shard reduction alone (qeb_cer_topk_segmented)."""
    from qeb_b200.mirror import dist as qdist
    rng = np.random.RandomState(3)
    v = (rng.randint(0, 40, size=5000) / 8.0).astype(np.float32)
    for k in (1, 33, 5000, 6000):
        got = qdist.global_topk(v, k).numpy()
        assert np.array_equal(got, np.argsort(-v.astype(np.float64), kind="stable")[:k])
