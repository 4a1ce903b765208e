"""The BASELINE.json configs beside the headline (bench_workloads.py): synthetic-data generators and the reference arm's
output contract on CPU; on the GPU the 512-patch jitter step (configs[2]) against the oracle, the 1 M-pair Levenshtein / CER
run (configs[4]) compared IN FULL with oracle/oracle.c, and one whole area minibatch (configs[3]) against the same
statements evaluated with the oracle."""
import copy
import ctypes
import json
import math
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT

import bench as BB
import bench_workloads as BW
from oracle import nn_oracle, pyoracle as po


# ------------------------------------------------------------------------------------------------ CPU: generators
def test_cer_pairs_are_consistent_csr_and_match_the_fixture_histogram():
    lab, lo, prd, pofs = BW.cer_pairs(20_000, 1)
    assert lo[0] == 0 and pofs[0] == 0 and lo[-1] == len(lab) and pofs[-1] == len(prd)
    assert (np.diff(lo) >= 1).all() and (np.diff(lo) <= 16).all() and (np.diff(pofs) >= 0).all()
    sl, sp = BW.csr_to_strings(lab, lo), BW.csr_to_strings(prd, pofs)
    assert [len(s) for s in sl] == np.diff(lo).tolist() and [len(s) for s in sp] == np.diff(pofs).tolist()
    assert all(c in BB.CHAR_SET for s in sl[:500] for c in s)
    _, _, dist, cer = po.compare_labels(sp, sl, return_all=True)
    exact = float((dist == 0).mean())
    assert 0.5 < exact < 0.75                       # the POS fixture: ~55 % of the strips are read correctly
    assert 0.02 < cer.mean() < 0.5
    # the host-side CSR encoder of the product path restates ord() per character
    from qeb_b200.mirror import utils as qutils
    f, o = qutils._encode_csr(sl + ["", "€uro", "a\ud800b"])
    f2, o2 = po.encode_csr(sl + ["", "€uro", "a\ud800b"])
    assert np.array_equal(f, f2) and np.array_equal(o, o2)


def test_label_generators_are_seeded_and_feasible():
    assert BW.vgg_labels(64, 3) == BW.vgg_labels(64, 3)
    for w in BW.vgg_labels(500, 4):
        assert 1 <= len(w) <= 23 and len(w) + sum(a == b for a, b in zip(w, w[1:])) <= 31 and w.isalnum()
    gt = BW.vgg_labels(2000, 5)
    ocr = BW.ocr_strings(gt, 6)
    assert ocr == BW.ocr_strings(gt, 6) and all(len(s) >= 1 for s in ocr)
    same = sum(a == b for a, b in zip(gt, ocr)) / len(gt)
    assert 0.5 < same < 0.62


@pytest.mark.parametrize("workload,unit", [("jitter_step", "patches/s"), ("cer_topk", "pairs/s")])
def test_reference_arm_of_the_other_workloads_keeps_the_contract(workload, unit):
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", workload, "--steps", "1",
                        "--warmup", "1"], cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == unit and d["higher_is_better"] is True and d["value"] > 0
    for k in ("metric", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["config"]["workload"].startswith({"jitter_step": "configs[2]", "cer_topk": "configs[4]"}[workload])
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_workload_flag_is_part_of_the_cli():
    a = BB.parse_args(["--workload", "area_step", "--gpus", "2"])
    assert a.workload == "area_step" and a.gpus == 2 and BB.parse_args([]).workload == "prep_step"
    assert set(BW.RUNNERS) == set(BB.WORKLOADS) - {"prep_step"}


# ------------------------------------------------------------------------------------------------ GPU
DEV = "cuda"


def cos(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return float((a @ b) / (a.norm() * b.norm()).clamp_min(1e-30))


def rel(a, b):
    return float((a.double().cpu() - b.double().cpu()).norm() / b.double().cpu().norm().clamp_min(1e-30))


@pytest.mark.gpu
def test_jitter_step_512_patches_vs_oracle():
    """configs[2] at its own size on one GPU: 64 base patches x 8 noised copies, one CRNN train-mode call per copy (its own
    batch statistics), CTC(mean) against that copy's OCR strings, gradients summed over the copies, one Adam step with
    weight decay. The oracle replays the kernel's own noise through the reference arithmetic and the torch graphs."""
    import qeb_b200  # noqa: F401
    from qeb_b200.mirror import ctc as qctc, train_ops, transform_helper as th
    from qeb_b200.mirror.models.model_crnn import CRNN
    torch.manual_seed(3)
    m = CRNN(95, False).to(DEV)
    m.register_backward_hook(m.backward_hook)
    mr = copy.deepcopy(m)
    m.train(); mr.train()
    opt = train_ops.Adam(m.parameters(), lr=BW.LR_CRNN, weight_decay=BW.WD_PATCH)
    optr = torch.optim.Adam(mr.parameters(), lr=BW.LR_CRNN, weight_decay=BW.WD_PATCH)
    base, labels = BB.synth_batch(64, 7)
    base = base.to(DEV)
    c2i = {c: i for i, c in enumerate(BB.CHAR_SET)}
    il = torch.full((64,), 31, dtype=torch.int32)
    g = torch.Generator().manual_seed(5)
    la_sum = lb_sum = 0.0
    for c in range(BW.INNER_LIMIT):
        sig = (torch.randint(0, 6, (64,), generator=g).double() / 100 + 1e-13).float()
        noisy, noise = th.jitter_batch(base, sig, seed=1000 + c, return_noise=True)
        assert torch.equal(noisy, torch.clamp(base - noise, 0, 1))                      # transform_helper.py:40-41, bit-exact
        y, ys = BB.encode(BW.ocr_strings(labels, 100 + c), c2i)
        la = qctc.CTCLoss()(m(noisy), y, il, ys)
        lb = torch.nn.CTCLoss()(nn_oracle.crnn_forward(mr, noisy), y.to(DEV), il.to(DEV), ys.to(DEV))
        assert abs(float(la) - float(lb)) < 1e-3 * abs(float(lb)), c
        la.backward(); lb.backward()                                                    # inside the loop: gradients accumulate
        la_sum += float(la); lb_sum += float(lb)
    assert abs(la_sum - lb_sum) < 1e-3 * abs(lb_sum)
    for (n, p), (_, r) in zip(m.named_parameters(), mr.named_parameters()):
        if n in ("convo.conv5.bias", "convo.conv6.bias"):      # analytically zero (bias before a train-mode BatchNorm)
            assert float(p.grad.abs().max()) < 1e-3 * float(m.convo.conv5.weight.grad.abs().max()) + 1e-4
            continue
        assert cos(p.grad, r.grad) > 0.999 and rel(p.grad, r.grad) < 0.04, (n, cos(p.grad, r.grad), rel(p.grad, r.grad))
    opt.step(); optr.step()
    for (n, p), (_, r) in zip(m.named_parameters(), mr.named_parameters()):
        assert float((p - r).abs().max()) <= 2.1e-4, n                                  # one Adam step moves a weight by ~lr
    assert int(m.convo.batchnorm1.num_batches_tracked) == BW.INNER_LIMIT                # one running-stat update per copy
    assert rel(m.convo.batchnorm1.running_var, mr.convo.batchnorm1.running_var) < 1e-3


@pytest.mark.gpu
def test_cer_topk_one_million_pairs_in_full_vs_c_oracle():
    """configs[4] at its own size, EVERY distance and fp64 CER against oracle/oracle.c, every segmented selection and the
    dataset-wide top-k against numpy stable sorts (bit-exact integer path)."""
    import qeb_b200  # noqa: F401
    from qeb_b200 import _lib
    from qeb_b200.mirror import selection_utils, utils as qutils
    n, seg = BW.N_PAIRS, BW.SEG
    lab, lo, prd, pofs = BW.cer_pairs(n, 42)
    dist_o, cer_o, tot = np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.float64), ctypes.c_double(0.0)
    p32 = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))
    correct = po.lib().oracle_compare_labels(p32(prd), p32(pofs), p32(lab), p32(lo), n, p32(dist_o),
                                             cer_o.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), ctypes.byref(tot))
    dl, dlo, dp, dpo = (torch.from_numpy(v).to(DEV) for v in (lab, lo, prd, pofs))
    dist, cer = qutils.cer_batch(dl, dlo, None, dp, dpo, None, n, int(max(np.diff(lo).max(), np.diff(pofs).max())))
    assert np.array_equal(dist.cpu().numpy(), dist_o)                    # all 1,000,000 distances
    assert np.array_equal(cer.cpu().numpy(), cer_o)                      # fp64, bit-exact
    assert int((dist == 0).sum()) == correct
    # the same through the string interface (host CSR encoding + H2D + D2H), and through uint8 char-set indices
    sl, sp = BW.csr_to_strings(lab, lo), BW.csr_to_strings(prd, pofs)
    d2, c2 = qutils.levenshtein_strings(sp, sl)
    assert np.array_equal(d2, dist_o) and np.array_equal(c2, cer_o)
    lut = np.zeros(0x20AD, dtype=np.uint8)
    lut[[ord(c) for c in BB.CHAR_SET]] = np.arange(len(BB.CHAR_SET), dtype=np.uint8)
    d8, c8 = qutils.cer_batch(torch.from_numpy(lut[lab]).to(DEV), dlo, None, torch.from_numpy(lut[prd]).to(DEV), dpo, None, n, 20)
    assert np.array_equal(d8.cpu().numpy(), dist_o) and np.array_equal(c8.cpu().numpy(), cer_o)
    # TopKCER over 15,625 minibatches of 64 (k = 32 and the 5 % setting) and the dataset-wide top 10 %
    c32 = cer.to(torch.float32)
    v = cer_o.astype(np.float32).reshape(-1, seg)
    n_seg = n // seg
    seg_off = torch.arange(0, n + 1, seg, dtype=torch.int32, device=DEV)
    for k in (32, 4):
        ks = torch.full((n_seg,), k, dtype=torch.int32, device=DEV)
        oo = torch.arange(0, n_seg * k, k, dtype=torch.int32, device=DEV)
        out = torch.empty(n_seg * k, dtype=torch.int64, device=DEV)
        _lib.call("qeb_cer_topk_segmented", c32.data_ptr(), seg_off.data_ptr(), ks.data_ptr(), oo.data_ptr(), n_seg, out.data_ptr(), _lib.stream())
        assert np.array_equal(out.view(n_seg, k).cpu().numpy(), np.argsort(-v, axis=1, kind="stable")[:, :k])
    top = selection_utils.topk_global_device(c32, n // 10).cpu().numpy()
    assert np.array_equal(top, np.argsort(-v.reshape(-1), kind="stable")[: n // 10])


@pytest.mark.gpu
def test_area_minibatch_phases_a_b_c_vs_oracle():
    """configs[3]: one minibatch of train_nn_area.py:212-304 through the mirror modules against the same statements on the
    oracle graphs: the TopKCER selection and the decoded strings / CERs / updated sampler dict must be IDENTICAL, the two
    losses within 1e-3."""
    import qeb_b200  # noqa: F401
    from qeb_b200.mirror import ctc as qctc, selection_utils, train_ops, transform_helper as th, utils as qutils
    from qeb_b200.mirror.models.model_crnn import CRNN
    from qeb_b200.mirror.models.model_unet import UNet
    torch.manual_seed(42)
    prep, crnn = UNet().to(DEV), CRNN(95, False).to(DEV)
    crnn.register_backward_hook(crnn.backward_hook)
    prep_r, crnn_r = copy.deepcopy(prep), copy.deepcopy(crnn)
    B = 64
    x, _ = BB.synth_batch(B, 7)
    x = x.to(DEV)
    labels = BW.vgg_labels(B, 11)
    names = [f"s{i}_{l}" for i, l in enumerate(labels)]
    ocr_all = BW.ocr_strings(labels, 300)
    c2i = {c: i for i, c in enumerate(BB.CHAR_SET)}
    i2c = {i: c for i, c in enumerate(BB.CHAR_SET)}
    cers0 = {n: (i * 7 % 11) / max(1, len(l)) for i, (n, l) in enumerate(zip(names, labels))}
    sampler = selection_utils.datasampler_factory("topKCER")(dict(cers0))
    k = max(1, math.ceil(B * 0.5))
    il = lambda n: torch.tensor([31] * n, dtype=torch.int32)
    opt_c = train_ops.Adam(crnn.parameters(), lr=1e-4, weight_decay=0)
    opt_p = train_ops.Adam(prep.parameters(), lr=5e-5, weight_decay=0)
    opt_cr = torch.optim.Adam(crnn_r.parameters(), lr=1e-4, weight_decay=0)
    opt_pr = torch.optim.Adam(prep_r.parameters(), lr=5e-5, weight_decay=0)
    # ---- phase A
    crnn.train(); prep.eval(); crnn_r.train(); prep_r.eval()
    img_all = prep(x)
    img_all_r = nn_oracle.unet_forward(prep_r, x)
    assert float((img_all - img_all_r).abs().max()) < 2e-3
    sel, _, idx = sampler.query(img_all, labels, k, names)
    want = po.topk_query(np.array([cers0[n] for n in names], dtype=np.float32), k)
    assert np.array_equal(idx.numpy(), want)
    noisy, noise = th.add_noise(sel.detach(), th.AddGaussianNoice(std=5, is_stochastic=True, return_noise=True))
    assert torch.equal(noisy, torch.clamp(sel.detach() - noise, 0, 1))
    y, ys = BB.encode([ocr_all[i] for i in idx.tolist()], c2i)
    la = qctc.CTCLoss()(crnn(noisy), y, il(k), ys)
    lb = torch.nn.CTCLoss()(nn_oracle.crnn_forward(crnn_r, noisy), y.to(DEV), il(k).to(DEV), ys.to(DEV))
    assert abs(float(la) - float(lb)) < 1e-3 * abs(float(lb))
    la.backward(); lb.backward(); opt_c.step(); opt_cr.step()
    # ---- phase B
    for p_, c_ in ((prep, crnn), (prep_r, crnn_r)):
        p_.train(); c_.train(); c_.apply(qutils.set_bn_eval); p_.zero_grad(); c_.zero_grad()
    img = prep(x); scores = crnn(img)
    img_r = nn_oracle.unet_forward(prep_r, x); scores_r = nn_oracle.crnn_forward(crnn_r, img_r)
    yg, ysg = BB.encode(labels, c2i)
    la = qctc.CTCLoss()(scores, yg, il(B), ysg) + train_ops.mse_to_ones(img)
    lb = torch.nn.CTCLoss()(scores_r, yg.to(DEV), il(B).to(DEV), ysg.to(DEV)) + torch.nn.functional.mse_loss(img_r, torch.ones_like(img_r))
    assert abs(float(la) - float(lb)) < 1e-3 * abs(float(lb))
    la.backward(); lb.backward(); opt_p.step(); opt_pr.step()
    # ---- phase C: device decode + CER against pred_to_string + compare_labels of the oracle ON THE SAME scores
    tg = qctc.pack_targets(yg, il(B), ysg, DEV)
    codes, lens, dist, cer = qutils.decode_and_cer(scores.detach(), tg.tg, tg.offs, tg.tl, 24)
    preds = po.pred_to_string(scores.detach().cpu().numpy(), i2c)
    assert qutils.pred_to_string(scores.detach(), labels, i2c) == preds
    want_cers = [po.compare_labels([preds[i]], [labels[i]])[1] for i in range(B)]
    assert cer.tolist() == want_cers                                          # fp64, bit-exact
    sampler.update_cer(cer.tolist(), names)
    assert sampler.cers == dict(zip(names, want_cers)) and all(sampler.all_cers[n] == [c] for n, c in zip(names, want_cers))


@pytest.mark.gpu
def test_crnn_warmup_loop_matches_reference_loop():
    """configs[0] / train_crnn.py:154-183 packaged as mirror/train_crnn.train_epoch + validate: three training steps and one
    validation pass against the same statements on the oracle graph (torch CTCLoss / Adam, pred_to_string + compare_labels of
    the oracle)."""
    import qeb_b200  # noqa: F401
    from qeb_b200.mirror import train_crnn as tc, train_ops
    from qeb_b200.mirror.models.model_crnn import CRNN
    torch.manual_seed(8)
    m = CRNN(95, False).to(DEV)
    mr = copy.deepcopy(m)
    c2i = {c: i for i, c in enumerate(BB.CHAR_SET)}
    i2c = {i: c for i, c in enumerate(BB.CHAR_SET)}
    batches = []
    for s in range(3):
        x, labels = BB.synth_batch(16, 20 + s)
        batches.append((x, labels))                               # CPU images, as a DataLoader hands them over
    opt = train_ops.Adam(m.parameters(), lr=1e-4)
    optr = torch.optim.Adam(mr.parameters(), lr=1e-4)
    total, steps = tc.train_epoch(m, batches, opt, c2i, DEV)
    assert steps == 3
    mr.train()
    ref_total = 0.0
    for x, labels in batches:
        mr.zero_grad()
        scores = nn_oracle.crnn_forward(mr, x.to(DEV))
        y, ys = BB.encode(labels, c2i)
        loss = torch.nn.CTCLoss()(scores, y.to(DEV), torch.tensor([31] * 16, dtype=torch.int32, device=DEV), ys.to(DEV))
        loss.backward(); optr.step()
        ref_total += loss.item()
    assert abs(total - ref_total) < 1e-3 * abs(ref_total)
    for (n, p), (_, r) in zip(m.named_parameters(), mr.named_parameters()):
        assert float((p - r).abs().max()) <= 3 * 2.1e-4, n           # three Adam steps of ~lr each
    vloss, crt, cer, nb = tc.validate(m, batches[:2], c2i, i2c, DEV)
    assert nb == 2 and not m.training
    with torch.no_grad():
        want_crt, want_cer, want_loss = 0, 0, 0.0
        for x, labels in batches[:2]:
            scores = m(x.to(DEV))                                     # same weights: decode / CER parity is exact
            preds = po.pred_to_string(scores.cpu().numpy(), i2c)
            c, e = po.compare_labels(preds, labels)
            want_crt += c; want_cer += e
            y, ys = BB.encode(labels, c2i)
            want_loss += float(torch.nn.functional.ctc_loss(scores, y.to(DEV), torch.tensor([31] * 16, dtype=torch.int32, device=DEV), ys.to(DEV)))
    assert (crt, cer) == (want_crt, want_cer)
    assert abs(vloss - want_loss) < 1e-4 * abs(want_loss)
