"""TEST INFRASTRUCTURE ONLY - writes tests/golden/* by running the UNMODIFIED reference (imported from
/root/reference through oracle/refload.py) in the build container. Run:  python -m oracle.gen_golden

The fixtures pin the Python-level reference functions on the hot path (the reference itself has no tests):
  decode.npz      utils.pred_to_string                         utils.py:74-92
  cer.json        utils.compare_labels (Levenshtein stubbed by the oracle restatement - distance itself unpinned)
  select.npz      TopKCERSampler.query / CerRangeSampler.query selection_utils.py:107-151 (real fixture CERs)
  jitter.npz      AddGaussianNoice                             transform_helper.py:26-45
  crop.npz        get_text_stack / padder (+autograd)          utils.py:118-141
  ctc.npz         torch.nn.CTCLoss CPU, reference call shapes  train_nn_area.py:146-148,174
  crnn.npz        models.model_crnn.CRNN fwd + CTC + bwd       models/model_crnn.py, train_crnn.py:157-162
  unet.npz        models.model_unet.UNet fwd + MSE + bwd       models/model_unet.py, train_nn_area.py:173-182
  tracking.json / tracking.npz  tracking_utils.generate_ctc_target_batches / weighted_ctc_loss (both weighting modes,
                  value and gradient at the scores) and LevenshteinWeightGenerator.gen_weights
                                                               tracking_utils.py:34-75, label_tracking/tracking_methods.py:72-101
  pruning.json    pruning/methods.topk on the reference's own artifacts (pruning/cer_artifacts/cers_pos*.json)
"""
import json
import os
import random

import numpy as np
import torch

from . import refload

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def seed_all(s):
    # utils.py:240-243 set_random_seeds
    torch.manual_seed(s)
    random.seed(s)
    np.random.seed(s)


def pos_labels(ref, n, seed):
    """Real POS label strings: the label field of the keys of cer_data_utils/pos_dataset_cers.json."""
    with open(os.path.join(refload.REFERENCE_ROOT, "cer_data_utils", "pos_dataset_cers.json")) as f:
        cers = json.load(f)
    keys = list(cers.keys())
    rng = random.Random(seed)
    picked = rng.sample(keys, n)
    labels = [k.split("_")[1] for k in picked]
    return picked, labels, cers


def gen_decode(ref):
    seed_all(42)
    V = len(ref.properties.char_set)
    _, index_to_char, _ = ref.utils.get_char_maps(ref.properties.char_set)
    T, B = 31, 24
    scores = torch.randn(T, B, V)
    # half of the batch: a structured path with repeats and blanks, so collapse/drop-blank are exercised
    g = torch.Generator().manual_seed(7)
    for b in range(B // 2):
        path = torch.randint(0, 6, (T,), generator=g)
        path = torch.where(path < 2, torch.zeros_like(path), path + 30 * (b % 3))
        for t in range(1, T):
            if torch.rand((), generator=g) < 0.4:
                path[t] = path[t - 1]
        scores[torch.arange(T), b, path] += 8.0
    scores[3, 5, :] = 0.25  # exact tie over all classes -> first index (blank)
    scores[4, 5, 10] = scores[4, 5, 20] = 9.0  # two-way tie -> lower index
    strings = ref.utils.pred_to_string(scores, [""] * B, index_to_char)
    np.savez_compressed(os.path.join(OUT, "decode.npz"), scores=scores.numpy(),
                        strings=np.array(strings, dtype=object), allow_pickle=True)
    return strings


def gen_cer(ref):
    names, labels, _ = pos_labels(ref, 200, 3)
    rng = random.Random(5)
    alphabet = ref.properties.char_set[1:]
    preds = []
    for l in labels:
        s = list(l)
        r = rng.random()
        if r < 0.55:
            pass
        else:
            for _ in range(rng.randint(1, 4)):
                op = rng.random()
                if op < 0.33 and s:
                    del s[rng.randrange(len(s))]
                elif op < 0.66:
                    s.insert(rng.randrange(len(s) + 1), rng.choice(alphabet))
                elif s:
                    s[rng.randrange(len(s))] = rng.choice(alphabet)
        preds.append("".join(s))
    labels += ["", "", "abc", "€10", "kitten", "flaw", "a" * 40]
    preds += ["", "xyz", "", "10€", "sitting", "lawn", "b" * 25]
    correct, total = ref.utils.compare_labels(preds, labels)
    per_pair = [ref.utils.compare_labels([p], [l]) for p, l in zip(preds, labels)]
    with open(os.path.join(OUT, "cer.json"), "w") as f:
        json.dump({"labels": labels, "preds": preds, "correct": correct, "total_cer": total,
                   "per_pair": per_pair,
                   "note": "Levenshtein.distance is the oracle restatement (python-Levenshtein 0.12.0 not installable)"},
                  f, ensure_ascii=False)


def gen_select(ref):
    su = ref.selection_utils
    names, labels, cers = pos_labels(ref, 64 * 6, 11)
    out = {}
    seg = 0
    # real fixture CERs (tie heavy: 55% zeros)
    for i in range(6):
        nm = names[64 * i: 64 * (i + 1)]
        lab = labels[64 * i: 64 * (i + 1)]
        imgs = torch.arange(64).float().reshape(64, 1)
        vals = torch.tensor([cers[n] for n in nm])
        for prop in (0.5, 0.87, 0.95):
            k = max(1, int(np.ceil(64 * (1 - prop))))
            # TopK: stable order shim for ties, plus the reference call itself for the value multiset
            sampler = su.TopKCERSampler(dict(cers))
            _, _, idx_ref = sampler.query(imgs, lab, k, nm)
            idx_stable = torch.argsort(vals, descending=True, stable=True)[:k]
            assert torch.equal(vals[idx_ref].sort().values, vals[idx_stable].sort().values)
            out[f"topk_{seg}_vals"] = vals.numpy()
            out[f"topk_{seg}_k"] = np.int64(k)
            out[f"topk_{seg}_idx_stable"] = idx_stable.numpy()
            out[f"topk_{seg}_ref_values"] = vals[idx_ref].numpy()
            # range sampler: reference call under a known host seed; the draws are re-derived by reseeding
            sampler = su.CerRangeSampler(dict(cers))
            torch.manual_seed(100 + seg)
            _, _, ridx = sampler.query(imgs, lab, k, nm)
            torch.manual_seed(100 + seg)
            rands = torch.rand(k)
            out[f"range_{seg}_rands"] = rands.numpy()
            out[f"range_{seg}_idx"] = ridx.numpy()
            seg += 1
    # tie-free synthetic segments of odd sizes, k > n included for the range sampler
    g = torch.Generator().manual_seed(1)
    for n, k in ((1, 1), (7, 3), (33, 40), (124, 62)):
        vals = torch.rand(n, generator=g) * 3
        d = {f"s{j}": float(vals[j]) for j in range(n)}
        nm = list(d.keys())
        imgs = torch.arange(n).float().reshape(n, 1)
        _, _, idx_ref = su.TopKCERSampler(d).query(imgs, nm, min(k, n), nm)
        out[f"topk_{seg}_vals"] = torch.tensor([d[x] for x in nm]).numpy()
        out[f"topk_{seg}_k"] = np.int64(min(k, n))
        out[f"topk_{seg}_idx_stable"] = idx_ref.numpy()  # tie-free: the reference order is the stable order
        out[f"topk_{seg}_ref_values"] = torch.tensor([d[x] for x in nm])[idx_ref].numpy()
        torch.manual_seed(200 + seg)
        _, _, ridx = su.CerRangeSampler(d).query(imgs, nm, k, nm)
        torch.manual_seed(200 + seg)
        out[f"range_{seg}_rands"] = torch.rand(k).numpy()
        out[f"range_{seg}_idx"] = ridx.numpy()
        seg += 1
    out["n_seg"] = np.int64(seg)
    np.savez_compressed(os.path.join(OUT, "select.npz"), **out)


def gen_jitter(ref):
    seed_all(42)
    imgs = torch.rand(3, 1, 32, 128)
    out = {"imgs": imgs.numpy()}
    noiser = ref.transform_helper.AddGaussianNoice(std=5, is_stochastic=True, return_noise=True)
    for i, coef in enumerate((1, 1, 0.5)):
        o, z = noiser(imgs[i], coef)
        out[f"out_{i}"] = o.numpy()
        out[f"noise_{i}"] = z.numpy()
        out[f"coef_{i}"] = np.float32(coef)
    np.savez_compressed(os.path.join(OUT, "jitter.npz"), **out)


def gen_crop(ref):
    seed_all(42)
    H, W = 80, 200
    img = torch.rand(1, H, W, requires_grad=True)
    boxes = [(10, 5, 90, 30), (0, 0, 127, 31), (150, 60, 200, 80), (20, 20, 21, 21), (60, 40, 180, 70), (100, 10, 100, 30),
             (5, 50, 133, 79)]
    labels = [{"label": f"l{i}", "x_min": b[0], "y_min": b[1], "x_max": b[2], "y_max": b[3]} for i, b in enumerate(boxes)]
    stack, labs = ref.utils.get_text_stack(img, labels, (32, 128))
    w = torch.rand_like(stack)
    (stack * w).sum().backward()
    np.savez_compressed(os.path.join(OUT, "crop.npz"), img=img.detach().numpy(), boxes=np.array(boxes, dtype=np.int32),
                        out=stack.detach().numpy(), w=w.numpy(), grad=img.grad.numpy())


def encode(labels, char_to_index):
    y = [char_to_index[c] for c in "".join(labels)]
    return torch.tensor(y, dtype=torch.int), torch.tensor([len(l) for l in labels], dtype=torch.int)


def gen_ctc(ref):
    seed_all(42)
    V = len(ref.properties.char_set)
    char_to_index, _, _ = ref.utils.get_char_maps(ref.properties.char_set)
    T, B = 31, 12
    _, labels, _ = pos_labels(ref, B, 21)
    labels[0] = "aabbcc"          # repeated characters
    labels[1] = ""                # zero-length target
    labels[2] = "A" * 40          # infeasible: longer than T -> inf loss, NaN grads
    labels[3] = "hello world!!"
    labels[4] = "1111111111111111"  # 16 repeats need 31 frames exactly
    y, y_size = encode(labels, char_to_index)
    out = {"targets": y.numpy(), "target_lengths": y_size.numpy(), "labels": np.array(labels, dtype=object)}
    logits = torch.randn(T, B, V) * 2
    lp0 = torch.log_softmax(logits, 2)
    out["log_probs"] = lp0.numpy()
    in_len = torch.tensor([T] * B, dtype=torch.int)
    for red in ("mean", "none"):
        lp = lp0.clone().requires_grad_(True)
        loss = torch.nn.CTCLoss(reduction=red)(lp, y, in_len, y_size)
        (loss if red == "mean" else (loss * torch.arange(1, B + 1).float()).sum()).backward()
        out[f"loss_{red}"] = loss.detach().numpy()
        out[f"grad_{red}"] = lp.grad.numpy()
    # feasible-only subset + zero_infinity + ragged input lengths
    in_len2 = torch.tensor([T - (b % 5) for b in range(B)], dtype=torch.int)
    lp = lp0.clone().requires_grad_(True)
    loss = torch.nn.CTCLoss(reduction="mean", zero_infinity=True)(lp, y, in_len2, y_size)
    loss.backward()
    out["input_lengths_ragged"] = in_len2.numpy()
    out["loss_zero_inf_ragged"] = loss.detach().numpy()
    out["grad_zero_inf_ragged"] = lp.grad.numpy()
    # weighted_ctc_loss subset form: scores[:, idx, :] (tracking_utils.py:65-68)
    idx = [7, 3, 9, 0]
    sub_labels = [labels[i] for i in idx]
    ys, ys_size = encode(sub_labels, char_to_index)
    lp = lp0.clone().requires_grad_(True)
    loss = torch.nn.CTCLoss()(lp[:, idx, :], ys, in_len[idx], ys_size)
    loss.backward()
    out["subset_idx"] = np.array(idx, dtype=np.int32)
    out["subset_targets"] = ys.numpy()
    out["subset_target_lengths"] = ys_size.numpy()
    out["subset_loss"] = loss.detach().numpy()
    out["subset_grad"] = lp.grad.numpy()
    np.savez_compressed(os.path.join(OUT, "ctc.npz"), **out, allow_pickle=True)


def param_digest(module):
    d = {}
    for k, v in module.state_dict().items():
        v = v.detach().double()
        d[k] = [float(v.sum()), float(v.abs().sum())]
    return d


def grad_digest(module):
    d = {}
    for k, p in module.named_parameters():
        g = p.grad
        d[k] = {"norm": float(g.double().norm()), "head": g.reshape(-1)[:8].tolist()}
    return d


def gen_crnn(ref):
    V = len(ref.properties.char_set)
    char_to_index, _, _ = ref.utils.get_char_maps(ref.properties.char_set)
    seed_all(42)
    model = ref.model_crnn.CRNN(V, False)
    model.register_backward_hook(model.backward_hook)
    digest0 = param_digest(model)
    B = 4
    _, labels, _ = pos_labels(ref, B, 33)
    labels[1] = "B" * 35  # infeasible sample: exercises the NaN hook
    x = torch.rand(B, 1, 32, 128)
    y, y_size = encode(labels, char_to_index)
    model.train()
    scores = model(x)
    in_len = torch.tensor([scores.shape[0]] * B, dtype=torch.int)
    loss = torch.nn.CTCLoss()(scores, y, in_len, y_size)
    loss.backward()
    out = {"x": x.numpy(), "targets": y.numpy(), "target_lengths": y_size.numpy(),
           "scores_train": scores.detach().numpy(), "loss_train": loss.detach().numpy(),
           "bn1_mean": model.convo.batchnorm1.running_mean.numpy().copy(),
           "bn1_var": model.convo.batchnorm1.running_var.numpy().copy(),
           "grad_lstm_w_ih_l0": model.lstm.weight_ih_l0.grad.numpy()[:16].copy(),
           "grad_conv7_w": model.convo.conv7.weight.grad.numpy()[:2].copy(),
           "grad_conv1_w": model.convo.conv1.weight.grad.numpy().copy(),
           "grad_linear_b": model.linear.bias.grad.numpy().copy()}
    gd = grad_digest(model)
    # phase-B mode: train() + set_bn_eval, grad w.r.t. the input image (train_nn_area.py:277-286)
    model.zero_grad()
    model.train()
    model.apply(ref.utils.set_bn_eval)
    xg = x.clone().requires_grad_(True)
    labels2 = list(labels)
    labels2[1] = "ok"
    y2, y2_size = encode(labels2, char_to_index)
    scores2 = model(xg)
    loss2 = torch.nn.CTCLoss()(scores2, y2, in_len, y2_size)
    loss2.backward()
    out.update({"targets_b": y2.numpy(), "target_lengths_b": y2_size.numpy(), "scores_bneval": scores2.detach().numpy(),
                "loss_bneval": loss2.detach().numpy(), "grad_x_bneval": xg.grad.numpy()})
    gd2 = grad_digest(model)
    np.savez_compressed(os.path.join(OUT, "crnn.npz"), **out)
    with open(os.path.join(OUT, "crnn_digest.json"), "w") as f:
        json.dump({"labels": labels, "labels_b": labels2, "param_digest_seed42": digest0, "grad_digest_train": gd,
                   "grad_digest_bneval": gd2}, f)


def gen_unet(ref):
    seed_all(42)
    model = ref.model_unet.UNet()
    digest0 = param_digest(model)
    x = torch.rand(2, 1, 32, 128)
    model.train()
    y = model(x)
    loss = torch.nn.MSELoss()(y, torch.ones(y.shape))
    loss.backward()
    out = {"x": x.numpy(), "y_train": y.detach().numpy(), "loss_train": loss.detach().numpy(),
           "enc1norm1_mean": model.encoder1.enc1norm1.running_mean.numpy().copy(),
           "enc1norm1_var": model.encoder1.enc1norm1.running_var.numpy().copy(),
           "grad_enc1conv1_w": model.encoder1.enc1conv1.weight.grad.numpy().copy(),
           "grad_upconv4_b": model.upconv4.bias.grad.numpy().copy(),
           "grad_conv_w": model.conv.weight.grad.numpy().copy()}
    gd = grad_digest(model)
    model.eval()
    with torch.no_grad():
        out["y_eval"] = model(x).numpy()
    np.savez_compressed(os.path.join(OUT, "unet.npz"), **out)
    with open(os.path.join(OUT, "unet_digest.json"), "w") as f:
        json.dump({"param_digest_seed42": digest0, "grad_digest_train": gd}, f)


def gen_tracking(ref):
    """Label-tracking CTC path: a trainer-like object with a label history per image, the reference's target batching,
    the Levenshtein loss weights and the weighted multi-target CTC loss in both weighting modes."""
    import importlib
    import types
    seed_all(42)
    tu = ref.tracking_utils
    tm = importlib.import_module("label_tracking.tracking_methods")
    char_to_index, _, V = ref.utils.get_char_maps(ref.properties.char_set)
    names, labels, _ = pos_labels(ref, 24, 5)
    rng = random.Random(3)
    alphabet = [c for c in ref.properties.char_set[1:] if c != "`"]

    def corrupt(s):
        s = list(s)
        for _ in range(rng.randint(0, 2)):
            op = rng.choice("sid")
            pos = rng.randrange(len(s) + 1)
            if op == "s" and s:
                s[min(pos, len(s) - 1)] = rng.choice(alphabet)
            elif op == "i" and len(s) < 14:
                s.insert(pos, rng.choice(alphabet))
            elif op == "d" and len(s) > 1:
                s.pop(min(pos, len(s) - 1))
        return "".join(s)

    window = 4
    tracked = {}
    for n_, l_ in zip(names[:20], labels[:20]):          # 4 images have no history at all
        tracked[n_] = [corrupt(l_) for _ in range(rng.randint(1, 6))]
    tracked[names[0]] = ["", corrupt(labels[0]), ""]       # empty OCR labels occur (tess returns '' on blank crops)
    batch_names = names[:16] + names[20:22]                # 16 with history + 2 without -> gen_weights skips those
    T, B = 31, 16
    scores = torch.randn(T, B, V).log_softmax(2).requires_grad_(True)
    pred_size = torch.tensor([T] * B, dtype=torch.int)
    out_np, out_js = {"scores": scores.detach().numpy()}, {"tracked": tracked, "names": batch_names, "window": window}
    for mode in ("levenshtein", "decaying"):
        obj = types.SimpleNamespace(window_size=window, tracked_labels=tracked, char_to_index=char_to_index, device="cpu",
                                    weightgen_method=mode, primary_loss_fn=torch.nn.CTCLoss(),
                                    primary_loss_fn_sample_wise=torch.nn.CTCLoss(reduction="none"))
        tb = tu.generate_ctc_target_batches(obj, batch_names[:B])
        args = types.SimpleNamespace(window_size=window, decay_factor=0.7)
        gen = tm.weightgenerator_factory(mode)(args, "cpu", char_to_index)
        w = gen.gen_weights(tracked, batch_names) if mode == "levenshtein" else gen.gen_weights(obj, batch_names)
        if mode == "levenshtein":
            out_np["lev_weights"] = w.numpy()
            out_js["target_batches"] = [[t.tolist(), ts.tolist(), idx] for t, ts, idx in tb]
            # the trainers index the weights of the images in the CTC batch; column 0 belongs to the current label
            lw = w[:B, 1:]
        else:
            out_np["decay_weights"] = w.numpy()
            lw = w
        if scores.grad is not None:
            scores.grad = None
        loss = tu.weighted_ctc_loss(obj, scores, pred_size, tb, lw)
        loss.backward()
        out_np[f"loss_{mode}"] = loss.detach().numpy()
        out_np[f"grad_{mode}"] = scores.grad.numpy().copy()
    np.savez_compressed(os.path.join(OUT, "tracking.npz"), **out_np)
    with open(os.path.join(OUT, "tracking.json"), "w") as f:
        json.dump(out_js, f)


def gen_pruning(ref):
    """The reference's own pruning artifacts (pruning/cer_artifacts/*.json, written by pruning/prune_dataset.py with
    pruning/methods.topk): input name -> mean CER and the kept names, in order, at 10 % and 50 % pruning. These ARE golden
    vectors held by the reference (data files); the live function is called as well."""
    import importlib
    import sys
    import types
    art = os.path.join(refload.REFERENCE_ROOT, "pruning", "cer_artifacts")
    cers = json.load(open(os.path.join(art, "cers_pos.json")))
    out = {"cers": cers}
    sys.modules.setdefault("apricot", types.SimpleNamespace(FacilityLocationSelection=None))   # methods.py imports it at the top
    sys.path.insert(0, os.path.join(refload.REFERENCE_ROOT, "pruning"))
    methods = importlib.import_module("methods")
    for pct in (10, 50):
        kept = json.load(open(os.path.join(art, f"cers_pos_topk_{pct}.json")))
        live = methods.topk(cers, len(kept))
        assert list(live.keys()) == list(kept.keys()) and live == kept
        out[f"topk_{pct}_names"] = list(kept.keys())
    with open(os.path.join(OUT, "pruning.json"), "w") as f:
        json.dump(out, f)


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = refload.load()
    torch.set_num_threads(8)
    gen_decode(ref)
    gen_cer(ref)
    gen_select(ref)
    gen_jitter(ref)
    gen_crop(ref)
    gen_ctc(ref)
    gen_crnn(ref)
    gen_unet(ref)
    gen_tracking(ref)
    gen_pruning(ref)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
