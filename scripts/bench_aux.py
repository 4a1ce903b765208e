#!/usr/bin/env python
"""Device-time benchmark of the integer / elementwise kernels beside the training step (BASELINE.json configs 3 and 5,
SURVEY.md 8(d)): batched Levenshtein + CER over 1 M pairs, segmented TopK / range selection over 1 M CERs, Gaussian jitter
of 512 patches, greedy decode and CTC forward/backward of 512 sequences. CUDA events around back-to-back launches with
inputs resident in HBM; algorithmic bytes per SURVEY.md 8(d); peak = measured copy bandwidth (MEASURED_PEAKS.json).
Prints one JSON object; `python scripts/bench_aux.py > profiles/aux_kernels_rNN.json`."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import qeb_b200  # noqa: F401
from qeb_b200 import _lib
from qeb_b200.mirror import ctc as qctc, transform_helper as th, utils as qutils

DEV = torch.device("cuda", 0)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0)
except OSError:
    PEAK = 6650.0


def timed(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def row(ms, bytes_alg, units, unit_name):
    gbs = bytes_alg / ms / 1e6
    return {"ms": ms, "algorithmic_MB": bytes_alg / 1e6, "GB/s": gbs, "frac_of_hbm_peak": gbs / PEAK, f"{unit_name}/s": units / (ms / 1e3)}


def main():
    out = {"hbm_peak_gbs": PEAK}
    rng = np.random.default_rng(0)
    # ---- Levenshtein + CER, 1 M POS-like pairs (length 1..16, 55 % identical, others 1-3 edits)
    n = 1_000_000
    la = rng.integers(1, 17, n)
    a = rng.integers(1, 95, int(la.sum())).astype(np.uint8)
    aoff = np.concatenate([[0], np.cumsum(la)]).astype(np.int32)
    b = a.copy()
    edit = rng.random(len(a)) < 0.08
    b[edit] = rng.integers(1, 95, int(edit.sum())).astype(np.uint8)
    for sym_dtype, name in ((np.uint8, "levenshtein_cer_1M_u8"), (np.int32, "levenshtein_cer_1M_utf32")):
        ta, tb = torch.from_numpy(a.astype(sym_dtype)).to(DEV), torch.from_numpy(b.astype(sym_dtype)).to(DEV)
        to = torch.from_numpy(aoff).to(DEV)
        ms = timed(lambda: qutils.cer_batch(ta, to, None, tb, to, None, n, 16))
        s = np.dtype(sym_dtype).itemsize
        alg = 2 * len(a) * s + 2 * 4 * n + 4 * n + 8 * n            # symbols + offsets + int32 distance + fp64 CER
        r = row(ms, alg, n, "pairs")
        r["dp_cells/s"] = float((la.astype(np.int64) ** 2).sum()) / (ms / 1e3)
        out[name] = r
    # ---- segmented TopK over 1 M CERs (15,625 minibatches of 64), k = 32 and k = 4
    n_seg, m = 15625, 64
    vals = torch.from_numpy((rng.integers(0, 18, (n_seg, m)) * (rng.random((n_seg, m)) > 0.55) / rng.integers(1, 17, (n_seg, m))).astype(np.float32)).to(DEV)
    seg_off = torch.arange(0, n_seg * m + 1, m, dtype=torch.int32, device=DEV)
    for k in (32, 4):
        ks = torch.full((n_seg,), k, dtype=torch.int32, device=DEV)
        oo = torch.arange(0, n_seg * k, k, dtype=torch.int32, device=DEV)
        res = torch.empty(n_seg * k, dtype=torch.int64, device=DEV)
        st = torch.cuda.current_stream().cuda_stream
        ms = timed(lambda: _lib.call("qeb_cer_topk_segmented", vals.data_ptr(), seg_off.data_ptr(), ks.data_ptr(), oo.data_ptr(), n_seg, res.data_ptr(), st))
        out[f"topk_segmented_1M_k{k}"] = row(ms, 4 * n_seg * m + 8 * n_seg * k + 12 * n_seg, n_seg * m, "cers")
    # ---- jitter, 512 patches (config 3: 64 x 8 noised copies)
    imgs = torch.rand(512, 1, 32, 128, device=DEV)
    sig = ((torch.arange(512) % 6).float() / 100 + 1e-13).to(DEV)
    ms = timed(lambda: th.jitter_batch(imgs, sig, seed=7))
    out["gauss_jitter_512"] = row(ms, 2 * imgs.numel() * 4, 512, "patches")
    # ---- greedy decode + CTC, 512 sequences
    T, B, V = 31, 512, 95
    lp = torch.randn(T, B, V, device=DEV).log_softmax(2)
    ms = timed(lambda: qutils.decode_batch(lp))
    out["greedy_decode_512"] = row(ms, T * B * V * 4 + B * 36, B, "sequences")
    tl = torch.from_numpy(rng.integers(1, 17, B).astype(np.int32))
    y = torch.from_numpy(rng.integers(1, V, int(tl.sum())).astype(np.int32))
    packed = qctc.pack_targets(y, torch.full((B,), T, dtype=torch.int32), tl, DEV)
    lpg = lp.clone().requires_grad_(True)
    loss_fn = qctc.CTCLoss()

    def ctc_fb():
        lpg.grad = None
        loss_fn(lpg, packed).backward()

    ms = timed(ctc_fb)
    out["ctc_fwd_bwd_512"] = row(ms, 2 * T * B * V * 4, B, "sequences")
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
