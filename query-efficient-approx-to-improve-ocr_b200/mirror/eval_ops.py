"""Validation / evaluation step of the trainers as one call.

The reference's per-epoch validation loops (train_nn_patch.py:366-388, train_nn_area.py:327-345) and eval_prep.py do, per
batch: preprocessor forward -> CRNN forward -> CTC + MSE loss -> pred_to_string (31*B `.item()` syncs) -> three
compare_labels calls (one Levenshtein C call per pair each): prediction vs ground truth, OCR vs ground truth,
prediction vs OCR. `validation_batch` returns the same numbers from: the two networks in inference mode (BatchNorm folded
into the conv epilogues, no activations saved for a backward), ONE greedy-decode launch and ONE batched Levenshtein
launch for all three comparisons. The OCR engine stays a host-side callable (`ocr_labels` are its strings).
"""
import torch

from . import ctc as qctc
from . import train_ops
from . import utils as qutils


def encode_labels(labels, char_to_index):
    """TrainNNPrep._call_model's label encoding (train_nn_area.py:163-171)."""
    y = torch.tensor([char_to_index[c] for l in labels for c in l], dtype=torch.int32)
    y_size = torch.tensor([len(l) for l in labels], dtype=torch.int32)
    return y, y_size


def compare_three_way(preds, labels, ocr_labels=None):
    """(crt, cer), (ocr_crt, ocr_cer), (matching_crt, matching_cer) exactly as three compare_labels calls
    (utils.py:95-110: exact-match count, CER summed in list order in float64) - with one Levenshtein launch."""
    n = len(labels)
    a, b = list(preds[:n]), list(labels)            # (predictions, denominators)
    if ocr_labels is not None:
        a += list(ocr_labels[:n]) + list(preds[:n])
        b += list(labels) + list(ocr_labels[:n])
    dist, cer = qutils.levenshtein_strings(a, b)
    out = []
    for k in range(len(a) // max(n, 1) if n else 0):
        d, c = dist[k * n:(k + 1) * n], cer[k * n:(k + 1) * n]
        total = 0
        for v in c.tolist():
            total += v
        out.append((int((d == 0).sum()), total))
    while len(out) < 3:
        out.append((0, 0))
    return tuple(out)


@torch.no_grad()
def validation_batch(prep_model, crnn_model, images, labels, char_to_index, index_to_char, ocr=None, loss_fn=None,
                     scalar=1.0):
    """One validation batch of train_nn_area.py:327-341. images: (B,1,32,W) on the models' device; labels: list[str];
    ocr: None or a callable images_cpu -> list[str] (the reference's self.ocr.get_labels).
    Returns dict(loss, preds, ocr_labels, crt, cer, ocr_crt, ocr_cer, matching_crt, matching_cer, img_preds)."""
    loss_fn = loss_fn or qctc.CTCLoss()
    img_preds = prep_model(images)
    scores = crnn_model(img_preds)
    y, y_size = encode_labels(labels, char_to_index)
    pred_size = torch.tensor([scores.shape[0]] * images.shape[0], dtype=torch.int32)
    loss = loss_fn(scores, y, pred_size, y_size) + scalar * train_ops.mse_to_ones(img_preds)
    preds = qutils.pred_to_string(scores, labels, index_to_char)
    ocr_labels = ocr(img_preds.cpu()) if ocr is not None else None
    (crt, cer), (ocr_crt, ocr_cer), (m_crt, m_cer) = compare_three_way(preds, labels, ocr_labels)
    return {"loss": float(loss), "preds": preds, "ocr_labels": ocr_labels, "crt": crt, "cer": cer, "ocr_crt": ocr_crt,
            "ocr_cer": ocr_cer, "matching_crt": m_crt, "matching_cer": m_cer, "img_preds": img_preds}
