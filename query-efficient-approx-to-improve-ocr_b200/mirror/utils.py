"""Hot-path helpers of the reference's utils.py with the same names, arguments and return types.

  get_char_maps   utils.py:22-40      (host logic, restated)
  pred_to_string  utils.py:74-92   -> qeb_greedy_decode   (one launch instead of T*B .item() syncs)
  compare_labels  utils.py:95-110  -> qeb_levenshtein_batch (one launch instead of one C call per pair)
  set_bn_eval     utils.py:113-115    (host logic, restated)
  get_text_stack  utils.py:118-141 -> qeb_crop_pad_gather / qeb_crop_pad_scatter (differentiable)
plus device-resident batch forms the trainers' phase C can use without host round trips
(decode_batch, cer_batch, decode_and_cer).
"""
import numpy as np
import torch

from .. import _lib


def get_char_maps(vocabulary=None):
    if vocabulary is None:
        vocab = ["-"] + [chr(ord("a") + i) for i in range(26)] + [chr(ord("A") + i) for i in range(26)] + \
                [chr(ord("0") + i) for i in range(10)]
    else:
        vocab = vocabulary
    char_to_index, index_to_char = {}, {}
    for cnt, c in enumerate(vocab):
        char_to_index[c] = cnt
        index_to_char[cnt] = c
    return char_to_index, index_to_char, len(vocab)


def set_bn_eval(module):
    if isinstance(module, torch.nn.modules.batchnorm._BatchNorm):
        module.eval()


def _require_cuda():
    if not torch.cuda.is_available():
        raise _lib.QebError("the qeb hot path needs a CUDA device (no CPU fallback)")


# ---------------------------------------------------------------------------------------------------------------
# greedy decode
def decode_batch(scores, blank=0, want_path=False):
    """scores (T,B,V) CUDA fp32 -> (codes (B,T) int32 padded with -1, lengths (B) int32), both on device.
    Scores that come straight out of the qeb CRNN carry the per-frame arg-max its head epilogue computed (`_qeb_path`, valid
    while the tensor has not been modified in place): then only that (T,B) int path is collapsed, the scores are not re-read."""
    attached = getattr(scores, "_qeb_path", None)
    if attached is not None and not want_path and blank >= 0 and attached[1] == scores._version and attached[0].shape == scores.shape[:2]:
        path = attached[0]
        T, B = path.shape
        codes = torch.empty((B, T), dtype=torch.int32, device=path.device)
        lens = torch.empty(B, dtype=torch.int32, device=path.device)
        _lib.call("qeb_greedy_collapse", path.data_ptr(), path.stride(0), path.stride(1), T, B, blank, codes.data_ptr(), lens.data_ptr(),
                  _lib.stream())
        return codes, lens
    if not scores.is_cuda:
        _require_cuda()
        scores = scores.cuda()
    scores = scores.float()
    if scores.stride(2) != 1:
        scores = scores.contiguous()
    T, B, V = scores.shape
    codes = torch.empty((B, T), dtype=torch.int32, device=scores.device)
    lens = torch.empty(B, dtype=torch.int32, device=scores.device)
    path = torch.empty((B, T), dtype=torch.int32, device=scores.device) if want_path else None
    _lib.call("qeb_greedy_decode", scores.data_ptr(), scores.stride(0), scores.stride(1), T, B, V, blank,
              codes.data_ptr(), lens.data_ptr(), _lib.ptr(path), _lib.stream())
    return (codes, lens, path) if want_path else (codes, lens)


def pred_to_string(scores, labels, index_to_char, show_text=False):
    codes, lens = decode_batch(scores)
    codes, lens = codes.cpu().numpy(), lens.cpu().numpy()
    preds = []
    for i in range(codes.shape[0]):
        out = "".join(index_to_char[int(c)] for c in codes[i, : lens[i]])
        preds.append(out)
        if show_text:
            print(labels[i], " -> ", out)
    return preds


# ---------------------------------------------------------------------------------------------------------------
# Levenshtein / CER
def _encode_csr(strings):
    """list[str] -> (int32 code points, int32 offsets[n+1]). One C-level UTF-32 encode of the joined strings instead of a
    Python loop over the characters (1 M strings: ~0.1 s instead of seconds); `surrogatepass` keeps lone surrogates as
    their code points, which is what ord() gives."""
    offs = np.zeros(len(strings) + 1, dtype=np.int32)
    if len(strings):
        np.cumsum(np.fromiter(map(len, strings), dtype=np.int64, count=len(strings)), out=offs[1:])
    flat = np.frombuffer(bytearray("".join(strings).encode("utf-32-le", "surrogatepass")), dtype="<i4")   # writable for torch
    return flat, offs


def cer_batch(a_syms, a_off, a_len, b_syms, b_off, b_len, n, max_len, want_cer=True):
    """Device-resident form. a = labels (CER denominator), b = predictions; int32 or uint8 symbols.
    Returns (dist int32 (n), cer float64 (n) or None) on device."""
    dev = a_off.device
    dist = torch.empty(n, dtype=torch.int32, device=dev)
    cer = torch.empty(n, dtype=torch.float64, device=dev) if want_cer else None
    scratch = torch.empty(1, dtype=torch.int32, device=dev)
    sym_bytes = a_syms.element_size()
    if b_syms.element_size() != sym_bytes or sym_bytes not in (1, 4):
        raise _lib.QebError("cer_batch: symbols must both be uint8 or both int32")
    _lib.call("qeb_levenshtein_batch", a_syms.data_ptr(), a_off.data_ptr(), _lib.ptr(a_len), b_syms.data_ptr(),
              b_off.data_ptr(), _lib.ptr(b_len), n, sym_bytes, max_len, dist.data_ptr(), _lib.ptr(cer), scratch.data_ptr(),
              _lib.stream())
    return dist, cer


def levenshtein_strings(preds, labels, device=None):
    """Batched Levenshtein.distance(labels[i], preds[i]) and distance / max(1, len(labels[i])).
    Returns (dist int32 ndarray, cer float64 ndarray)."""
    _require_cuda()
    n = len(labels)
    if n == 0:
        return np.zeros(0, np.int32), np.zeros(0, np.float64)
    device = device or torch.device("cuda", torch.cuda.current_device())
    lf, lo = _encode_csr(labels)
    pf, po = _encode_csr(list(preds[:n]))
    max_len = int(max(np.diff(lo).max(), np.diff(po).max()))
    nl, npd = max(len(lf), 1), max(len(pf), 1)
    host = torch.zeros(nl + npd + 2 * (n + 1), dtype=torch.int32, pin_memory=True)
    host[: len(lf)] = torch.from_numpy(lf)
    host[nl: nl + len(pf)] = torch.from_numpy(pf)
    host[nl + npd: nl + npd + n + 1] = torch.from_numpy(lo)
    host[nl + npd + n + 1:] = torch.from_numpy(po)
    d = host.to(device, non_blocking=True)
    dist, cer = cer_batch(d[:nl], d[nl + npd: nl + npd + n + 1], None, d[nl: nl + npd], d[nl + npd + n + 1:], None, n, max_len)
    return dist.cpu().numpy(), cer.cpu().numpy()


def compare_labels(preds, labels):
    """utils.py:95-110: (exact-match count, sum over pairs of distance / max(1, len(label))) with the sum
    accumulated in list order in float64, as the Python loop does."""
    if not isinstance(labels, (list, tuple)):
        labels = [labels]
        print(labels)
    dist, cer = levenshtein_strings(preds, labels)
    correct_count = int((dist == 0).sum())
    total_cer = 0
    for c in cer.tolist():  # sequential fp64 accumulation == the reference loop
        total_cer += c
    return correct_count, total_cer


def decode_and_cer(scores, targets, target_offsets, target_lengths, max_target_len, blank=0):
    """Phase C without host round trips (train_nn_area.py:290-304): greedy-decode `scores` and score each sample
    against its ground-truth class indices (the CTC targets already on the device).
    Returns (codes, lens, dist, cer) device tensors. Equivalent to pred_to_string + per-sample compare_labels
    because char<->index is a bijection on properties.char_set."""
    codes, lens = decode_batch(scores, blank)
    B, T = codes.shape
    row_off = torch.arange(0, B * T, T, dtype=torch.int32, device=codes.device)
    dist, cer = cer_batch(targets, target_offsets, target_lengths, codes, row_off, lens, B, max(max_target_len, T))
    return codes, lens, dist, cer


# ---------------------------------------------------------------------------------------------------------------
# crop + pad of text strips
class _CropPad(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image, boxes, oh, ow):
        img = image.contiguous()
        H, W = img.shape[-2], img.shape[-1]
        n = boxes.shape[0]
        out = torch.empty((n, 1, oh, ow), dtype=torch.float32, device=img.device)
        _lib.call("qeb_crop_pad_gather", img.data_ptr(), H, W, boxes.data_ptr(), n, oh, ow, out.data_ptr(), _lib.stream())
        ctx.save_for_backward(boxes)
        ctx.shape = (tuple(image.shape), H, W, n, oh, ow)
        return out

    @staticmethod
    def backward(ctx, gout):
        (boxes,) = ctx.saved_tensors
        shape, H, W, n, oh, ow = ctx.shape
        gimg = torch.zeros(shape, dtype=torch.float32, device=gout.device)
        g = gout.contiguous()
        _lib.call("qeb_crop_pad_scatter", g.data_ptr(), H, W, boxes.data_ptr(), n, oh, ow, gimg.data_ptr(), _lib.stream())
        return gimg, None, None, None


def get_text_stack(image, labels, input_size):
    """utils.py:128-141: image (1,H,W) CUDA fp32, labels = list of dicts with label/x_min/y_min/x_max/y_max.
    Returns (strips (n,1,h,w), list of label strings); differentiable w.r.t. image."""
    if not image.is_cuda or image.dtype != torch.float32 or image.shape[0] != 1:
        raise _lib.QebError("qeb get_text_stack needs a (1,H,W) CUDA fp32 image (no CPU fallback)")
    labels_out = [lbl["label"] for lbl in labels]
    if len(labels) == 0:
        raise RuntimeError("stack expects a non-empty TensorList")  # same failure as torch.stack([])
    host = torch.tensor([[lbl["x_min"], lbl["y_min"], lbl["x_max"], lbl["y_max"]] for lbl in labels], dtype=torch.int32).pin_memory()
    boxes = host.to(image.device, non_blocking=True)
    return _CropPad.apply(image, boxes, int(input_size[0]), int(input_size[1])), labels_out
