"""TEST INFRASTRUCTURE ONLY - Python face of the CPU oracle (oracle.c through ctypes, numpy, torch-CPU).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may import this
module; the product package never does. Every function cites the reference lines it restates.
See oracle.c for the parity-pinning statement (Levenshtein: "parity unpinned").
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    """Compile oracle.c -> liboracle.so (gcc). Building the checker is not using it."""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "liboracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        i32p = ctypes.POINTER(ctypes.c_int32)
        i64p = ctypes.POINTER(ctypes.c_int64)
        f32p = ctypes.POINTER(ctypes.c_float)
        f64p = ctypes.POINTER(ctypes.c_double)
        L.oracle_levenshtein.argtypes = [i32p, ctypes.c_int, i32p, ctypes.c_int]
        L.oracle_levenshtein.restype = ctypes.c_int
        L.oracle_compare_labels.argtypes = [i32p, i32p, i32p, i32p, ctypes.c_int, i32p, f64p, f64p]
        L.oracle_compare_labels.restype = ctypes.c_int
        L.oracle_greedy_decode.argtypes = [f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int, i32p, i32p]
        L.oracle_topk_stable.argtypes = [f32p, ctypes.c_int, ctypes.c_int, i64p]
        L.oracle_range_select.argtypes = [f32p, ctypes.c_int, f32p, ctypes.c_int, i64p, f32p]
        L.oracle_crop_pad.argtypes = [f32p, ctypes.c_int, ctypes.c_int, i32p, ctypes.c_int, ctypes.c_int, ctypes.c_int, f32p]
        L.oracle_jitter_apply.argtypes = [f32p, f32p, ctypes.c_float, ctypes.c_size_t, f32p]
        L.oracle_philox4x32_10.argtypes = [ctypes.c_uint32] * 6 + [ctypes.POINTER(ctypes.c_uint32)]
        _LIB = L
    return _LIB


def _p(a, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct))


def encode_csr(strings):
    """list[str] -> (int32 code points, int32 offsets[n+1])"""
    offs = np.zeros(len(strings) + 1, dtype=np.int32)
    np.cumsum(np.fromiter(map(len, strings), dtype=np.int64, count=len(strings)), out=offs[1:])
    # == np.fromiter((ord(c) for s in strings for c in s), int32), as one C-level UTF-32 encode (1 M strings in 0.1 s)
    flat = np.frombuffer(bytearray("".join(strings).encode("utf-32-le", "surrogatepass")), dtype="<i4")
    return flat, offs


# --- utils.py:103-109 / python-Levenshtein 0.12.0 `distance` -------------------------------------------------
def levenshtein(a, b):
    A = np.fromiter((ord(c) for c in a), dtype=np.int32, count=len(a))
    B = np.fromiter((ord(c) for c in b), dtype=np.int32, count=len(b))
    return int(lib().oracle_levenshtein(_p(A, ctypes.c_int32), len(a), _p(B, ctypes.c_int32), len(b)))


def levenshtein_py(a, b):
    """Pure-Python two-row restatement (small cases; cross-checks the C one)."""
    prev = list(range(len(b) + 1))
    for i, ca in enumerate(a, 1):
        cur = [i]
        for j, cb in enumerate(b, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (ca != cb)))
        prev = cur
    return prev[-1]


def compare_labels(preds, labels, return_all=False):
    """utils.py:95-110. Returns (correct_count, total_cer) (+ per-pair distances and CERs)."""
    if not isinstance(labels, (list, tuple)):
        labels = [labels]
    n = len(labels)
    pf, po = encode_csr(list(preds[:n]))
    lf, lo = encode_csr(list(labels))
    dist = np.zeros(n, dtype=np.int32)
    cer = np.zeros(n, dtype=np.float64)
    tot = ctypes.c_double(0.0)
    if len(pf) == 0:
        pf = np.zeros(1, dtype=np.int32)
    if len(lf) == 0:
        lf = np.zeros(1, dtype=np.int32)
    c = lib().oracle_compare_labels(_p(pf, ctypes.c_int32), _p(po, ctypes.c_int32), _p(lf, ctypes.c_int32),
                                    _p(lo, ctypes.c_int32), n, _p(dist, ctypes.c_int32), _p(cer, ctypes.c_double),
                                    ctypes.byref(tot))
    if return_all:
        return int(c), float(tot.value), dist, cer
    return int(c), float(tot.value)


# --- utils.py:74-92 -------------------------------------------------------------------------------------------
def greedy_decode(scores):
    """scores (T,B,V) float32 array -> (codes (B,T) int32 padded -1, lengths (B))"""
    s = np.ascontiguousarray(np.asarray(scores, dtype=np.float32))
    T, B, V = s.shape
    out = np.zeros((B, T), dtype=np.int32)
    ln = np.zeros(B, dtype=np.int32)
    lib().oracle_greedy_decode(_p(s, ctypes.c_float), T, B, V, _p(out, ctypes.c_int32), _p(ln, ctypes.c_int32))
    return out, ln


def pred_to_string(scores, index_to_char):
    codes, ln = greedy_decode(scores)
    return ["".join(index_to_char[int(c)] for c in codes[b, : ln[b]]) for b in range(codes.shape[0])]


# --- selection_utils.py:144-151 / :107-135 --------------------------------------------------------------------
def topk_query(cers_f32, k):
    v = np.ascontiguousarray(np.asarray(cers_f32, dtype=np.float32))
    k = min(int(k), len(v))
    out = np.zeros(max(k, 1), dtype=np.int64)
    lib().oracle_topk_stable(_p(v, ctypes.c_float), len(v), k, _p(out, ctypes.c_int64))
    return out[:k]


def range_query(cers_f32, rands_f32, return_points=False):
    v = np.ascontiguousarray(np.asarray(cers_f32, dtype=np.float32))
    r = np.ascontiguousarray(np.asarray(rands_f32, dtype=np.float32))
    out = np.zeros(max(len(r), 1), dtype=np.int64)
    pts = np.zeros(max(len(r), 1), dtype=np.float32)
    if len(v):
        lib().oracle_range_select(_p(v, ctypes.c_float), len(v), _p(r, ctypes.c_float), len(r), _p(out, ctypes.c_int64),
                                  _p(pts, ctypes.c_float))
    if return_points:
        return out[: len(r)], pts[: len(r)]
    return out[: len(r)]


# --- utils.py:118-141 -----------------------------------------------------------------------------------------
def crop_pad(img_hw, boxes, oh, ow):
    img = np.ascontiguousarray(np.asarray(img_hw, dtype=np.float32))
    bx = np.ascontiguousarray(np.asarray(boxes, dtype=np.int32)).reshape(-1, 4)
    out = np.zeros((len(bx), oh, ow), dtype=np.float32)
    if len(bx):
        lib().oracle_crop_pad(_p(img, ctypes.c_float), img.shape[0], img.shape[1], _p(bx, ctypes.c_int32), len(bx), oh, ow,
                              _p(out, ctypes.c_float))
    return out


def crop_pad_backward(gout, boxes, H, W):
    """Adjoint of crop_pad (autograd of slice + ConstantPad2d + stack): scatter-add in float64, numpy loops."""
    g = np.zeros((H, W), dtype=np.float64)
    n, oh, ow = gout.shape
    for b in range(n):
        x0, y0, x1, y1 = [int(v) for v in boxes[b]]
        x0 = min(max(x0, 0), W); y0 = min(max(y0, 0), H)
        x1 = min(max(x1, x0), W); y1 = min(max(y1, y0), H)
        cw, ch = x1 - x0, y1 - y0
        pl, pt = (ow - cw) // 2, (oh - ch) // 2
        for y in range(oh):
            cy = y - pt
            if cy < 0 or cy >= ch:
                continue
            xs = np.arange(ow)
            cx = xs - pl
            m = (cx >= 0) & (cx < cw)
            g[y0 + cy, x0 + cx[m]] += gout[b, y, xs[m]]
    return g


# --- transform_helper.py:33-45 --------------------------------------------------------------------------------
def jitter_apply(img, noise, coef=1.0):
    a = np.ascontiguousarray(np.asarray(img, dtype=np.float32))
    z = np.ascontiguousarray(np.asarray(noise, dtype=np.float32))
    out = np.empty_like(a)
    lib().oracle_jitter_apply(_p(a, ctypes.c_float), _p(z, ctypes.c_float), float(coef), a.size, _p(out, ctypes.c_float))
    return out


def philox4x32_10(counter, key):
    out = (ctypes.c_uint32 * 4)()
    lib().oracle_philox4x32_10(*[int(c) & 0xFFFFFFFF for c in counter], int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF, out)
    return [int(x) for x in out]


def philox_noise(n_img, hw, sigma, mean, seed):
    """The jitter kernel's in-kernel noise stream restated in numpy/float64 (Box-Muller on Philox4x32-10 with
    counter=(group, image_lo, image_hi, 0), key=(seed_lo, seed_hi)); compare with a float tolerance."""
    out = np.zeros((n_img, hw), dtype=np.float64)
    for im in range(n_img):
        for g in range(hw // 4):
            r = philox4x32_10((g, im & 0xFFFFFFFF, im >> 32, 0), (seed & 0xFFFFFFFF, seed >> 32))
            for h in range(2):
                u1 = ((r[2 * h] >> 8) + 1.0) / 16777216.0
                u2 = (r[2 * h + 1] >> 8) / 16777216.0
                rad = np.sqrt(-2.0 * np.log(np.float32(u1).astype(np.float64)))
                ang = 2.0 * np.pi * u2
                out[im, 4 * g + 2 * h] = mean + sigma[im] * rad * np.cos(ang)
                out[im, 4 * g + 2 * h + 1] = mean + sigma[im] * rad * np.sin(ang)
    return out


# --- torch.nn.CTCLoss call sites (train_nn_patch.py:143-144,178; train_nn_area.py:146-148,174) ---------------
def ctc_loss(log_probs, targets, input_lengths, target_lengths, reduction="mean", zero_infinity=False, blank=0):
    """ATen's native CPU CTC = what the reference's loss object runs on CPU. Returns (loss, grad wrt log_probs)."""
    import torch

    lp = torch.as_tensor(np.asarray(log_probs), dtype=torch.float32).clone().requires_grad_(True)
    loss = torch.nn.functional.ctc_loss(lp, torch.as_tensor(np.asarray(targets), dtype=torch.int32),
                                        torch.as_tensor(np.asarray(input_lengths), dtype=torch.int32),
                                        torch.as_tensor(np.asarray(target_lengths), dtype=torch.int32),
                                        blank=blank, reduction=reduction, zero_infinity=zero_infinity)
    (loss.sum() if reduction == "none" else loss).backward()
    return loss.detach().numpy(), lp.grad.numpy()
