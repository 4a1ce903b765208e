"""Drop-in for torch.nn.CTCLoss as the reference constructs and calls it.

Reference: ctor train_nn_patch.py:143-144, train_nn_area.py:146-148, train_crnn.py:130-131
           (CTCLoss() and CTCLoss(reduction='none'), blank=0, zero_infinity=False);
           calls train_nn_patch.py:178,294, train_nn_area.py:174,265, train_crnn.py:160,
           tracking_utils.py:68,72 (on scores[:, img_indices, :] subsets).
Call signature is unchanged: loss(log_probs[T,B,V], targets int32 (1-D concatenated or 2-D padded, CPU or CUDA),
input_lengths[B], target_lengths[B]). The work runs in qeb_ctc_fwd / qeb_ctc_bwd (csrc/ctc.cu).
"""
import numpy as np
import torch

from .. import _lib

_RED = {"none": 0, "mean": 1, "sum": 2}


def _pack_int_args(targets, input_lengths, target_lengths, B, device):
    """One pinned staging buffer + one H2D copy for [targets | offsets | input_lengths | target_lengths]."""
    tl = torch.as_tensor(target_lengths).to("cpu", torch.int32).reshape(-1)
    il = torch.as_tensor(input_lengths).to("cpu", torch.int32).reshape(-1)
    tg = torch.as_tensor(targets).to("cpu", torch.int32)
    if tl.numel() != B or il.numel() != B:
        raise RuntimeError(f"CTCLoss: expected {B} input/target lengths, got {il.numel()}/{tl.numel()}")
    tl_np = tl.numpy()
    if tg.dim() == 2:  # padded (B, S) form: compact it
        rows = [tg[b, : tl_np[b]] for b in range(B)]
        tg = torch.cat(rows) if rows else tg.reshape(-1)
    tg = tg.reshape(-1)
    n_t = int(tl_np.sum()) if B else 0
    if tg.numel() < n_t:
        raise RuntimeError("CTCLoss: targets shorter than sum(target_lengths)")
    max_len = int(tl_np.max()) if B else 0
    offs = np.zeros(B, dtype=np.int32)
    if B > 1:
        np.cumsum(tl_np[:-1], out=offs[1:])
    n_pad = max(n_t, 1)
    host = torch.empty(n_pad + 3 * B, dtype=torch.int32, pin_memory=True)
    host[:n_t] = tg[:n_t]
    if n_t == 0:
        host[0] = 0
    host[n_pad:n_pad + B] = torch.from_numpy(offs)
    host[n_pad + B:n_pad + 2 * B] = il
    host[n_pad + 2 * B:] = tl
    dev = host.to(device, non_blocking=True)
    return dev[:n_pad], dev[n_pad:n_pad + B], dev[n_pad + B:n_pad + 2 * B], dev[n_pad + 2 * B:], max_len


class PackedTargets:
    """Targets / lengths already staged on the device (pack_targets): lets a caller that keeps its labels on the GPU
    skip the per-call host staging + H2D copy that the reference's int32-CPU-tensor calling convention implies."""

    def __init__(self, tg, offs, il, tl, max_len, B):
        self.tg, self.offs, self.il, self.tl, self.max_len, self.B = tg, offs, il, tl, max_len, B


def pack_targets(targets, input_lengths, target_lengths, device):
    B = int(torch.as_tensor(target_lengths).numel())
    return PackedTargets(*_pack_int_args(targets, input_lengths, target_lengths, B, torch.device(device)), B)


class _CTC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, log_probs, targets, input_lengths, target_lengths, blank, reduction, zero_infinity, batch_index):
        if not log_probs.is_cuda:
            raise _lib.QebError("qeb CTCLoss needs CUDA log_probs (no CPU fallback)")
        if log_probs.dtype != torch.float32:
            raise _lib.QebError("qeb CTCLoss computes in fp32; got " + str(log_probs.dtype))
        lp = log_probs if log_probs.stride(2) == 1 else log_probs.contiguous()
        T, Bfull, V = lp.shape
        dev = lp.device
        bidx = None
        B = Bfull
        if batch_index is not None:
            bidx = torch.as_tensor(batch_index, dtype=torch.int32).to(dev)
            B = bidx.numel()
        if isinstance(targets, PackedTargets):
            if targets.B != B:
                raise RuntimeError(f"CTCLoss: packed targets hold {targets.B} samples, log_probs {B}")
            tg, offs, il, tl, max_len = targets.tg, targets.offs, targets.il, targets.tl, targets.max_len
        else:
            tg, offs, il, tl, max_len = _pack_int_args(targets, input_lengths, target_lengths, B, dev)
        red = _RED[reduction]
        log_alpha = torch.empty(_lib.load().qeb_ctc_workspace_bytes(B, T, max_len) // 4, dtype=torch.float32, device=dev)
        nll = torch.empty(B, dtype=torch.float32, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev) if red else None
        _lib.call("qeb_ctc_fwd", lp.data_ptr(), lp.stride(0), lp.stride(1), _lib.ptr(bidx), tg.data_ptr(), offs.data_ptr(),
                  il.data_ptr(), tl.data_ptr(), B, T, V, blank, max_len, red, int(zero_infinity), log_alpha.data_ptr(),
                  nll.data_ptr(), _lib.ptr(loss), _lib.stream())
        ctx.save_for_backward(lp, tg, offs, il, tl, log_alpha, nll)
        ctx.bidx = bidx
        ctx.cfg = (B, T, V, blank, max_len, red, int(zero_infinity), Bfull)
        if red:
            return loss
        return torch.where(torch.isinf(nll), torch.zeros_like(nll), nll) if zero_infinity else nll.clone()

    @staticmethod
    def backward(ctx, grad_out):
        lp, tg, offs, il, tl, log_alpha, nll = ctx.saved_tensors
        B, T, V, blank, max_len, red, zinf, Bfull = ctx.cfg
        go = grad_out.contiguous().to(torch.float32)
        if ctx.bidx is not None:
            grad = torch.zeros((T, Bfull, V), dtype=torch.float32, device=lp.device)
        else:
            grad = torch.empty((T, Bfull, V), dtype=torch.float32, device=lp.device)
        _lib.call("qeb_ctc_bwd", lp.data_ptr(), lp.stride(0), lp.stride(1), _lib.ptr(ctx.bidx), tg.data_ptr(),
                  offs.data_ptr(), il.data_ptr(), tl.data_ptr(), B, T, V, blank, max_len, red, zinf, log_alpha.data_ptr(),
                  nll.data_ptr(), go.data_ptr(), grad.data_ptr(), grad.stride(0), grad.stride(1), _lib.stream())
        return grad, None, None, None, None, None, None, None


def ctc_loss(log_probs, targets, input_lengths, target_lengths, blank=0, reduction="mean", zero_infinity=False,
             batch_index=None):
    """Functional form. `batch_index` selects columns of log_probs without materialising scores[:, idx, :]
    (the gather weighted_ctc_loss does at tracking_utils.py:65)."""
    return _CTC.apply(log_probs, targets, input_lengths, target_lengths, blank, reduction, zero_infinity, batch_index)


class CTCLoss(torch.nn.Module):
    """Same constructor and call signature as torch.nn.CTCLoss."""

    def __init__(self, blank=0, reduction="mean", zero_infinity=False):
        super().__init__()
        if reduction not in _RED:
            raise ValueError(f"{reduction} is not a valid value for reduction")
        self.blank = blank
        self.reduction = reduction
        self.zero_infinity = zero_infinity

    def forward(self, log_probs, targets, input_lengths=None, target_lengths=None):
        return ctc_loss(log_probs, targets, input_lengths, target_lengths, self.blank, self.reduction, self.zero_infinity)


class _LogSoftmax(torch.autograd.Function):
    """fn.log_softmax(x, -1) of models/model_crnn.py:20 (the node CRNN.backward_hook sees)."""

    @staticmethod
    def forward(ctx, x):
        x = x.contiguous()
        y = torch.empty_like(x)
        V = x.shape[-1]
        _lib.call("qeb_log_softmax_fwd", x.data_ptr(), y.data_ptr(), x.numel() // V, V, _lib.stream())
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        dy = dy.contiguous()
        dx = torch.empty_like(y)
        V = y.shape[-1]
        _lib.call("qeb_log_softmax_bwd", y.data_ptr(), dy.data_ptr(), dx.data_ptr(), y.numel() // V, V, _lib.stream())
        return dx


class _LogSoftmaxDone(torch.autograd.Function):
    """The log-softmax NODE for a tensor whose values the CRNN head's GEMM epilogue already turned into log-probs
    (qeb_crnn_forward_fused, log_softmax = 1): forward is the identity, backward is log-softmax's (dx = dy - exp(y) sum dy),
    so the autograd graph - and what CRNN.backward_hook sees (models/model_crnn.py:30-32) - is the unfused one."""

    @staticmethod
    def forward(ctx, y):
        ctx.save_for_backward(y)
        return y.view_as(y)

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        dy = dy.contiguous()
        dx = torch.empty_like(y)
        V = y.shape[-1]
        _lib.call("qeb_log_softmax_bwd", y.data_ptr(), dy.data_ptr(), dx.data_ptr(), y.numel() // V, V, _lib.stream())
        return dx


def log_softmax_node(log_probs):
    return _LogSoftmaxDone.apply(log_probs)


def log_softmax(x):
    if not x.is_cuda or x.dtype != torch.float32:
        raise _lib.QebError("qeb log_softmax needs a CUDA fp32 tensor (no CPU fallback)")
    return _LogSoftmax.apply(x)
