"""Drop-in for the reference's models/model_crnn.py (CRNN :5-32, Convolutional :34-56).

Same class names, constructor signatures, sub-module / parameter names and state_dict keys, so pickles, optimizers,
`.train()/.eval()/.apply(set_bn_eval)`, `register_backward_hook(model.backward_hook)` and `zero_grad()` keep working.
The sub-modules only hold the parameters; `CRNN.forward` runs the whole network in libqeb_sm100.so
(qeb_crnn_forward / qeb_crnn_backward, csrc/engine_crnn.cu) followed by the log-softmax op (csrc/ctc.cu), which is
its own autograd node so that the reference's module backward hook sees the gradient at the logits exactly as it does
on the reference (models/model_crnn.py:30-32, train_nn_patch.py:94).
"""
import ctypes

import os

import torch
import torch.nn as nn

from ... import _lib
from ..ctc import log_softmax

_POISON = os.environ.get("QEB_POISON_WORKSPACE", "0") == "1"


class Convolutional(nn.Module):
    """Parameter container of the conv stack (models/model_crnn.py:34-45)."""

    def __init__(self):
        super(Convolutional, self).__init__()
        self.conv1 = nn.Conv2d(1, 64, kernel_size=3, stride=1, padding=1)
        self.conv2 = nn.Conv2d(64, 128, kernel_size=3, stride=1, padding=1)
        self.conv3 = nn.Conv2d(128, 256, kernel_size=3, stride=1, padding=1)
        self.conv4 = nn.Conv2d(256, 256, kernel_size=3, stride=1, padding=1)
        self.conv5 = nn.Conv2d(256, 512, kernel_size=3, stride=1, padding=1)
        self.batchnorm1 = nn.BatchNorm2d(num_features=512)
        self.conv6 = nn.Conv2d(512, 512, kernel_size=3, stride=1, padding=1)
        self.batchnorm2 = nn.BatchNorm2d(num_features=512)
        self.conv7 = nn.Conv2d(512, 512, kernel_size=2, stride=1, padding=0)

    def forward(self, x):
        raise _lib.QebError("the qeb Convolutional stack runs inside CRNN.forward (one fused launch sequence); "
                            "call the CRNN module")


def _ptr_array(tensors):
    arr = (ctypes.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = None if t is None else t.data_ptr()
    return arr


def _alloc_grads(params, need):
    """Zero-filled gradient tensors for the parameters that need one, carved out of ONE flat buffer (one memset; the
    kernels accumulate into them). AccumulateGrad adopts these views as .grad without a copy when .grad is None."""
    sizes = [(p.numel() + 63) // 64 * 64 if n else 0 for p, n in zip(params, need)]
    flat = torch.zeros(sum(sizes), dtype=torch.float32, device=params[0].device)
    grads, off = [], 0
    for p, n, s in zip(params, need, sizes):
        grads.append(flat[off:off + p.numel()].view_as(p) if n else None)
        off += s
    return grads


class _CRNNTrunk(torch.autograd.Function):
    """x (B,1,32,W) + the 36 parameters -> logits (T,B,V)."""

    @staticmethod
    def forward(ctx, x, bn_train, buffers, *params):
        lib = _lib.load()
        B, C, H, W = x.shape
        V = params[-1].shape[0]
        nbytes = lib.qeb_crnn_workspace_bytes(B, W, V)
        if C != 1 or H != 32 or nbytes == 0:
            raise _lib.QebError(f"qeb CRNN: unsupported input {tuple(x.shape)} (needs (B,1,32,W), W % 4 == 0, vocab <= 96)")
        ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
        if _POISON:  # debugging aid: any read of workspace memory the pass did not write shows up as NaN
            ws.view(torch.float32).fill_(float("nan"))
        T = W // 4 - 1
        logits = torch.empty((T, B, V), dtype=torch.float32, device=x.device)
        _lib.call("qeb_crnn_forward", x.data_ptr(), B, W, V, _ptr_array(params), _ptr_array(buffers), int(bn_train),
                  ws.data_ptr(), logits.data_ptr(), _lib.stream())
        ctx.save_for_backward(x, *params)
        ctx.ws = ws
        ctx.cfg = (B, W, V, int(bn_train))
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        x, *params = ctx.saved_tensors
        B, W, V, bn_train = ctx.cfg
        dlogits = dlogits.contiguous()
        grads = _alloc_grads(params, ctx.needs_input_grad[3:])
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        _lib.call("qeb_crnn_backward", x.data_ptr(), B, W, V, _ptr_array(params), bn_train, ctx.ws.data_ptr(),
                  dlogits.data_ptr(), _ptr_array(grads), _lib.ptr(dx), _lib.stream())
        ctx.ws = None
        return (dx, None, None) + tuple(grads)


class CRNN(nn.Module):

    def __init__(self, vocab_size, multi_gpu=True):
        super(CRNN, self).__init__()
        self.lstm = nn.LSTM(512, 256, 2, bidirectional=True)
        self.linear = nn.Linear(512, vocab_size)
        if multi_gpu:
            # kept for state_dict compatibility (keys convo.module.*); data parallelism is process-per-GPU in qeb
            self.convo = nn.DataParallel(Convolutional())
        else:
            self.convo = Convolutional()

    def _stack(self):
        return self.convo.module if isinstance(self.convo, nn.DataParallel) else self.convo

    def qeb_parameters(self):
        """The 36 parameters in the order of the C ABI (csrc/engine_crnn.cu)."""
        c = self._stack()
        ps = [c.conv1.weight, c.conv1.bias, c.conv2.weight, c.conv2.bias, c.conv3.weight, c.conv3.bias, c.conv4.weight,
              c.conv4.bias, c.conv5.weight, c.conv5.bias, c.batchnorm1.weight, c.batchnorm1.bias, c.conv6.weight,
              c.conv6.bias, c.batchnorm2.weight, c.batchnorm2.bias, c.conv7.weight, c.conv7.bias]
        for layer in range(2):
            for suffix in ("", "_reverse"):
                for name in ("weight_ih", "weight_hh", "bias_ih", "bias_hh"):
                    ps.append(getattr(self.lstm, f"{name}_l{layer}{suffix}"))
        ps += [self.linear.weight, self.linear.bias]
        return ps

    def qeb_buffers(self):
        c = self._stack()
        return [c.batchnorm1.running_mean, c.batchnorm1.running_var, c.batchnorm1.num_batches_tracked,
                c.batchnorm2.running_mean, c.batchnorm2.running_var, c.batchnorm2.num_batches_tracked]

    def forward_logits(self, x):
        if not x.is_cuda or x.dtype != torch.float32:
            raise _lib.QebError("qeb CRNN needs a CUDA fp32 input (no CPU fallback)")
        c = self._stack()
        if c.batchnorm1.training != c.batchnorm2.training:
            raise _lib.QebError("qeb CRNN: batchnorm1 and batchnorm2 must be in the same mode")
        params = self.qeb_parameters()
        for p in params:
            if not p.is_contiguous() or p.device != x.device:
                raise _lib.QebError("qeb CRNN: parameters must be contiguous and on the input's device")
        return _CRNNTrunk.apply(x.contiguous(), c.batchnorm1.training, self.qeb_buffers(), *params)

    def forward(self, x):
        return log_softmax(self.forward_logits(x))

    def map_to_sequence(self, map):
        batch, channel, height, width = map.size()
        sequence = map.permute(3, 0, 1, 2)
        sequence = sequence.contiguous().view(width, batch, -1)
        return sequence

    def backward_hook(self, module, grad_input, grad_output):
        for g in grad_input:
            g[g != g] = 0
