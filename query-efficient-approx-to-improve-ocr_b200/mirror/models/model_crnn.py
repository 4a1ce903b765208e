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
from ..ctc import log_softmax, log_softmax_node

_POISON = os.environ.get("QEB_POISON_WORKSPACE", "0") == "1"
# 1 (default): CRNN.forward takes log_softmax (+ the per-frame arg-max pred_to_string needs) from the Linear GEMM's epilogue;
# 0: logits from the GEMM, log-softmax as its own launch (same-box A/B, debugging)
_FUSED_HEAD = os.environ.get("QEB_FUSED_HEAD", "1") != "0"


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
    """x (B,1,32,W) + the 36 parameters -> logits (T,B,V); with `opts` the fused variants of the C ABI
    (qeb_crnn_forward_fused): opts["log_softmax"] -> the output already holds log_softmax(logits) and opts["path"] receives
    the per-frame arg-max (T,B) int32; opts["jitter"] = dict(sigma, mean, coef, seed, seed_dev, return_noise) -> x is the
    CLEAN batch, the Gaussian jitter rides conv1's input load, opts["noisy"] (and opts["noise"]) receive the jittered batch."""

    @staticmethod
    def forward(ctx, x, bn_train, buffers, opts, *params):
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
        x_saved = x
        if not opts:
            _lib.call("qeb_crnn_forward", x.data_ptr(), B, W, V, _ptr_array(params), _ptr_array(buffers), int(bn_train),
                      ws.data_ptr(), logits.data_ptr(), _lib.stream())
        else:
            lsm = bool(opts.get("log_softmax"))
            path = torch.empty((T, B), dtype=torch.int32, device=x.device) if lsm else None
            jit = opts.get("jitter")
            sigma = noisy = noise = seed_dev = None
            mean, coef, seed = 0.0, 1.0, 0
            if jit is not None:
                sigma = jit["sigma"]
                if sigma.numel() != B or not sigma.is_cuda or sigma.dtype != torch.float32 or not sigma.is_contiguous():
                    raise _lib.QebError("qeb CRNN jitter: sigma must be a contiguous CUDA fp32 tensor with one value per image")
                mean, coef, seed, seed_dev = float(jit.get("mean", 0.0)), float(jit.get("coef", 1.0)), int(jit.get("seed", 0)), jit.get("seed_dev")
                noisy = jit.get("out")
                if noisy is None:
                    noisy = torch.empty_like(x)
                elif noisy.shape != x.shape or not noisy.is_contiguous() or noisy.device != x.device or noisy.dtype != torch.float32:
                    raise _lib.QebError("qeb CRNN jitter: `out` must be a contiguous fp32 tensor of the input's shape on its device")
                noise = torch.empty_like(x) if jit.get("return_noise") else None
                x_saved = noisy   # conv1's weight gradient reads the image the network saw
            _lib.call("qeb_crnn_forward_fused", x.data_ptr(), B, W, V, _ptr_array(params), _ptr_array(buffers), int(bn_train),
                      ws.data_ptr(), logits.data_ptr(), int(lsm), _lib.ptr(path), _lib.ptr(sigma), mean, coef, seed, _lib.ptr(seed_dev),
                      _lib.ptr(noisy), _lib.ptr(noise), _lib.stream())
            opts["path"], opts["noisy"], opts["noise"] = path, noisy, noise
        ctx.save_for_backward(x_saved, *params)
        ctx.ws = ws
        ctx.cfg = (B, W, V, int(bn_train))
        ctx.jittered = bool(opts) and opts.get("jitter") is not None
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        x, *params = ctx.saved_tensors
        B, W, V, bn_train = ctx.cfg
        dlogits = dlogits.contiguous()
        grads = _alloc_grads(params, ctx.needs_input_grad[4:])
        # a jittered input is not differentiated (the reference adds the noise to detached CPU copies, train_nn_area.py:227,260)
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] and not ctx.jittered else None
        _lib.call("qeb_crnn_backward", x.data_ptr(), B, W, V, _ptr_array(params), bn_train, ctx.ws.data_ptr(),
                  dlogits.data_ptr(), _ptr_array(grads), _lib.ptr(dx), _lib.stream())
        ctx.ws = None
        return (dx, None, None, None) + tuple(grads)


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

    def _trunk(self, x, opts):
        if not x.is_cuda or x.dtype != torch.float32:
            raise _lib.QebError("qeb CRNN needs a CUDA fp32 input (no CPU fallback)")
        c = self._stack()
        if c.batchnorm1.training != c.batchnorm2.training:
            raise _lib.QebError("qeb CRNN: batchnorm1 and batchnorm2 must be in the same mode")
        params = self.qeb_parameters()
        for p in params:
            if not p.is_contiguous() or p.device != x.device:
                raise _lib.QebError("qeb CRNN: parameters must be contiguous and on the input's device")
        return _CRNNTrunk.apply(x.contiguous(), c.batchnorm1.training, self.qeb_buffers(), opts, *params)

    def forward_logits(self, x):
        return self._trunk(x, None)

    def forward(self, x):
        if not _FUSED_HEAD:
            return log_softmax(self.forward_logits(x))
        opts = {"log_softmax": True}
        out = log_softmax_node(self._trunk(x, opts))     # its own autograd node: backward_hook sees the gradient at the logits
        out._qeb_path = (opts["path"], out._version)     # pred_to_string collapses this instead of re-reading (T,B,V) scores
        return out

    def forward_jittered(self, images, sigmas, mean=0.0, noise_coef=1, seed=None, seed_dev=None, return_noise=False, out=None):
        """`self(AddGaussianNoice(...)(images))` for a whole batch in ONE pass: the jitter of transform_helper.py:33-45 (one
        sigma per image, `sigmas`: (B) host or device) rides conv1's input load. Returns (log_probs, noisy_images[, noise]);
        noisy_images is what the OCR engine gets (train_nn_area.py:260-261) and is bit-identical to
        transform_helper.jitter_batch(images, sigmas, mean, noise_coef, seed) under the same seed. `seed_dev`: device-resident
        Philox key (CUDA-graph replays draw fresh noise), `out`: preallocated noisy batch. No gradient flows to `images`."""
        sg = torch.as_tensor(sigmas, dtype=torch.float32)
        if not sg.is_cuda:
            sg = sg.pin_memory().to(images.device, non_blocking=True)
        if seed is None:
            seed = 0 if seed_dev is not None else int(torch.randint(0, 2 ** 62, (1,)).item())
        opts = {"log_softmax": True, "jitter": {"sigma": sg.contiguous(), "mean": mean, "coef": noise_coef, "seed": seed,
                                                "seed_dev": seed_dev, "return_noise": return_noise, "out": out}}
        lp = log_softmax_node(self._trunk(images, opts))
        lp._qeb_path = (opts["path"], lp._version)
        return (lp, opts["noisy"], opts["noise"]) if return_noise else (lp, opts["noisy"])

    def map_to_sequence(self, map):
        batch, channel, height, width = map.size()
        sequence = map.permute(3, 0, 1, 2)
        sequence = sequence.contiguous().view(width, batch, -1)
        return sequence

    def backward_hook(self, module, grad_input, grad_output):
        for g in grad_input:
            g[g != g] = 0
