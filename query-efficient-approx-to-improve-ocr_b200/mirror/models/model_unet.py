"""Drop-in for the reference's models/model_unet.py (UNet :7-109).

Same class name, constructor signature, sub-module names and state_dict keys (encoder1.enc1conv1.weight,
encoder1.enc1norm1.*, ..., upconv4.weight/bias, conv.weight/bias). The sub-modules only hold the parameters and
BatchNorm buffers; `UNet.forward` runs the whole network in libqeb_sm100.so (qeb_unet_forward / qeb_unet_backward,
csrc/engine_unet.cu).
"""
from collections import OrderedDict

import ctypes
import os

import torch
import torch.nn as nn

from ... import _lib
from .model_crnn import _alloc_grads, _ptr_array


_POISON = os.environ.get("QEB_POISON_WORKSPACE", "0") == "1"


class _UNetFn(torch.autograd.Function):
    """x (B,1,H,W) + the 64 parameters -> sigmoid(conv(dec1)) (B,1,H,W)."""

    @staticmethod
    def forward(ctx, x, bn_train, buffers, tail_event, *params):
        lib = _lib.load()
        B, C, H, W = x.shape
        ctx.tail_event = tail_event
        nbytes = lib.qeb_unet_workspace_bytes(B, H, W)
        if C != 1 or nbytes == 0:
            raise _lib.QebError(f"qeb UNet: unsupported input {tuple(x.shape)} (needs (B,1,H,W) with H, W multiples of 16)")
        ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
        if _POISON:  # debugging aid: any read of workspace memory the pass did not write shows up as NaN
            ws.view(torch.float32).fill_(float("nan"))
        y = torch.empty_like(x)
        _lib.call("qeb_unet_forward", x.data_ptr(), B, H, W, _ptr_array(params), _ptr_array(buffers), int(bn_train),
                  ws.data_ptr(), y.data_ptr(), _lib.stream())
        ctx.save_for_backward(x, y, *params)
        ctx.ws = ws
        ctx.cfg = (B, H, W, int(bn_train))
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y, *params = ctx.saved_tensors
        B, H, W, bn_train = ctx.cfg
        need = list(ctx.needs_input_grad[4:])
        need[-1] = need[-2] = True  # the final conv's gradients are always produced by the fused sigmoid backward
        grads = _alloc_grads(params, need)
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        if ctx.tail_event is not None:
            # data-parallel overlap (mirror/dist.BucketedAllReduce): two events, fired when parameters [30, 64) / [18, 30) are final
            evs = (ctypes.c_void_p * 2)(ctx.tail_event[0].cuda_event, ctx.tail_event[1].cuda_event)
            _lib.call("qeb_unet_backward_bucketed", x.data_ptr(), B, H, W, _ptr_array(params), bn_train, ctx.ws.data_ptr(),
                      y.data_ptr(), dy.contiguous().data_ptr(), _ptr_array(grads), _lib.ptr(dx), evs, _lib.stream())
        else:
            _lib.call("qeb_unet_backward", x.data_ptr(), B, H, W, _ptr_array(params), bn_train, ctx.ws.data_ptr(),
                      y.data_ptr(), dy.contiguous().data_ptr(), _ptr_array(grads), _lib.ptr(dx), _lib.stream())
        ctx.ws = None
        out = [g if n else None for g, n in zip(grads, ctx.needs_input_grad[4:])]
        return (dx, None, None, None) + tuple(out)


class UNet(nn.Module):

    def __init__(self, in_channels=1, out_channels=1, init_features=32):
        super(UNet, self).__init__()
        if in_channels != 1 or out_channels != 1 or init_features != 32:
            raise _lib.QebError("qeb UNet implements the reference configuration UNet(1, 1, 32) only")
        features = init_features
        self.encoder1 = UNet._block(in_channels, features, name="enc1")
        self.pool1 = nn.MaxPool2d(kernel_size=2, stride=2)
        self.encoder2 = UNet._block(features, features * 2, name="enc2")
        self.pool2 = nn.MaxPool2d(kernel_size=2, stride=2)
        self.encoder3 = UNet._block(features * 2, features * 4, name="enc3")
        self.pool3 = nn.MaxPool2d(kernel_size=2, stride=2)
        self.encoder4 = UNet._block(features * 4, features * 8, name="enc4")
        self.pool4 = nn.MaxPool2d(kernel_size=2, stride=2)
        self.bottleneck = UNet._block(features * 8, features * 16, name="bottleneck")
        self.upconv4 = nn.ConvTranspose2d(features * 16, features * 8, kernel_size=2, stride=2)
        self.decoder4 = UNet._block((features * 8) * 2, features * 8, name="dec4")
        self.upconv3 = nn.ConvTranspose2d(features * 8, features * 4, kernel_size=2, stride=2)
        self.decoder3 = UNet._block((features * 4) * 2, features * 4, name="dec3")
        self.upconv2 = nn.ConvTranspose2d(features * 4, features * 2, kernel_size=2, stride=2)
        self.decoder2 = UNet._block((features * 2) * 2, features * 2, name="dec2")
        self.upconv1 = nn.ConvTranspose2d(features * 2, features, kernel_size=2, stride=2)
        self.decoder1 = UNet._block(features * 2, features, name="dec1")
        self.conv = nn.Conv2d(in_channels=features, out_channels=out_channels, kernel_size=1)

    def _blocks(self):
        return [(self.encoder1, "enc1"), (self.encoder2, "enc2"), (self.encoder3, "enc3"), (self.encoder4, "enc4"),
                (self.bottleneck, "bottleneck"), (self.decoder4, "dec4"), (self.decoder3, "dec3"), (self.decoder2, "dec2"),
                (self.decoder1, "dec1")]

    def qeb_parameters(self):
        """The 64 parameters in the order of the C ABI (csrc/engine_unet.cu)."""
        ps = []
        for blk, n in self._blocks():
            for k in ("1", "2"):
                norm = getattr(blk, f"{n}norm{k}")
                ps += [getattr(blk, f"{n}conv{k}").weight, norm.weight, norm.bias]
        for up in (self.upconv4, self.upconv3, self.upconv2, self.upconv1):
            ps += [up.weight, up.bias]
        ps += [self.conv.weight, self.conv.bias]
        return ps

    QEB_BUCKET_STARTS = (30, 18)   # ABI order: parameters [30, 64) are final first in the backward pass, then [18, 30), then [0, 18)

    def qeb_buffers(self):
        bs = []
        for blk, n in self._blocks():
            for k in ("1", "2"):
                norm = getattr(blk, f"{n}norm{k}")
                bs += [norm.running_mean, norm.running_var, norm.num_batches_tracked]
        return bs

    def forward(self, x):
        if not x.is_cuda or x.dtype != torch.float32:
            raise _lib.QebError("qeb UNet needs a CUDA fp32 input (no CPU fallback)")
        modes = {getattr(blk, f"{n}norm{k}").training for blk, n in self._blocks() for k in ("1", "2")}
        if len(modes) != 1:
            raise _lib.QebError("qeb UNet: all BatchNorm layers must be in the same mode")
        params = self.qeb_parameters()
        for p in params:
            if not p.is_contiguous() or p.device != x.device:
                raise _lib.QebError("qeb UNet: parameters must be contiguous and on the input's device")
        return _UNetFn.apply(x.contiguous(), modes.pop(), self.qeb_buffers(), getattr(self, "_qeb_tail_event", None), *params)

    @staticmethod
    def _block(in_channels, features, name):
        return nn.Sequential(
            OrderedDict([
                (name + "conv1", nn.Conv2d(in_channels=in_channels, out_channels=features, kernel_size=3, padding=1,
                                           bias=False)),
                (name + "norm1", nn.BatchNorm2d(num_features=features)),
                (name + "relu1", nn.ReLU(inplace=True)),
                (name + "conv2", nn.Conv2d(in_channels=features, out_channels=features, kernel_size=3, padding=1,
                                           bias=False)),
                (name + "norm2", nn.BatchNorm2d(num_features=features)),
                (name + "relu2", nn.ReLU(inplace=True)),
            ]))
