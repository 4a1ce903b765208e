"""CPU study: where does the network-gradient error against the fp32 oracle come from, and what is the floor?

Phase-B step (UNet train-mode BN -> CRNN, frozen BN -> CTC + MSE), per-tensor relative L2 of the UNet conv weight gradients
against the SAME graph in float64 ("truth"):
  fp32          the fp32 oracle itself (what any fp32 implementation differs by from the exact gradient)
  fwd16         forward conv operands rounded to 11-bit significands (fp16 RN), backward exact
  bwd_trunc     forward exact, backward conv operands truncated to tf32
  bwd_rn        forward exact, backward conv operands rounded to nearest tf32
  fwd16+bwd_rn  the product's arithmetic
    python scripts/exp/precision_floor.py [batch]
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import copy
import torch
import torch.nn.functional as F
from torch.nn.grad import conv2d_input, conv2d_weight

import bench
from oracle import nn_oracle
from qeb_b200.mirror.models.model_crnn import CRNN
from qeb_b200.mirror.models.model_unet import UNet
from qeb_b200.mirror.utils import set_bn_eval

MODE = {"fwd": "exact", "bwd": "exact"}


def trunc_tf32(t):
    return (t.contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32)


def rn_tf32(t):   # round to nearest, ties away (cvt.rna.tf32.f32)
    return ((t.contiguous().view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)


def q(t, mode):
    if mode == "exact" or t.dtype != torch.float32:
        return t
    return {"fp16": lambda v: v.half().float(), "trunc": trunc_tf32, "rn": rn_tf32}[mode](t)


class QConv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, stride, padding):
        ctx.save_for_backward(x, w)
        ctx.cfg = (stride, padding, b is not None)
        return F.conv2d(q(x, MODE["fwd"]), q(w, MODE["fwd"]), b, stride, padding)

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        stride, padding, has_b = ctx.cfg
        m = MODE["bwd"]
        gx = conv2d_input(x.shape, q(w, m), q(gy, m), stride, padding) if ctx.needs_input_grad[0] else None
        gw = conv2d_weight(q(x, m), w.shape, q(gy, m), stride, padding)
        return gx, gw, (gy.sum((0, 2, 3)) if has_b else None), None, None


def patched_forward(self, x):
    return QConv.apply(x, self.weight, self.bias, self.stride, self.padding)


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count() or 1)
    x, labels = bench.synth_batch(B, 11)
    unet, crnn = UNet(), CRNN(95, False)
    c2i = {c: i for i, c in enumerate(bench.CHAR_SET)}
    y, ys = bench.encode(labels, c2i)
    il = torch.full((B,), 31, dtype=torch.int32)

    def grads(un, cr, xx):
        un.train(); cr.train(); cr.apply(set_bn_eval)
        for p in un.parameters():
            p.grad = None
        img = nn_oracle.unet_forward(un, xx)
        lp = nn_oracle.crnn_forward(cr, img)
        loss = F.ctc_loss(lp, y, il, ys) + F.mse_loss(img, torch.ones_like(img))
        loss.backward()
        return {n: p.grad.detach().double().clone() for n, p in un.named_parameters()}

    truth = grads(copy.deepcopy(unet).double(), copy.deepcopy(crnn).double(), x.double())
    orig = torch.nn.Conv2d.forward
    torch.nn.Conv2d.forward = patched_forward
    try:
        for name, f, b in (("fp32", "exact", "exact"), ("fwd16", "fp16", "exact"), ("bwd_trunc", "exact", "trunc"), ("bwd_rn", "exact", "rn"),
                           ("fwd16+bwd_rn", "fp16", "rn"), ("fwd16+bwd_trunc", "fp16", "trunc")):
            MODE["fwd"], MODE["bwd"] = f, b
            g = grads(copy.deepcopy(unet), copy.deepcopy(crnn), x)
            errs = {n: float((g[n] - truth[n]).norm() / truth[n].norm()) for n in g}
            worst = max(errs, key=errs.get)
            sel = ["conv.weight", "decoder1.dec1conv2.weight", "encoder1.enc1conv1.weight", "encoder4.enc4conv1.weight", "bottleneck.bottleneckconv1.weight"]
            print(f"{name:16s} worst {errs[worst]:.2e} ({worst})  " + "  ".join(f"{s.split('.')[-2]}={errs[s]:.1e}" for s in sel))
    finally:
        torch.nn.Conv2d.forward = orig


if __name__ == "__main__":
    main()
