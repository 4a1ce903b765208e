"""CPU study for round 2: what does a 16-bit gradient operand cost in gradient accuracy?

The backward contractions (dgrad, wgrad) read fp32 operands as tf32 today (the tensor core truncates to a 10-bit mantissa).
Candidates with half the operand bytes: bf16 (7-bit mantissa), fp16 with one power-of-two scale per gradient tensor
(10-bit mantissa, range handled by the scale). This script evaluates the reference graphs (oracle/nn_oracle.py, torch CPU)
with every Conv2d's BACKWARD computed from operands rounded in one of these ways, and reports the error of the parameter
gradients against exact fp32 - cosine and relative norm error per network - plus the fraction of gradient elements that a
scaled fp16 copy flushes to zero. Forward operands are fp16-rounded in all modes (as in the product).

    python scripts/exp/grad_operand_study.py [batch]          (SLACK=k: the scale comes from a bound 2^k above the true maximum)
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import torch.nn.functional as F
from torch.nn.grad import conv2d_input, conv2d_weight

import bench
from oracle import nn_oracle
from qeb_b200.mirror.models.model_crnn import CRNN
from qeb_b200.mirror.models.model_unet import UNet

MODE = ["exact"]
SLACK = int(os.environ.get("SLACK", "0"))
STATS = {"n": 0, "flushed": 0}


def trunc_tf32(t):   # what the tensor core does with an fp32 operand: drop the low 13 mantissa bits
    return (t.contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32)


def fp16_scaled(g):  # one power-of-two scale per tensor so that max|g| lands in [2^14, 2^15)
    m = float(g.abs().max())
    if m == 0.0 or m != m:
        return g
    s = 2.0 ** (14 - SLACK - int(torch.floor(torch.log2(torch.tensor(m)))))   # SLACK: binades wasted by a conservative bound on max|g|
    q = (g * s).half().float() / s
    STATS["n"] += g.numel()
    STATS["flushed"] += int(((q == 0) & (g != 0)).sum())
    return q


class QConv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, stride, padding):
        ctx.save_for_backward(x, w)
        ctx.cfg = (stride, padding, b is not None)
        return F.conv2d(x.half().float(), w.half().float(), b, stride, padding)

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        stride, padding, has_b = ctx.cfg
        mode = MODE[0]
        if mode == "exact":
            xq, wq, gq = x, w, gy
        elif mode == "tf32":
            xq, wq, gq = trunc_tf32(x), trunc_tf32(w), trunc_tf32(gy)
        elif mode == "bf16":
            xq, wq, gq = x.bfloat16().float(), w.bfloat16().float(), gy.bfloat16().float()
        elif mode == "fp16s":   # activations / weights: the forward's fp16 copies; gradient: scaled fp16
            xq, wq, gq = x.half().float(), w.half().float(), fp16_scaled(gy)
        gx = conv2d_input(x.shape, wq, gq, stride, padding) if ctx.needs_input_grad[0] else None
        gw = conv2d_weight(xq, w.shape, gq, stride, padding)
        gb = gy.sum((0, 2, 3)) if has_b else None
        return gx, gw, gb, None, None


def patched_forward(self, x):
    return QConv.apply(x, self.weight, self.bias, self.stride, self.padding)


def grads(fn, params, mode):
    MODE[0] = mode
    for p in params:
        p.grad = None
    fn().backward()
    return torch.cat([p.grad.reshape(-1) for p in params if p.grad is not None]).double()


def report(name, fn, params):
    ref = grads(fn, params, "exact")
    print(f"{name}: |g| = {float(ref.norm()):.3e}")
    for mode in ("tf32", "fp16s", "bf16"):
        STATS["n"] = STATS["flushed"] = 0
        g = grads(fn, params, mode)
        cos = float((g @ ref) / (g.norm() * ref.norm()))
        rel = float((g - ref).norm() / ref.norm())
        extra = f"  flushed to zero: {STATS['flushed'] / max(1, STATS['n']):.2e} of the gradient elements" if mode == "fp16s" else ""
        print(f"  {mode:6s} relative error {rel:.3e}   1 - cosine {1 - cos:.3e}{extra}")


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    torch.manual_seed(42)
    torch.set_num_threads(os.cpu_count() or 1)
    orig = torch.nn.Conv2d.forward
    torch.nn.Conv2d.forward = patched_forward
    try:
        x, labels = bench.synth_batch(B, 7)
        unet, crnn = UNet(), CRNN(95, False)
        unet.train(); crnn.train()
        c2i = {c: i for i, c in enumerate(bench.CHAR_SET)}
        y, ys = bench.encode(labels, c2i)
        il = torch.full((B,), 31, dtype=torch.int32)

        def step():   # phase-B loss: CTC through the surrogate + MSE-to-white
            img = nn_oracle.unet_forward(unet, x)
            lp = nn_oracle.crnn_forward(crnn, img)
            return F.ctc_loss(lp, y, il, ys) + F.mse_loss(img, torch.ones_like(img))

        report(f"UNet conv weights (phase-B step, B = {B})", step, [p for n, p in unet.named_parameters() if "conv" in n and p.dim() == 4 and "upconv" not in n])
        report(f"CRNN conv weights (same step)", step, [p for n, p in crnn.named_parameters() if "conv" in n and p.dim() == 4])
        x.requires_grad_(True)   # the end of the input-gradient chain: every dgrad of both networks in sequence
        report("gradient at the UNet input (all input-gradient contractions chained)", step, [x])
    finally:
        torch.nn.Conv2d.forward = orig


if __name__ == "__main__":
    main()
