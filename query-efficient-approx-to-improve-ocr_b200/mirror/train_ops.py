"""Training-step pieces around the two networks, with the reference's call shapes.

  MSEOnesLoss   MSELoss()(img_preds, torch.ones(img_preds.shape))  train_nn_patch.py:181-184, train_nn_area.py:177-181
  Adam          torch.optim.Adam(params, lr, weight_decay)         train_nn_patch.py:146-152, train_nn_area.py:149-154
Adam keeps torch's per-parameter state layout (step, exp_avg, exp_avg_sq), so optimizer state_dicts saved by the
reference (`optim_*_latest`, train_nn_patch.py:458-464) load, and the other way round; the update itself is ONE
launch over all tensors (qeb_adam_multi, csrc/train_ops.cu).
"""
import ctypes

import numpy as np
import torch

from .. import _lib


class _MSEOnes(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        if not x.is_cuda or x.dtype != torch.float32:
            raise _lib.QebError("qeb MSE needs a CUDA fp32 tensor (no CPU fallback)")
        x = x.contiguous()
        loss = torch.empty((), dtype=torch.float32, device=x.device)
        _lib.call("qeb_mse_ones_fwd", x.data_ptr(), x.numel(), loss.data_ptr(), _lib.stream())
        ctx.save_for_backward(x)
        return loss

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        dx = torch.empty_like(x)
        _lib.call("qeb_mse_ones_bwd", x.data_ptr(), x.numel(), g.contiguous().data_ptr(), dx.data_ptr(), _lib.stream())
        return dx


def mse_to_ones(x):
    """mean((x - 1)^2): the secondary loss of _get_loss without materialising the ones tensor."""
    return _MSEOnes.apply(x)


class MSELoss(torch.nn.Module):
    """torch.nn.MSELoss() as the reference uses it; a target that is all ones takes the fused path."""

    def forward(self, input, target):
        if target.numel() == input.numel() and bool((target == 1).all()):
            return mse_to_ones(input)
        raise _lib.QebError("qeb MSELoss implements the reference's MSE-to-white (all-ones target) only")


class Adam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self._tables = {}

    def _table(self, gi, group, ps):
        # the device table embeds raw pointers of the parameters, gradients AND moment buffers: all of them are part of the
        # key, so optimizer.load_state_dict() (which replaces exp_avg / exp_avg_sq) rebuilds the table
        key = (gi, tuple(p.data_ptr() for p in ps), tuple(p.grad.data_ptr() for p in ps),
               tuple(self.state[p]["exp_avg"].data_ptr() for p in ps), tuple(self.state[p]["exp_avg_sq"].data_ptr() for p in ps))
        cached = self._tables.get(gi)
        if cached is not None and cached[0] == key:
            return cached[1], cached[2]
        esz = _lib.load().qeb_adam_table_entry_bytes()
        assert esz == 48
        host = np.zeros((len(ps), 6), dtype=np.int64)
        chunk = 0
        for i, p in enumerate(ps):
            st = self.state[p]
            host[i] = (p.data_ptr(), p.grad.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), p.numel(), chunk)
            chunk += (p.numel() + 1023) // 1024
        dev = torch.from_numpy(host).pin_memory().to(ps[0].device, non_blocking=True)
        self._tables[gi] = (key, dev, chunk)
        return dev, chunk

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._tables = {}   # the moment tensors were replaced

    def __setstate__(self, state):
        super().__setstate__(state)
        self._tables = {}

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            for p in ps:
                if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous() or not p.grad.is_contiguous():
                    raise _lib.QebError("qeb Adam needs contiguous CUDA fp32 parameters and gradients (no CPU fallback)")
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            steps = {float(self.state[p]["step"]) for p in ps}
            if len(steps) != 1:
                raise _lib.QebError("qeb Adam: parameters of one group must share the step count")
            step = int(steps.pop()) + 1
            table, n_chunks = self._table(gi, group, ps)
            b1, b2 = group["betas"]
            _lib.call("qeb_adam_multi", table.data_ptr(), len(ps), n_chunks, float(group["lr"]), float(b1), float(b2),
                      float(group["eps"]), float(group["weight_decay"]), step, _lib.stream())
            for p in ps:
                self.state[p]["step"] += 1
        return loss
