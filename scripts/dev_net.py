"""Dev check on the GPU box: LSTM kernels, CRNN and UNet engines against torch fp32 on the same device."""
import copy, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import qeb_b200
from qeb_b200 import _lib
from qeb_b200.mirror.models.model_crnn import CRNN
from qeb_b200.mirror.models.model_unet import UNet
from oracle import nn_oracle as torch_ref
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = "cuda"
which = sys.argv[1] if len(sys.argv) > 1 else "all"

def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))

def cmp_grads(m, mr, tag, thr=2e-2):
    if os.environ.get('SYNC'): torch.cuda.synchronize()
    worst = ("", 0.0)
    for (n, p), (_, q) in zip(m.named_parameters(), mr.named_parameters()):
        if q.grad is None:
            continue
        if p.grad is None:
            print("   MISSING grad", n); continue
        e = rel(p.grad, q.grad)
        if e > 1e3:
            big = ((p.grad.abs() > 1e3) | ~torch.isfinite(p.grad)).reshape(-1)
            idx = big.nonzero().reshape(-1)
            print("   DEBUG", n, tuple(p.shape), "count", int(big.sum()), "idx", idx[:10].tolist(), idx[-3:].tolist(), "vals",
                  p.grad.reshape(-1)[idx[:5]].tolist(), "ptr", hex(p.grad.data_ptr()), "base",
                  None if p.grad._base is None else (hex(p.grad._base.data_ptr()), p.grad._base.numel()))
        if q.grad.norm() < 1e-6:  # analytically-zero gradients (conv bias before train-mode BN)
            e = float((p.grad - q.grad).abs().max())
        if e > worst[1]: worst = (n, e)
        if e > thr: print(f"   {tag} grad {n}: rel {e:.2e}  |ref| {float(q.grad.norm()):.3e}")
    print(f"  {tag} worst grad err: {worst[0]} {worst[1]:.2e}")

if which in ("all", "lstm"):
    T, B = 31, 20
    torch.manual_seed(0)
    lstm = torch.nn.LSTM(512, 256, 1, bidirectional=True).to(dev)
    x = torch.randn(T, B, 512, device=dev)
    y_ref, _ = lstm(x)
    dy = torch.randn_like(y_ref)
    y_ref.backward(dy)
    gates = torch.empty(T, B, 2, 1024, device=dev)
    gates[:, :, 0] = x @ lstm.weight_ih_l0.T + lstm.bias_ih_l0 + lstm.bias_hh_l0
    gates[:, :, 1] = x @ lstm.weight_ih_l0_reverse.T + lstm.bias_ih_l0_reverse + lstm.bias_hh_l0_reverse
    cells = torch.empty(T, B, 2, 256, device=dev)
    y = torch.empty(T, B, 512, device=dev)
    _lib.call("qeb_lstm_layer_fwd", gates.data_ptr(), lstm.weight_hh_l0.data_ptr(), lstm.weight_hh_l0_reverse.data_ptr(),
              cells.data_ptr(), y.data_ptr(), T, B, _lib.stream())
    torch.cuda.synchronize()
    print("lstm fwd rel", rel(y, y_ref))
    _lib.call("qeb_lstm_layer_bwd", gates.data_ptr(), cells.data_ptr(), dy.contiguous().data_ptr(), lstm.weight_hh_l0.data_ptr(),
              lstm.weight_hh_l0_reverse.data_ptr(), T, B, _lib.stream())
    torch.cuda.synchronize()
    dg = gates  # d(pre-activations)
    dwih = dg[:, :, 0].reshape(T * B, 1024).T @ x.reshape(T * B, 512)
    print("lstm bwd dW_ih rel", rel(dwih, lstm.weight_ih_l0.grad), " db rel", rel(dg[:, :, 1].sum((0, 1)), lstm.bias_ih_l0_reverse.grad))
    hprev = torch.zeros(T, B, 256, device=dev); hprev[1:] = y[:-1, :, :256]
    dwhh = dg[:, :, 0].reshape(T * B, 1024).T @ hprev.reshape(T * B, 256)
    print("lstm bwd dW_hh rel", rel(dwhh, lstm.weight_hh_l0.grad))

if which in ("all", "crnn"):
    for (B, W, mode) in ((4, 128, "train"), (64, 128, "train"), (8, 128, "bneval"), (3, 64, "train")):
        torch.manual_seed(1)
        m = CRNN(95, False).to(dev)
        with torch.no_grad():  # non-trivial BN statistics / affine
            for bn in (m.convo.batchnorm1, m.convo.batchnorm2):
                bn.running_mean.normal_(0, 0.1); bn.running_var.uniform_(0.5, 1.5); bn.weight.uniform_(0.5, 1.5); bn.bias.normal_(0, 0.1)
        mr = copy.deepcopy(m)
        m.train(); mr.train()
        if mode == "bneval":
            for mm in (m, mr):
                mm.convo.batchnorm1.eval(); mm.convo.batchnorm2.eval()
        x = torch.rand(B, 1, 32, W, device=dev)
        xr = x.clone().requires_grad_(True); xq = x.clone().requires_grad_(True)
        lp = m(xq)
        lpr = torch_ref.crnn_forward(mr, xr)
        print(f"crnn B{B} W{W} {mode}: log_probs rel {rel(lp, lpr):.2e} max abs {float((lp - lpr).abs().max()):.2e}")
        g = torch.randn_like(lp) / lp.numel() ** 0.5
        lp.backward(g); lpr.backward(g)
        print(f"  dx rel {rel(xq.grad, xr.grad):.2e}")
        cmp_grads(m, mr, "crnn")
        # calibration: torch's own TF32 path against its fp32 path on the same problem
        mt = copy.deepcopy(mr); mt.zero_grad()
        if mode == "train":
            with torch.no_grad():
                for bn in (mt.convo.batchnorm1, mt.convo.batchnorm2): pass
        torch.backends.cudnn.allow_tf32 = True; torch.backends.cuda.matmul.allow_tf32 = True
        xt = x.clone().requires_grad_(True)
        lpt = torch_ref.crnn_forward(mt, xt); lpt.backward(g)
        torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
        print(f"  [torch tf32 vs fp32] log_probs rel {rel(lpt, lpr):.2e} dx rel {rel(xt.grad, xr.grad):.2e}")
        cmp_grads(mt, mr, "torch-tf32", thr=1.0)
        if mode == "train":
            print("  running_mean rel", rel(m.convo.batchnorm1.running_mean, mr.convo.batchnorm1.running_mean),
                  "running_var rel", rel(m.convo.batchnorm2.running_var, mr.convo.batchnorm2.running_var),
                  "nbt", int(m.convo.batchnorm1.num_batches_tracked))

if which in ("all", "unet"):
    for (B, H, W, mode) in ((2, 32, 128, "train"), (16, 32, 128, "train"), (1, 64, 96, "train"), (2, 32, 128, "eval")):
        torch.manual_seed(2)
        m = UNet().to(dev)
        with torch.no_grad():
            for mod in m.modules():
                if isinstance(mod, torch.nn.BatchNorm2d):
                    mod.running_mean.normal_(0, 0.1); mod.running_var.uniform_(0.5, 1.5); mod.weight.uniform_(0.5, 1.5); mod.bias.normal_(0, 0.1)
        mr = copy.deepcopy(m)
        if mode == "train": m.train(); mr.train()
        else: m.eval(); mr.eval()
        x = torch.rand(B, 1, H, W, device=dev)
        xq = x.clone().requires_grad_(True); xr = x.clone().requires_grad_(True)
        y = m(xq); yr = torch_ref.unet_forward(mr, xr)
        print(f"unet B{B} {H}x{W} {mode}: y rel {rel(y, yr):.2e} max abs {float((y - yr).abs().max()):.2e}")
        g = torch.randn_like(y)
        y.backward(g); yr.backward(g)
        print(f"  dx rel {rel(xq.grad, xr.grad):.2e}")
        cmp_grads(m, mr, "unet")
        if mode == "train":
            print("  enc1norm1 running_mean rel", rel(m.encoder1.enc1norm1.running_mean, mr.encoder1.enc1norm1.running_mean),
                  "dec1norm2 running_var rel", rel(m.decoder1.dec1norm2.running_var, mr.decoder1.dec1norm2.running_var))
print("launches", _lib.launch_count())
