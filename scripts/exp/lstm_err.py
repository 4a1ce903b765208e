"""Experiment: error of the tensor-core LSTM recurrence against cuDNN fp32, forward and backward."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import qeb_b200
from qeb_b200 import _lib
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
DEV = "cuda"
st = torch.cuda.current_stream().cuda_stream
def rel(a, b):
    a, b = a.detach().double(), b.detach().double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))
for T, B in [(31, 64), (31, 20), (15, 3), (63, 9)]:
    torch.manual_seed(T + B)
    lstm = torch.nn.LSTM(512, 256, 1, bidirectional=True).to(DEV)
    x = torch.randn(T, B, 512, device=DEV)
    y_ref, _ = lstm(x)
    dy = torch.randn_like(y_ref)
    y_ref.backward(dy)
    gates = torch.empty(T, B, 2, 1024, device=DEV)
    with torch.no_grad():
        gates[:, :, 0] = x @ lstm.weight_ih_l0.T + lstm.bias_ih_l0 + lstm.bias_hh_l0
        gates[:, :, 1] = x @ lstm.weight_ih_l0_reverse.T + lstm.bias_ih_l0_reverse + lstm.bias_hh_l0_reverse
    cells = torch.empty(T, B, 2, 256, device=DEV); y = torch.empty(T, B, 512, device=DEV)
    _lib.call("qeb_lstm_layer_fwd", gates.data_ptr(), lstm.weight_hh_l0.data_ptr(), lstm.weight_hh_l0_reverse.data_ptr(), cells.data_ptr(), y.data_ptr(), T, B, st)
    e_y = rel(y, y_ref)
    _lib.call("qeb_lstm_layer_bwd", gates.data_ptr(), cells.data_ptr(), dy.contiguous().data_ptr(), lstm.weight_hh_l0.data_ptr(), lstm.weight_hh_l0_reverse.data_ptr(), T, B, st)
    dg = gates.reshape(T * B, 2, 1024); xf = x.reshape(T * B, 512)
    hprev = torch.zeros(T, B, 256, device=DEV); hprev[1:] = y[:-1, :, :256]
    hnext = torch.zeros(T, B, 256, device=DEV); hnext[:-1] = y[1:, :, 256:]
    print(T, B, "y %.1e" % e_y, "dWih %.1e %.1e" % (rel(dg[:, 0].T @ xf, lstm.weight_ih_l0.grad), rel(dg[:, 1].T @ xf, lstm.weight_ih_l0_reverse.grad)),
          "db %.1e" % rel(dg[:, 1].sum(0), lstm.bias_hh_l0_reverse.grad),
          "dWhh %.1e %.1e" % (rel(dg[:, 0].T @ hprev.reshape(T * B, 256), lstm.weight_hh_l0.grad), rel(dg[:, 1].T @ hnext.reshape(T * B, 256), lstm.weight_hh_l0_reverse.grad)))
