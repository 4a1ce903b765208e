"""Context measurement (not a bench value): the same phase-B step on torch eager / cuDNN on this GPU."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import qeb_b200
from oracle import nn_oracle
from qeb_b200.mirror.models.model_crnn import CRNN
from qeb_b200.mirror.models.model_unet import UNet
from qeb_b200.mirror.utils import set_bn_eval
dev = "cuda"
for tf32 in (True, False):
    torch.backends.cudnn.allow_tf32 = tf32; torch.backends.cuda.matmul.allow_tf32 = tf32
    torch.backends.cudnn.benchmark = True
    torch.manual_seed(42)
    prep, crnn = UNet().to(dev), CRNN(95, False).to(dev)
    crnn.lstm.flatten_parameters()
    opt = torch.optim.Adam(prep.parameters(), lr=5e-5)
    ctc, mse = torch.nn.CTCLoss(), torch.nn.MSELoss()
    x, labels = bench.synth_batch(64, 7); x = x.to(dev)
    c2i = {c: i for i, c in enumerate(bench.CHAR_SET)}
    y, ys = bench.encode(labels, c2i); ps = torch.tensor([31] * 64, dtype=torch.int32)
    y, ys, ps = y.to(dev), ys.to(dev), ps.to(dev)
    ones = torch.ones(64, 1, 32, 128, device=dev)
    def step():
        prep.train(); crnn.train(); crnn.apply(set_bn_eval)
        prep.zero_grad(); crnn.zero_grad()
        img = nn_oracle.unet_forward(prep, x)
        scores = nn_oracle.crnn_forward(crnn, img)
        loss = ctc(scores, y, ps, ys) + mse(img, ones)
        loss.backward(); opt.step()
    for _ in range(5): step()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20): step()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 20
    print(f"torch eager cuDNN tf32={tf32}: {dt*1e3:.2f} ms/step, {64/dt:.0f} patches/s")
