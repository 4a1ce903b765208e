"""TEST INFRASTRUCTURE ONLY (oracle): plain PyTorch fp32 restatement of the two networks' forward passes.

The functions take a module whose sub-modules are the standard torch layers with the reference's names (the mirror
modules' parameter containers qualify, and so do the reference's own modules) and evaluate the reference graph with
torch's library ops - on CPU this is the "reference modules executed on torch CPU fp32" oracle of SURVEY.md 8(c).
  crnn_logits / crnn_forward  models/model_crnn.py:16-28,47-56 (conv stack, map_to_sequence, LSTM, Linear, log_softmax)
  unet_forward                models/model_unet.py:49-76
Pinned against the real reference by tests/test_oracle_cpu.py (golden fixtures crnn.npz / unet.npz written by
oracle/gen_golden.py from the unmodified reference modules). Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this file; the product path never does."""
import torch
import torch.nn.functional as F


def crnn_logits(m, x):
    c = m._stack()
    x = F.max_pool2d(F.relu(c.conv1(x)), (2, 2))
    x = F.max_pool2d(F.relu(c.conv2(x)), (2, 2))
    x = F.relu(c.conv3(x))
    x = F.max_pool2d(F.relu(c.conv4(x)), (2, 1))
    x = F.relu(c.batchnorm1(c.conv5(x)))
    x = F.relu(c.batchnorm2(c.conv6(x)))
    x = F.max_pool2d(x, (2, 1))
    x = c.conv7(x)
    b, ch, h, w = x.shape
    x = x.permute(3, 0, 1, 2).reshape(w, b, ch * h)
    x, _ = m.lstm(x)
    return m.linear(x)


def crnn_forward(m, x):
    return F.log_softmax(crnn_logits(m, x), 2)


def unet_forward(m, x):
    e1 = m.encoder1(x)
    e2 = m.encoder2(m.pool1(e1))
    e3 = m.encoder3(m.pool2(e2))
    e4 = m.encoder4(m.pool3(e3))
    bt = m.bottleneck(m.pool4(e4))
    d4 = m.decoder4(torch.cat((m.upconv4(bt), e4), dim=1))
    d3 = m.decoder3(torch.cat((m.upconv3(d4), e3), dim=1))
    d2 = m.decoder2(torch.cat((m.upconv2(d3), e2), dim=1))
    d1 = m.decoder1(torch.cat((m.upconv1(d2), e1), dim=1))
    return torch.sigmoid(m.conv(d1))
