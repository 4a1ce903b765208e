"""Small run of the non-network kernels and of the fused / window variants for compute-sanitizer (scripts/sanitize.sh)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import qeb_b200
from qeb_b200.mirror import ctc as qctc, selection_utils, transform_helper as th, utils as qutils
from qeb_b200.mirror.models.model_crnn import CRNN
dev = "cuda"
torch.manual_seed(2)
labels = ["abc", "", "hello world", "x" * 40, "€uro"]
preds = ["abd", "q", "hello wrld", "x" * 35, "euro"]
print(qutils.compare_labels(preds, labels))
v = torch.rand(8, 64)
print(selection_utils._segmented(list(v.numpy()), [32] * 8, torch.device(dev))[0][:4])
print(selection_utils.topk_global_device(torch.rand(5000, device=dev), 17)[:4])
m = CRNN(95, False).to(dev); m.train()
x = torch.rand(4, 1, 32, 64, device=dev)
lp, noisy = m.forward_jittered(x, torch.full((4,), 0.03), seed=5)          # jitter fused into conv1, log-softmax head
y = torch.randint(1, 95, (12,), dtype=torch.int32); yl = torch.tensor([3, 3, 3, 3], dtype=torch.int32)
loss = qctc.CTCLoss()(lp, y, torch.full((4,), 15, dtype=torch.int32), yl)
loss.backward()
print(float(loss), qutils.decode_batch(lp)[1].tolist())
torch.cuda.synchronize()
print("done")
