import sys, os, json, torch
sys.path.insert(0, '/root/repo')
import qeb_b200
from qeb_b200 import _lib
from qeb_b200.mirror.models.model_crnn import CRNN
from qeb_b200.mirror import ctc as qctc, utils as qutils
torch.manual_seed(0)
m = CRNN(95, False).cuda(); m.train(); m.apply(qutils.set_bn_eval)
x = torch.rand(64,1,32,128, device='cuda')
for fused in (0,1):
    import qeb_b200.mirror.models.model_crnn as mc
    mc._FUSED_HEAD = bool(fused)
    with torch.no_grad():
        for _ in range(3): m(x)
        _lib.prof_enable(True); torch.cuda.synchronize(); _lib.prof_report()
        for _ in range(20): m(x)
        torch.cuda.synchronize(); rep = _lib.prof_report(); _lib.prof_enable(False)
    print('fused', fused, {k: (v['launches']//20, round(v['ms']/20*1e3,1)) for k,v in rep.items() if k.startswith('tc_') or 'softmax' in k})
