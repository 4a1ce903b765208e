"""Per-tensor parity of every UNet / CRNN weight gradient against the fp32 oracle (oracle/nn_oracle.py on torch CPU).

  python scripts/grad_parity.py [--batch 16] [--out gpurun_out/grad_parity.json]

Phase B (train_nn_area.py:277-287): UNet train-mode BN -> CRNN (train(), BN frozen) -> CTC + MSE; gradients of the UNet.
Phase A (train_nn_area.py:245-275): CRNN train mode -> CTC; gradients of the CRNN.
Prints rel-L2 and cosine per tensor, worst first; the JSON feeds bench.py's `parity` block and DESIGN.md.
"""
import argparse
import copy
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def cos(a, b):
    a, b = a.detach().double().cpu().reshape(-1), b.detach().double().cpu().reshape(-1)
    return float((a @ b) / (a.norm() * b.norm()).clamp_min(1e-30))


def measure(batch=16, seed=0, competitors=False, prime=False):
    import qeb_b200  # noqa: F401
    from bench import CHAR_SET, encode, synth_batch
    from oracle import nn_oracle
    from qeb_b200.mirror import ctc as qctc
    from qeb_b200.mirror import train_ops
    from qeb_b200.mirror.models.model_crnn import CRNN
    from qeb_b200.mirror.models.model_unet import UNet
    from qeb_b200.mirror.utils import set_bn_eval

    dev = torch.device("cuda", 0)
    torch.manual_seed(seed)
    V = len(CHAR_SET)
    crnn_cpu, unet_cpu = CRNN(V, False), UNet()
    crnn, unet = copy.deepcopy(crnn_cpu).to(dev), copy.deepcopy(unet_cpu).to(dev)
    x, labels = synth_batch(batch, 11)
    c2i = {c: i for i, c in enumerate(CHAR_SET)}
    y, y_size = encode(labels, c2i)
    il = torch.full((batch,), 31, dtype=torch.int32)
    out = {}

    # ---- phase B
    for m in (crnn_cpu, unet_cpu, crnn, unet):
        m.train()
    crnn_cpu.apply(set_bn_eval); crnn.apply(set_bn_eval)
    if prime:
        # a network's FIRST backward call reads tf32 operands and records the gradient maxima; from the second call on the conv
        # backward contractions read scaled fp16 shadows (csrc/nn.cuh GradShadow). Measure that second call: one throw-away step on
        # another batch, BatchNorm buffers restored.
        keep = copy.deepcopy(unet.state_dict())
        x0, _ = synth_batch(batch, 12)
        img0 = unet(x0.to(dev)); sc0 = crnn(img0)
        (qctc.CTCLoss()(sc0, y, il, y_size) + train_ops.mse_to_ones(img0)).backward()
        unet.zero_grad(); crnn.zero_grad(); unet.load_state_dict(keep)
    img = unet(x.to(dev)); scores = crnn(img)
    loss = qctc.CTCLoss()(scores, y, il, y_size) + train_ops.mse_to_ones(img)
    loss.backward()
    img_r = nn_oracle.unet_forward(unet_cpu, x); scores_r = nn_oracle.crnn_forward(crnn_cpu, img_r)
    loss_r = torch.nn.CTCLoss()(scores_r, y, il, y_size) + torch.nn.MSELoss()(img_r, torch.ones_like(img_r))
    loss_r.backward()
    rows = {}
    for (n, p), (_, pr) in zip(unet.named_parameters(), unet_cpu.named_parameters()):
        rows["unet." + n] = [rel(p.grad, pr.grad), cos(p.grad, pr.grad)]
    out["phase_b"] = {"loss_rel": abs(float(loss) - float(loss_r)) / abs(float(loss_r)), "img_rel": rel(img, img_r),
                      "scores_rel": rel(scores, scores_r), "grads": rows}
    if competitors:
        # the same step on torch eager + cuDNN on this GPU (the competing implementation on the same box), fp32 and TF32
        # operands, and the fp32 CPU oracle against float64 (what "exact fp32" itself is off by on this problem)
        def lib_step(un, cr, xx, dt=torch.float32):
            un.train(); cr.train(); cr.apply(set_bn_eval)
            i_ = nn_oracle.unet_forward(un, xx); s_ = nn_oracle.crnn_forward(cr, i_)
            (torch.nn.functional.ctc_loss(s_, y.to(xx.device), il.to(xx.device), y_size.to(xx.device)) +
             torch.nn.functional.mse_loss(i_, torch.ones_like(i_))).backward()
            return {"unet." + n: p.grad for n, p in un.named_parameters()}
        ref_g = {"unet." + n: p.grad for n, p in unet_cpu.named_parameters()}
        comp = {}
        for name, tf32 in (("cudnn_fp32", False), ("cudnn_tf32", True)):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            g_ = lib_step(copy.deepcopy(unet_cpu).to(dev), copy.deepcopy(crnn_cpu).to(dev), x.to(dev))
            comp[name] = {k: rel(g_[k], ref_g[k]) for k in ref_g}
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        g64 = lib_step(copy.deepcopy(unet_cpu).double(), copy.deepcopy(crnn_cpu).double(), x.double())
        comp["fp32_oracle_vs_fp64"] = {k: rel(ref_g[k], g64[k]) for k in ref_g}
        comp["qeb_vs_fp64"] = {k: rel(p.grad, g64["unet." + n]) for (n, p), k in zip(unet.named_parameters(), ref_g)}
        out["phase_b"]["competitors_worst"] = {k: max(v.values()) for k, v in comp.items()}
        out["phase_b"]["competitors_median"] = {k: sorted(v.values())[len(v) // 2] for k, v in comp.items()}

    # ---- phase A
    for m in (crnn_cpu, crnn):
        m.zero_grad(); m.train()
    xa = x.to(dev)
    if prime:
        keep = copy.deepcopy(crnn.state_dict())
        x0, _ = synth_batch(batch, 12)
        qctc.CTCLoss()(crnn(x0.to(dev)), y, il, y_size).backward()
        crnn.zero_grad(); crnn.load_state_dict(keep)
    scores = crnn(xa)
    loss = qctc.CTCLoss()(scores, y, il, y_size)
    loss.backward()
    scores_r = nn_oracle.crnn_forward(crnn_cpu, x)
    loss_r = torch.nn.CTCLoss()(scores_r, y, il, y_size)
    loss_r.backward()
    rows = {}
    for (n, p), (_, pr) in zip(crnn.named_parameters(), crnn_cpu.named_parameters()):
        rows["crnn." + n] = [rel(p.grad, pr.grad), cos(p.grad, pr.grad)]
    out["phase_a"] = {"loss_rel": abs(float(loss) - float(loss_r)) / abs(float(loss_r)), "scores_rel": rel(scores, scores_r), "grads": rows}
    for ph in ("phase_a", "phase_b"):
        g = out[ph]["grads"]
        worst = max(g, key=lambda k: g[k][0])
        out[ph]["worst_tensor"] = worst
        out[ph]["worst_rel_l2"] = g[worst][0]
        out[ph]["min_cos"] = min(v[1] for v in g.values())
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--out", default=None)
    ap.add_argument("--competitors", action="store_true", help="also torch eager + cuDNN (fp32 / TF32) on this GPU and the fp64 floor")
    ap.add_argument("--prime", action="store_true", help="measure the SECOND backward call of each network (scaled fp16 gradient operands)")
    a = ap.parse_args()
    res = measure(a.batch, competitors=a.competitors, prime=a.prime)
    res["primed"] = bool(a.prime)
    for ph in ("phase_b", "phase_a"):
        r = res[ph]
        print(f"== {ph}: loss rel {r['loss_rel']:.2e}, scores rel {r['scores_rel']:.2e}, worst {r['worst_tensor']} {r['worst_rel_l2']:.2e}, min cos {r['min_cos']:.6f}")
        for k, v in sorted(r["grads"].items(), key=lambda kv: -kv[1][0])[:12]:
            print(f"   {k:44s} rel {v[0]:.2e} cos {v[1]:.6f}")
    if "competitors_worst" in res["phase_b"]:
        print("phase-B UNet weight gradients, rel-L2 against the fp32 CPU oracle (worst / median tensor):")
        for k in res["phase_b"]["competitors_worst"]:
            print(f"   {k:22s} {res['phase_b']['competitors_worst'][k]:.2e} / {res['phase_b']['competitors_median'][k]:.2e}")
        g = res["phase_b"]["grads"]
        print(f"   {'qeb':22s} {max(v[0] for v in g.values()):.2e} / {sorted(v[0] for v in g.values())[len(g) // 2]:.2e}")
    if a.out:
        json.dump(res, open(a.out, "w"), indent=1)
