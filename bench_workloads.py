#!/usr/bin/env python
"""The BASELINE.json configs beside the headline, as `bench.py --workload ...` (same output contract, one JSON line):

  jitter_step  configs[2]  the black-box-approximation step on jittered patches (train_nn_patch.py:278-309,
                           transform_helper.py:33-45): 64 base patches x inner_limit = 8 noised copies, each copy one
                           CRNN train-mode call (its own BatchNorm batch statistics) + CTC(mean) against that copy's OCR
                           strings + backward; the copies' gradients are SUMMED (the patch trainer calls backward() inside
                           the loop, :301-303), then ONE Adam step (lr 1e-4, weight decay 5e-4). The copies are sharded over the
                           ranks (8 / N each - strong scaling, the job is always 512 patches) and the flat 35 MB CRNN gradient
                           is all-reduced with SUM.
  area_step    configs[3]  one whole minibatch of train_nn_area.py:212-304: phase A (UNet eval -> TopKCER query -> jitter ->
                           CRNN train step on the synthetic OCR strings), phase B (the headline's UNet update through the
                           frozen-BN surrogate) and phase C (greedy decode -> per-sample CER -> sampler.update_cer), 64
                           VGG-style patches (alphanumeric words, 1-23 symbols) per rank, both gradient all-reduces (AVG).
  cer_topk     configs[4]  Levenshtein + CER of 1 M (prediction, ground truth) pairs, TopKCER selection of 15,625 minibatches of
                           64 (utils.py:95-110, selection_utils.py:144-151) and the dataset-wide top-k of
                           pruning/methods.py:5-8; integer path, bit-exact.

The external OCR engine is replaced by cached synthetic strings (north_star); everything else of a step is inside the
timed region. `--impl reference` runs the same workload on the host cores with the UNMODIFIED reference modules /
functions from baseline/_ref when that copy is present (else the oracle port).
"""
import json
import math
import os
import random
import time

import numpy as np
import torch

import bench as BB

INNER_LIMIT = 8                      # BASELINE.json configs[2]
LR_CRNN, WD_PATCH = 1e-4, 5e-4       # patch_cli.py defaults; the area trainer uses weight_decay 0 (train_nn_area.py:149-154)
NOISE_STD = 5                        # --std 5, --random_std (patch_cli.py / area_cli.py)
SUBSET_PROP = 0.5                    # --minibatch_subset_prop
GFLOP_CRNN_TRAIN = 3 * 1.778         # per patch, SURVEY.md 8(d): fprop + dgrad + wgrad (conv1's dgrad is not computed)
N_PAIRS, SEG = 1_000_000, 64         # configs[4]
ALNUM = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789"


# ------------------------------------------------------------------------------------------------------ synthetic data
def ocr_strings(labels, seed):
    """Synthetic "OCR" output: the ground truth with seeded random edits at about the CER fixture's rate (55 % exact)."""
    rng = random.Random(seed)
    out = []
    for l in labels:
        if rng.random() < 0.55:
            out.append(l)
            continue
        s = list(l)
        for _ in range(1 + int(rng.random() < 0.3)):
            op = rng.random()
            pos = rng.randrange(len(s) + 1)
            if op < 0.4 and s:
                s[min(pos, len(s) - 1)] = rng.choice(BB.CHAR_SET[1:])
            elif op < 0.7 and len(s) > 1:
                del s[min(pos, len(s) - 1)]
            elif len(s) < 24:
                s.insert(pos, rng.choice(BB.CHAR_SET[1:]))
        out.append("".join(s) or l)
    return out


def vgg_labels(n, seed, T=31):
    """VGG-style words: alphanumeric, 1-23 symbols (mean ~7.5), feasible for CTC at T frames (len + repeats <= T)."""
    rng = random.Random(seed)
    out = []
    while len(out) < n:
        ln = max(1, min(23, int(rng.gauss(7.5, 3.5))))
        w = "".join(rng.choice(ALNUM) for _ in range(ln))
        if ln + sum(a == b for a, b in zip(w, w[1:])) <= T:
            out.append(w)
    return out


def cer_pairs(n, seed):
    """n (prediction, ground truth) pairs as CSR int32 code points + the CER histogram of the POS fixture (55 % zeros):
    ground truth 1-16 symbols of the char set, prediction = ground truth with substitutions / deletions / insertions."""
    rng = np.random.default_rng(seed)
    cs = np.array([ord(c) for c in BB.CHAR_SET[1:]], dtype=np.int32)
    la = rng.integers(1, 17, n)
    a = cs[rng.integers(0, len(cs), int(la.sum()))]
    aoff = np.zeros(n + 1, dtype=np.int32)
    np.cumsum(la, out=aoff[1:])
    noisy = np.repeat(rng.random(n) >= 0.55, la)                 # pairs that carry edits
    r = rng.random(len(a))
    sub = noisy & (r < 0.10)
    dele = noisy & (r >= 0.10) & (r < 0.16)
    ins = noisy & (r >= 0.16) & (r < 0.22)
    b = a.copy()
    b[sub] = cs[rng.integers(0, len(cs), int(sub.sum()))]
    rep = np.ones(len(a), dtype=np.int64)
    rep[dele] = 0
    rep[ins] = 2
    pair_id = np.repeat(np.arange(n), la)
    b = np.repeat(b, rep)
    lb = np.bincount(np.repeat(pair_id, rep), minlength=n)
    boff = np.zeros(n + 1, dtype=np.int32)
    np.cumsum(lb, out=boff[1:])
    return a, aoff, b.astype(np.int32), boff


def csr_to_strings(flat, off):
    s = flat.astype("<u4").tobytes().decode("utf-32-le")
    return [s[off[i]:off[i + 1]] for i in range(len(off) - 1)]


# ------------------------------------------------------------------------------------------------------ shared plumbing
class Env:
    def __init__(self, args, rank, world, local_rank, own_process_group):
        import torch.distributed as dist
        from qeb_b200 import _lib
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device - the qeb hot path has no CPU fallback (use --impl reference for the CPU arm)")
        self.args, self.rank, self.world, self.local_rank, self.dist = args, rank, world, local_rank, dist
        torch.cuda.set_device(local_rank)
        self.dev = torch.device("cuda", local_rank)
        self.own = own_process_group and world > 1
        if self.own:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
            dist.init_process_group("nccl", device_id=self.dev)
        assert _lib.load().qeb_check_device() == 0, _lib.load().qeb_last_error()
        self.lib = _lib

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def timed(self, fn, steps):
        """EXACTLY `steps` calls between two barriers + synchronisations, device time from CUDA events, max over ranks."""
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = None
        for _ in range(steps):
            out = fn()
        e1.record()
        self.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(ms, op=self.dist.ReduceOp.MAX)
        return float(ms) / steps, (float(out.detach()) if torch.is_tensor(out) and out.numel() == 1 else None if torch.is_tensor(out) else out)

    def profile(self, fn, steps):
        """Per-kernel-family device times of `steps` calls (CUDA events around every launch of the library, side stream off)."""
        self.lib.prof_enable(True)
        torch.cuda.synchronize()
        self.lib.prof_report()
        for _ in range(steps):
            fn()
        torch.cuda.synchronize()
        rep = self.lib.prof_report()
        self.lib.prof_enable(False)
        return rep

    def close(self):
        if self.own:
            self.dist.destroy_process_group()


def peaks():
    try:
        return json.load(open(os.path.join(BB.ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        return {}


def roofline_of(rep, psteps, prefer=None, bytes_override=None):
    """The family with the largest share of the profiled step (or `prefer`) against the measured peak of its bound."""
    if not rep:
        return None, None
    rep, kinds = BB.fold_kinds(rep)
    total = sum(v["ms"] for v in rep.values()) or 1e-9
    kernels = {k: {"launches_per_step": v["launches"] / psteps, "ms_per_step": v["ms"] / psteps, "share": v["ms"] / total,
                   "tflops": v["flops"] / v["ms"] / 1e9 if v["ms"] > 0 else 0.0,
                   "gbs": v["bytes"] / v["ms"] / 1e6 if v["ms"] > 0 else 0.0} for k, v in rep.items()}
    top = prefer if prefer in rep else max(rep, key=lambda k: rep[k]["ms"])
    v, pk = rep[top], peaks()
    if top.startswith("tc_"):
        peak, ach, unit, bound = pk.get("bf16_tflops", 1590.0), v["flops"] / v["ms"] / 1e9, "TFLOP/s", "tensor"
        src = "measured bf16 dense burst (MEASURED_PEAKS.json)" if pk else "fallback 1590 bf16"
    else:
        nbytes = bytes_override * v["launches"] if bytes_override is not None else v["bytes"]
        peak, ach, unit, bound = pk.get("hbm_gbs", 6650.0), nbytes / v["ms"] / 1e6, "GB/s", "hbm"
        src = "measured copy bandwidth (MEASURED_PEAKS.json)" if pk else "fallback 6650"
    extra = {"by_operand_kind": BB.by_operand_kind(kinds[top], BB.tensor_peaks(), psteps)} if top in kinds else {}
    rf = {"kernel": top, "bound": bound, "achieved": ach, "peak": peak, "unit": unit, "frac": ach / peak, "traffic": None,
          "peak_source": src, "share_of_step": v["ms"] / total, "launches_per_step": v["launches"] / psteps,
          "avg_launch_ms": v["ms"] / v["launches"]}
    rf.update(extra)
    return rf, kernels


def cpu_time(step, units, what, max_s=15.0, min_n=2, max_n=12, threads=None, kind="port"):
    threads = threads or (os.cpu_count() or 1)
    torch.set_num_threads(threads)
    step()
    t0, n = time.perf_counter(), 0
    while n < min_n or (time.perf_counter() - t0 < max_s and n < max_n):
        step()
        n += 1
    dt = time.perf_counter() - t0
    return {"value": units * n / dt, "unit": None, "cores": threads, "kind": kind, "sample": f"{n} x {what} ({dt:.1f} s)"}


reference_modules = BB.reference_modules


def line_of(env, name, metric, unit, units_per_step, ms_step, ms_e2e, h2d, d2h, launches, clocks, config, dtype, scaling,
            roofline, kernels, cpu_baseline, extra=None):
    a = env.args
    value = units_per_step / (ms_step / 1e3)
    if cpu_baseline is not None:
        cpu_baseline["unit"] = unit
    line = {"metric": metric, "value": value, "unit": unit, "n_gpus": env.world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": dtype,
            "data": "synthetic", "config": config,
            "e2e": {"value": units_per_step / (ms_e2e / 1e3), "unit": unit, "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h},
            "gpu_launches": launches * a.steps, "gpu_launches_per_step": launches, "clocks": clocks, "roofline": roofline,
            "kernels": kernels, "cpu_baseline": cpu_baseline}
    line.update(extra or {})
    return line


# ====================================================================================================== jitter_step
def jitter_config(world):
    return {"workload": "configs[2]: jittered black-box approximation step - 64 base patches x inner_limit 8 noised copies (sigma = "
                        "randint(0..5)/100 per image) through CRNN(train, batch-stat BN per copy) + CTC(mean) vs synthetic OCR strings, "
                        "gradients summed over the copies, Adam(lr 1e-4, wd 5e-4) on the CRNN; V=95, T=31",
            "copies": INNER_LIMIT, "copies_per_gpu": INNER_LIMIT // world, "patches_per_step": BB.BATCH * INNER_LIMIT,
            "patch": [BB.H, BB.W], "parallelism": f"dp{world} over the noised copies, all-reduce SUM of the flat 35 MB CRNN gradient",
            "operand_precision": "tensor-core operands with an 11-bit significand (fp16 forward copies, power-of-two-scaled fp16 gradient "
                                 "copies in the conv backward, tf32 rounded-to-nearest in the LSTM / Linear backward), fp32 accumulation / "
                                 "activations / gradients / parameters",
            "l2": "no explicit flush: one copy streams ~0.6 GB of activations, > 126 MB L2",
            "launch": "forward (jitter fused into conv1's input load, log-softmax into the head's epilogue) + CTC + backward of the rank's copies replayed as one CUDA graph; all-reduce and Adam outside it"}


def run_jitter(env):
    from qeb_b200.graphs import GraphedStep, StaticTargets
    from qeb_b200.mirror import ctc as qctc, dist as qdist, train_ops, transform_helper as th
    from qeb_b200.mirror.models.model_crnn import CRNN
    from qeb_b200.mirror.models.model_unet import UNet
    a, dev, world, rank = env.args, env.dev, env.world, env.rank
    if INNER_LIMIT % world:
        raise SystemExit(f"jitter_step shards {INNER_LIMIT} noised copies: --gpus must divide {INNER_LIMIT}")
    local = INNER_LIMIT // world
    copies = list(range(rank * local, (rank + 1) * local))
    torch.manual_seed(42)
    prep, crnn = UNet().to(dev), CRNN(len(BB.CHAR_SET), False).to(dev)
    crnn.register_backward_hook(crnn.backward_hook)                       # train_nn_patch.py:94
    ctc_loss = qctc.CTCLoss()
    params = list(crnn.parameters())
    opt = train_ops.Adam(params, lr=LR_CRNN, weight_decay=WD_PATCH)       # train_nn_patch.py:146-148
    c2i = {c: i for i, c in enumerate(BB.CHAR_SET)}
    x_host, labels = BB.synth_batch(BB.BATCH, 7)                          # every rank holds the same 64 base patches
    prep.eval()
    with torch.no_grad():
        base = prep(x_host.to(dev)).detach()                              # phase A0: the preprocessor's output (BN eval)
    base_pin = base.cpu().pin_memory()                                    # `text_crops.detach().cpu()`, train_nn_patch.py:261,270
    ocr = {c: ocr_strings(labels, 100 + c) for c in range(INNER_LIMIT)}   # cached "OCR" strings per noised copy
    g = torch.Generator().manual_seed(5)
    sig_all = (torch.randint(0, NOISE_STD + 1, (INNER_LIMIT, BB.BATCH), generator=g).double() / 100 + 1e-13).float()
    sig = sig_all.to(dev)
    T = BB.W // 4 - 1
    pred_size = torch.tensor([T] * BB.BATCH, dtype=torch.int32)
    tgs = [StaticTargets(BB.BATCH, 24, dev) for _ in copies]
    for tg, c in zip(tgs, copies):
        yy, ys = BB.encode(ocr[c], c2i)
        tg.load(yy, pred_size, ys)
    base_static = base.clone()
    seed_dev = torch.full((1,), 1234 + 1000 * rank, dtype=torch.int64, device=dev)
    noisy = [torch.empty_like(base) for _ in copies]
    crnn.train()

    def fwd_bwd():
        total = None
        for j, c in enumerate(copies):
            # add_noise + _call_model (train_nn_patch.py:289-292) in one pass: the jitter rides conv1's input load, the noisy
            # batch is still materialised (noisy[j]) for the OCR hand-off
            scores, _ = crnn.forward_jittered(base_static, sig[c], seed=j, seed_dev=seed_dev, out=noisy[j])
            loss = ctc_loss(scores, tgs[j])
            loss.backward()                                               # inside the loop: gradients accumulate (:301-303)
            total = loss.detach() if total is None else total + loss.detach()
        return total

    def finish():
        qdist.allreduce_grads(params, average=False)                      # SUM over the ranks' copies
        opt.step()                                                        # :309

    def step_eager():
        crnn.zero_grad(set_to_none=True)
        seed_dev.add_(local)
        out = fwd_bwd()
        finish()
        return out

    ms_eager = None
    if not a.skip_eager:
        for _ in range(a.warmup):
            step_eager()
        ms_eager, _ = env.timed(step_eager, a.steps)
    roofline = kernels = None
    if not a.skip_profile and not a.skip_eager:
        psteps = min(a.steps, 3)
        roofline, kernels = roofline_of(env.profile(step_eager, psteps), psteps)

    gstep = GraphedStep(fwd_bwd, modules=[crnn], warmup=3)

    def step_graph():
        seed_dev.add_(local)                                              # new Philox key per step, read by the captured kernels
        out = gstep()
        finish()
        return out

    # every step copies its own crops from pinned memory and encodes its own OCR strings on the host; graphs.BatchStager issues
    # those one step ahead on a copy stream, behind the running replay (bench.py --e2e-serial: in line on the step's stream)
    from qeb_b200.graphs import BatchStager
    stager = BatchStager(base_static, tgs)

    def host_sets():
        sets = []
        for c in copies:
            yy, ys = BB.encode(ocr[c], c2i)                               # _call_model's host-side label encoding per copy
            sets.append((yy, pred_size, ys))
        return sets

    def step_e2e_serial():
        base_static.copy_(base_pin, non_blocking=True)                    # host crops -> HBM
        for tg, (yy, il_, ys) in zip(tgs, host_sets()):
            tg.load(yy, il_, ys)
        seed_dev.add_(local)
        out = gstep()
        finish()
        return out.item()                                                 # temp_loss += loss.item()

    def step_e2e_staged():
        stager.commit()
        seed_dev.add_(local)
        out = gstep()
        finish()
        stager.stage(base_pin, host_sets())                               # the next step's crops and labels, behind the replay
        return out.item()                                                 # temp_loss += loss.item()

    step_e2e = step_e2e_serial if getattr(a, "e2e_serial", False) else step_e2e_staged
    if step_e2e is step_e2e_staged:
        stager.stage(base_pin, host_sets())

    for _ in range(a.warmup):
        step_graph()
    sampler = BB.ClockSampler(env.local_rank)
    if rank == 0:
        sampler.start()
    ms_step, last = env.timed(step_graph, a.steps)
    clocks = sampler.stop() if rank == 0 else None
    for _ in range(2):
        step_e2e()
    ms_e2e, _ = env.timed(step_e2e, a.steps)

    cpu = None
    if rank == 0 and world == 1 and not a.skip_cpu_baseline:
        cstep, kind, what = jitter_cpu_step(2)
        cpu = cpu_time(cstep, 2 * BB.BATCH, what, kind=kind)
    if rank != 0:
        return None
    units = BB.BATCH * INNER_LIMIT
    h2d = base_pin.numel() * 4 + sum(t._host.numel() * 4 for t in tgs)
    return line_of(env, "jitter_step", "patches/sec per jittered approximation step (CRNN+CTC, 64x8 noised copies)", "patches/s", units,
                   ms_step, ms_e2e, h2d, 4, gstep.launches + 1, clocks, jitter_config(world), "fp16/tf32", "strong", roofline, kernels, cpu,
                   {"loss_sum_of_local_copies": last, "tflops_algorithmic": GFLOP_CRNN_TRAIN * units / (ms_step / 1e3) / 1e3,
                    "variants": None if ms_eager is None else {"eager_modules": {"value": units / (ms_eager / 1e3), "ms_per_step": ms_eager}}})


def jitter_cpu_step(n_copies):
    """n_copies noised copies of the 64 base patches through the reference path on the host: AddGaussianNoice per image on
    the CPU (transform_helper.py:33-45), CRNN.train() forward, CTC, backward inside the loop, one Adam step."""
    UNet, CRNN, _, unet_f, crnn_f, kind, ns = reference_modules()
    torch.manual_seed(42)
    UNet()                                   # same initialisation order as the GPU arm
    crnn = CRNN(len(BB.CHAR_SET), False)
    crnn.register_backward_hook(crnn.backward_hook)
    ctc = torch.nn.CTCLoss()
    opt = torch.optim.Adam(crnn.parameters(), lr=LR_CRNN, weight_decay=WD_PATCH)
    c2i = {c: i for i, c in enumerate(BB.CHAR_SET)}
    x, labels = BB.synth_batch(BB.BATCH, 7)
    if ns is not None:
        noiser = ns.transform_helper.AddGaussianNoice(std=NOISE_STD, is_stochastic=True)
    else:
        def noiser(img):
            r = torch.randint(0, NOISE_STD + 1, (1,)).item() / 100 + 1e-13
            return torch.clamp(img - torch.normal(0.0, r, img.shape), 0, 1)
    ocr = [ocr_strings(labels, 100 + c) for c in range(n_copies)]
    crnn.train()

    def step():
        crnn.zero_grad()
        for c in range(n_copies):
            noisy = torch.stack([noiser(img) for img in x])               # add_noise, train_nn_patch.py:187-191
            scores = crnn_f(crnn, noisy)
            y, ys = BB.encode(ocr[c], c2i)
            loss = ctc(scores, y, torch.tensor([scores.shape[0]] * BB.BATCH, dtype=torch.int32), ys)
            loss.backward()
        opt.step()
        return float(loss)

    return step, kind, f"{n_copies} noised copies of 64 patches per step, torch {torch.__version__} CPU fp32"


# ====================================================================================================== area_step
def area_config(world):
    return {"workload": "configs[3]: whole train_nn_area minibatch - phase A (UNet eval -> TopKCER query, 50 % -> jitter -> CRNN(train) + CTC "
                        "vs synthetic OCR strings -> Adam on the CRNN), phase B (UNet(train BN) -> CRNN(BN frozen) -> CTC + MSE-to-white -> "
                        "Adam on the UNet), phase C (greedy decode -> per-sample Levenshtein CER -> sampler.update_cer); 64 VGG-style "
                        "32x128 patches per GPU (alphanumeric words of 1-23 symbols), inner_limit 1, V=95, T=31",
            "batch_per_gpu": BB.BATCH, "global_batch": BB.BATCH * world, "patch": [BB.H, BB.W],
            "parallelism": f"dp{world}: all-reduce AVG of the flat CRNN gradient (35 MB) after phase A and of the UNet gradient (31 MB) after phase B",
            "operand_precision": "11-bit-significand tensor-core operands, fp32 everything else",
            "l2": "no explicit flush: a step streams > 2 GB of activations",
            "launch": "mirror modules called eagerly, as the unmodified trainer does (the sampler query is host logic between launches)"}


def run_area(env):
    from qeb_b200.mirror import ctc as qctc, dist as qdist, selection_utils, train_ops, transform_helper as th, utils as qutils
    from qeb_b200.mirror.models.model_crnn import CRNN
    from qeb_b200.mirror.models.model_unet import UNet
    a, dev, world, rank = env.args, env.dev, env.world, env.rank
    torch.manual_seed(42)
    prep, crnn = UNet().to(dev), CRNN(len(BB.CHAR_SET), False).to(dev)
    crnn.register_backward_hook(crnn.backward_hook)                       # train_nn_area.py:92
    ctc_loss = qctc.CTCLoss()
    p_crnn, p_prep = list(crnn.parameters()), list(prep.parameters())
    opt_crnn = train_ops.Adam(p_crnn, lr=LR_CRNN, weight_decay=0)         # :149-154
    opt_prep = train_ops.Adam(p_prep, lr=BB.LR_PREP, weight_decay=0)
    overlap_prep = qdist.BucketedAllReduce(prep, average=True)           # UNet exchange overlapped with the encoder's backward
    c2i = {c: i for i, c in enumerate(BB.CHAR_SET)}
    x_host, _ = BB.synth_batch(BB.BATCH, 7 + rank)
    labels = vgg_labels(BB.BATCH, 11 + rank)
    names = [f"r{rank}_{i}_{l}" for i, l in enumerate(labels)]
    ocr_all = ocr_strings(labels, 300 + rank)
    rng = random.Random(3 + rank)
    sampler = selection_utils.datasampler_factory("topKCER")({n: rng.randrange(0, 8) / max(1, len(l)) for n, l in zip(names, labels)})
    noiser = th.AddGaussianNoice(std=NOISE_STD, is_stochastic=True, return_noise=True)
    x_pin, x_dev = x_host.pin_memory(), x_host.to(dev)
    k_sel = max(1, math.ceil(BB.BATCH * (1 - SUBSET_PROP)))               # :221-224
    T = BB.W // 4 - 1
    y_gt, ys_gt = BB.encode(labels, c2i)
    tg_gt = qctc.pack_targets(y_gt, torch.tensor([T] * BB.BATCH, dtype=torch.int32), ys_gt, dev)
    state = {}

    def step(x_in, host_inputs):
        # ---- phase A: the surrogate learns the OCR engine on the selected, jittered patches
        crnn.train(); prep.eval()
        prep.zero_grad(set_to_none=True); crnn.zero_grad(set_to_none=True)
        X = x_in.to(dev, non_blocking=True) if host_inputs else x_in      # :217
        img_all = prep(X)                                                 # :218
        img_sel, _, idx = sampler.query(img_all, labels, k_sel, names)    # :225
        img_sel = img_sel.detach()
        noisy, _ = th.add_noise(img_sel, noiser)                          # :260
        ocr_labels = [ocr_all[i] for i in idx.tolist()]                   # self.ocr.get_labels(noisy_imgs): cached strings
        scores = crnn(noisy)                                              # _call_model :262
        y, ys = BB.encode(ocr_labels, c2i)
        loss_a = ctc_loss(scores, y, torch.tensor([scores.shape[0]] * len(ocr_labels), dtype=torch.int32), ys)   # :265
        loss_a.backward()                                                 # :271
        qdist.allreduce_grads(p_crnn, average=True)
        opt_crnn.step()                                                   # :275
        # ---- phase B: the preprocessor learns through the surrogate
        prep.train(); crnn.train(); crnn.apply(qutils.set_bn_eval)        # :277-279
        prep.zero_grad(set_to_none=True); crnn.zero_grad(set_to_none=True)
        img = prep(X)                                                     # :283
        scores = crnn(img)                                                # :284
        if host_inputs:
            yb, ysb = BB.encode(labels, c2i)
            pri = ctc_loss(scores, yb, torch.tensor([scores.shape[0]] * BB.BATCH, dtype=torch.int32), ysb)
        else:
            pri = ctc_loss(scores, tg_gt)
        loss_b = pri + BB.SCALAR * train_ops.mse_to_ones(img)             # :285
        loss_b.backward()                                                 # :286
        overlap_prep()
        opt_prep.step()                                                   # :287
        # ---- phase C: CER bookkeeping (decode + Levenshtein on the device against the CTC targets already there)
        _, _, _, cer = qutils.decode_and_cer(scores.detach(), tg_gt.tg, tg_gt.offs, tg_gt.tl, 24)   # :290, :297-303
        la, lb = float(loss_a), float(loss_b)                             # temp_loss / training_loss += loss.item()
        cers = cer.tolist()
        sampler.update_cer(cers, names)                                   # :304
        state["cers"] = cers
        return la + lb

    step_dev = lambda: step(x_dev, False)
    step_e2e = lambda: step(x_pin, True)
    for _ in range(a.warmup):
        step_dev()
    n0 = env.lib.launch_count()
    step_dev()
    launches = env.lib.launch_count() - n0
    sampler_clk = BB.ClockSampler(env.local_rank)
    if rank == 0:
        sampler_clk.start()
    ms_step, last = env.timed(step_dev, a.steps)
    clocks = sampler_clk.stop() if rank == 0 else None
    for _ in range(2):
        step_e2e()
    ms_e2e, _ = env.timed(step_e2e, a.steps)
    roofline = kernels = None
    if not a.skip_profile:
        psteps = min(a.steps, 3)
        roofline, kernels = roofline_of(env.profile(step_dev, psteps), psteps)
    cpu = None
    if rank == 0 and world == 1 and not a.skip_cpu_baseline:
        cstep, kind, what = area_cpu_step()
        cpu = cpu_time(cstep, BB.BATCH, what, kind=kind, max_n=8)
    if rank != 0:
        return None
    units = BB.BATCH * world
    gflop = 1.504 + k_sel / BB.BATCH * GFLOP_CRNN_TRAIN + BB.GFLOP_PER_PATCH   # UNet eval fwd + phase A on the subset + phase B
    h2d = x_pin.numel() * 4 + 4 * (len(y_gt) + 3 * BB.BATCH) + 4 * (k_sel * 8 + 3 * k_sel)
    return line_of(env, "area_step", "patches/sec per whole train_nn_area minibatch (phases A+B+C)", "patches/s", units, ms_step, ms_e2e,
                   h2d, 8 + 8 * BB.BATCH + 8 * k_sel, launches, clocks, area_config(world), "fp16/tf32", "weak", roofline, kernels, cpu,
                   {"loss_a_plus_b": last, "mean_cer_after": float(np.mean(state["cers"])), "selected_per_batch": k_sel,
                    "tflops_algorithmic": gflop * units / (ms_step / 1e3) / 1e3})


def area_cpu_step():
    UNet, CRNN, set_bn_eval, unet_f, crnn_f, kind, ns = reference_modules()
    torch.manual_seed(42)
    prep, crnn = UNet(), CRNN(len(BB.CHAR_SET), False)
    crnn.register_backward_hook(crnn.backward_hook)
    ctc, mse = torch.nn.CTCLoss(), torch.nn.MSELoss()
    opt_crnn = torch.optim.Adam(crnn.parameters(), lr=LR_CRNN, weight_decay=0)
    opt_prep = torch.optim.Adam(prep.parameters(), lr=BB.LR_PREP, weight_decay=0)
    c2i = {c: i for i, c in enumerate(BB.CHAR_SET)}
    i2c = {i: c for i, c in enumerate(BB.CHAR_SET)}
    x, _ = BB.synth_batch(BB.BATCH, 7)
    labels = vgg_labels(BB.BATCH, 11)
    names = [f"r0_{i}_{l}" for i, l in enumerate(labels)]
    ocr_all = ocr_strings(labels, 300)
    rng = random.Random(3)
    cers = {n: rng.randrange(0, 8) / max(1, len(l)) for n, l in zip(names, labels)}
    k_sel = max(1, math.ceil(BB.BATCH * (1 - SUBSET_PROP)))
    if ns is not None:
        sampler = ns.selection_utils.datasampler_factory("topKCER")(cers)
        noiser = ns.transform_helper.AddGaussianNoice(std=NOISE_STD, is_stochastic=True)
        pred_to_string, compare_labels = ns.utils.pred_to_string, ns.utils.compare_labels
    else:
        from oracle import pyoracle

        class _S:
            def __init__(self, c):
                self.cers = c

            def query(self, images, labels_, k, names_):
                idx = torch.from_numpy(pyoracle.topk_query(np.array([self.cers[n] for n in names_], dtype=np.float32), k))
                return images[idx], [labels_[i] for i in idx], idx

            def update_cer(self, cs, ns_):
                self.cers.update(zip(ns_, cs))

        sampler = _S(cers)

        def noiser(img):
            r = torch.randint(0, NOISE_STD + 1, (1,)).item() / 100 + 1e-13
            return torch.clamp(img - torch.normal(0.0, r, img.shape), 0, 1)

        pred_to_string = lambda s, l, m: pyoracle.pred_to_string(s.detach().numpy(), m)
        compare_labels = pyoracle.compare_labels

    def step():
        crnn.train(); prep.eval(); prep.zero_grad(); crnn.zero_grad()
        img_all = unet_f(prep, x)
        img_sel, _, idx = sampler.query(img_all, labels, k_sel, names)
        img_sel = img_sel.detach().cpu()
        noisy = torch.stack([noiser(im) for im in img_sel])
        ocr_labels = [ocr_all[i] for i in idx.tolist()]
        scores = crnn_f(crnn, noisy)
        y, ys = BB.encode(ocr_labels, c2i)
        loss = ctc(scores, y, torch.tensor([scores.shape[0]] * len(ocr_labels), dtype=torch.int32), ys)
        loss.backward(); opt_crnn.step()
        prep.train(); crnn.train(); crnn.apply(set_bn_eval); prep.zero_grad(); crnn.zero_grad()
        img = unet_f(prep, x)
        scores = crnn_f(crnn, img)
        y, ys = BB.encode(labels, c2i)
        loss = ctc(scores, y, torch.tensor([scores.shape[0]] * BB.BATCH, dtype=torch.int32), ys) + BB.SCALAR * mse(img, torch.ones(img.shape))
        loss.backward(); opt_prep.step()
        preds = pred_to_string(scores, labels, i2c)
        batch_cers = [compare_labels([preds[i]], [labels[i]])[1] for i in range(len(labels))]
        sampler.update_cer(batch_cers, names)
        return float(loss)

    return step, kind, f"whole minibatches of 64 patches (phases A+B+C), torch {torch.__version__} CPU fp32"


# ====================================================================================================== cer_topk
def cer_config(world, n_pairs):
    n_seg = n_pairs // SEG
    return {"workload": f"configs[4]: Levenshtein distance + CER (fp64) of {n_pairs:,} (prediction, ground truth) pairs per GPU (POS-like: 1-16 "
                        f"symbols, 55 % exact), TopKCER selection of {n_seg:,} minibatches of {SEG} (k = 32) and the stable top-10 % of all CERs",
            "pairs_per_gpu": n_pairs, "segments": n_seg, "segment_size": SEG, "k_segment": SEG // 2, "k_global": n_pairs // 10,
            "symbols": "int32 code points (UTF-32), CSR", "parallelism": f"dp{world}: pairs sharded by range, no data-path collective for the "
            "scoring and the segmented selection; the global top-k merges k candidates per rank (all-gather)",
            "l2": "the 1 M-pair CSR arrays + results are ~60 MB (< L2): 256 MB are written between timed iterations to flush it",
            "launch": "eager C-ABI calls"}


def run_cer(env, n_pairs=N_PAIRS):
    from qeb_b200.mirror import dist as qdist, selection_utils, utils as qutils
    a, dev, world, rank = env.args, env.dev, env.world, env.rank
    n_seg, k_seg, k_glob = n_pairs // SEG, SEG // 2, n_pairs // 10
    lab, lab_off, prd, prd_off = cer_pairs(n_pairs, 42 + rank)
    max_len = int(max(np.diff(lab_off).max(), np.diff(prd_off).max()))
    d_lab, d_laboff = torch.from_numpy(lab).to(dev), torch.from_numpy(lab_off).to(dev)
    d_prd, d_prdoff = torch.from_numpy(prd).to(dev), torch.from_numpy(prd_off).to(dev)
    seg_off = torch.arange(0, n_seg * SEG + 1, SEG, dtype=torch.int32, device=dev)
    seg_k = torch.full((n_seg,), k_seg, dtype=torch.int32, device=dev)
    out_off = torch.arange(0, n_seg * k_seg, k_seg, dtype=torch.int32, device=dev)
    sel = torch.empty(n_seg * k_seg, dtype=torch.int64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    res = {}

    def score_and_select(la, lo, pa, po):
        dist, cer = qutils.cer_batch(la, lo, None, pa, po, None, n_pairs, max_len)             # utils.py:103-109 per pair
        cer32 = cer.to(torch.float32)                                                            # torch.tensor(list_of_floats) -> fp32
        env.lib.call("qeb_cer_topk_segmented", cer32.data_ptr(), seg_off.data_ptr(), seg_k.data_ptr(), out_off.data_ptr(), n_seg,
                     sel.data_ptr(), env.lib.stream())                                           # selection_utils.py:144-151 per minibatch
        top = selection_utils.topk_global_device(cer32, k_glob)                                  # pruning/methods.py:5-8
        res.update(dist=dist, cer=cer, sel=sel, top=top)

    def step_dev():
        flush.zero_()
        score_and_select(d_lab, d_laboff, d_prd, d_prdoff)

    # e2e: host strings -> CSR code points (one C-level utf-32 encode per side) -> H2D -> kernels -> distances, CERs and
    # selections back on the host
    s_lab, s_prd = csr_to_strings(lab, lab_off), csr_to_strings(prd, prd_off)

    def step_e2e():
        flush.zero_()
        lf, lo = qutils._encode_csr(s_lab)
        pf, po = qutils._encode_csr(s_prd)
        hs = [torch.from_numpy(v).pin_memory().to(dev, non_blocking=True) for v in (lf, lo, pf, po)]
        score_and_select(*hs)
        res["host"] = [res[k].cpu() for k in ("dist", "cer", "sel", "top")]

    for _ in range(a.warmup):
        step_dev()
    n0 = env.lib.launch_count()
    step_dev()
    launches = env.lib.launch_count() - n0
    clk = BB.ClockSampler(env.local_rank)
    if rank == 0:
        clk.start()
    ms_step, _ = env.timed(step_dev, a.steps)
    clocks = clk.stop() if rank == 0 else None
    # the L2 flush is part of every iteration: measure it alone and take it off
    ms_flush, _ = env.timed(lambda: (flush.zero_(), None)[1], a.steps)
    step_e2e()
    ms_e2e, _ = env.timed(step_e2e, max(2, min(a.steps, 5)))
    ms_step_net, ms_e2e_net = max(ms_step - ms_flush, 1e-6), max(ms_e2e - ms_flush, 1e-6)
    lev_bytes = 4.0 * (len(lab) + len(prd)) + 2 * 4.0 * n_pairs + 4.0 * n_pairs + 8.0 * n_pairs   # symbols + offsets + distance + fp64 CER
    roofline = kernels = None
    if not a.skip_profile:
        psteps = min(a.steps, 5)
        roofline, kernels = roofline_of(env.profile(step_dev, psteps), psteps, prefer="levenshtein", bytes_override=lev_bytes)
        if roofline is not None:
            cells = float((np.diff(lab_off).astype(np.int64) * np.diff(prd_off).astype(np.int64)).sum())
            roofline["dp_cells_per_s"] = cells / (roofline["avg_launch_ms"] / 1e3)
            roofline["note"] = "bytes = SURVEY.md 8(d): code points + offsets + int32 distance + fp64 CER per pair"
    # multi-GPU: the dataset-wide top-k over the ranks' shards (values stay sharded; k candidates per rank are merged)
    merged = None
    if world > 1:
        merged = qdist.global_topk(res["cer"].float().cpu(), 1000)
    cpu = None
    if rank == 0 and world == 1 and not a.skip_cpu_baseline:
        cstep, kind, what = cer_cpu_step(lab, lab_off, prd, prd_off, n_pairs)
        cpu = cpu_time(cstep, n_pairs, what, kind=kind, threads=1, max_n=6)
    if rank != 0:
        return None
    units = n_pairs * world
    d2h = 4 * n_pairs + 8 * n_pairs + 8 * n_seg * k_seg + 8 * k_glob
    h2d = 4 * (len(lab) + len(prd)) + 8 * (n_pairs + 1)
    return line_of(env, "cer_topk", "pairs/sec, batched Levenshtein CER + TopKCER selection", "pairs/s", units, ms_step_net, ms_e2e_net, h2d,
                   d2h, launches, clocks, cer_config(world, n_pairs), "int32/f64", "weak", roofline, kernels, cpu,
                   {"ms_l2_flush_subtracted": ms_flush, "exact_pairs": int((res["dist"] == 0).sum()), "mean_cer": float(res["cer"].mean()),
                    "global_topk_merged": None if merged is None else int(merged.numel())})


def cer_cpu_step(lab, lab_off, prd, prd_off, n_pairs):
    """compare_labels over all pairs FROM THE HOST STRINGS (CSR encoding + the C restatement of the Levenshtein extension
    the reference calls once per pair), the TopKCER argsort per minibatch and the dataset-wide stable sort, on one host core."""
    from oracle import pyoracle
    s_lab, s_prd = csr_to_strings(lab, lab_off), csr_to_strings(prd, prd_off)

    def step():
        _, tot, _, cer = pyoracle.compare_labels(s_prd, s_lab, return_all=True)
        c32 = torch.from_numpy(cer.astype(np.float32)).view(-1, SEG)
        torch.argsort(c32, dim=1, descending=True, stable=True)[:, : SEG // 2]
        torch.argsort(c32.reshape(-1), descending=True, stable=True)[: n_pairs // 10]
        return tot

    return step, "port", (f"{n_pairs:,} string pairs: oracle/pyoracle.compare_labels (CSR encode + oracle/oracle.c) + torch CPU stable argsort "
                          "per minibatch and over all CERs, 1 core")


# ====================================================================================================== entry points
RUNNERS = {"jitter_step": run_jitter, "area_step": run_area, "cer_topk": run_cer}


def run(args, rank, world, local_rank, own_process_group=True):
    import qeb_b200  # noqa: F401
    env = Env(args, rank, world, local_rank, own_process_group)
    try:
        return RUNNERS[args.workload](env)
    finally:
        env.close()


def run_reference(args, rank):
    """The same workloads on the host cores: the unmodified reference modules / functions when baseline/_ref is present."""
    if rank != 0:
        return None
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    wl = args.workload
    if wl == "jitter_step":
        n_copies = 1 if args.steps + args.warmup > 6 else 2
        step, kind, what = jitter_cpu_step(n_copies)
        units, cfg, metric, unit, scaling = n_copies * BB.BATCH, jitter_config(1), "patches/sec per jittered approximation step (CRNN+CTC, 64x8 noised copies)", "patches/s", "strong"
    elif wl == "area_step":
        step, kind, what = area_cpu_step()
        units, cfg, metric, unit, scaling = BB.BATCH, area_config(1), "patches/sec per whole train_nn_area minibatch (phases A+B+C)", "patches/s", "weak"
    else:
        n = 200_000
        lab, lab_off, prd, prd_off = cer_pairs(n, 42)
        step, kind, what = cer_cpu_step(lab, lab_off, prd, prd_off, n)
        threads = 1
        units, cfg, metric, unit, scaling = n, cer_config(1, N_PAIRS), "pairs/sec, batched Levenshtein CER + TopKCER selection", "pairs/s", "weak"
    for _ in range(max(1, args.warmup)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = units * args.steps / dt
    return {"impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": scaling,
            "vs_baseline": None, "dtype": "f32" if wl != "cer_topk" else "int32/f64", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": value, "unit": unit, "cores": threads, "kind": kind, "sample": f"{args.steps} x {what}"},
            "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
