#!/usr/bin/env python
"""Benchmark of the hot path: one preprocessor-update training step (phase B of the reference's trainers,
train_nn_area.py:277-287 / train_nn_patch.py:312-345): UNet (BatchNorm in train mode) -> CRNN surrogate (train(),
BatchNorm frozen by set_bn_eval) -> CTC(mean) + 1.0 * MSE-to-white -> backward through both networks -> Adam step on
the UNet. Batch 64 synthetic 32x128 patches per GPU (BASELINE.json configs[1]); metric = patches/sec.

  python bench.py --gpus N --steps K --warmup W            our arm (sm_100a kernels behind the mirror modules)
  python bench.py --impl reference ...                      the reference's CPU path (torch CPU, all host threads)
For N > 1 launch with torch.distributed.run (one rank per GPU); the batch is sharded 64/GPU (weak scaling) and the
UNet gradients are all-reduced (AVG) over NCCL once per step.

Prints ONE JSON line on rank 0. `value`: inputs resident in HBM, CUDA-event timing, max over ranks; forward + losses +
backward are replayed as one CUDA graph (qeb_b200.graphs.GraphedStep), the all-reduce and Adam follow it; the eagerly
launched modules are timed too (`variants.eager_modules`). `e2e`: the same
step through the trainers' calling convention with HOST inputs (pinned image batch -> H2D, label strings encoded on
the host -> int32 CPU targets -> H2D, loss read back -> D2H) inside the timed region. `roofline`: the kernel family
with the largest share of the step, measured with CUDA events around every launch of the library in a separate
profiled pass. `cpu_baseline`: the oracle port (oracle/nn_oracle.py + torch CPU CTC/MSE/Adam) on the box's host cores.
"""
import argparse
import json
import os
import random
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CHAR_SET = ['`', ' ', '!', '"', '#', '$', '%', '&', "'", '(', ')', '*', '+', ',', '-', '.', '0', '1', '2', '3', '4',
            '5', '6', '7', '8', '9', ':', ';', '<', '=', '>', '?', '@', 'A', 'B', 'C', 'D', 'E', 'F', 'G', 'H', 'I', 'J',
            'K', 'L', 'M', 'N', 'O', 'P', 'Q', 'R', 'S', 'T', 'U', 'V', 'W', 'X', 'Y', 'Z', '[', ']', '^', 'a', 'b', 'c',
            'd', 'e', 'f', 'g', 'h', 'i', 'j', 'k', 'l', 'm', 'n', 'o', 'p', 'q', 'r', 's', 't', 'u', 'v', 'w', 'x', 'y',
            'z', '{', '|', '~', '€', '}', '\\', '/']
BATCH = 64          # patches per GPU (BASELINE.json configs[1])
H, W = 32, 128      # properties.input_size
LR_PREP = 5e-5      # patch_cli.py / area_cli.py default
SCALAR = 1.0        # weight of the MSE term (default --scalar 1)
# algorithmic FLOPs per patch of this step: UNet fprop+dgrad+wgrad 3 x 1.504 + CRNN fprop+dgrad 2 x 1.778
# (+ the CRNN weight gradients the reference also computes and discards: + 1.778)
GFLOP_PER_PATCH = 3 * 1.504 + 3 * 1.778


def synth_batch(n, seed):
    """POS-shaped synthetic data: near-white background with darker strokes; labels 1..16 chars (mean ~4.8)."""
    g = torch.Generator().manual_seed(seed)
    x = 0.9 + 0.1 * torch.rand(n, 1, H, W, generator=g)
    strokes = (torch.rand(n, 1, H // 4, W // 4, generator=g) < 0.18).float()
    strokes = torch.nn.functional.interpolate(strokes, scale_factor=4, mode="nearest")
    x = (x - 0.75 * strokes * torch.rand(n, 1, H, W, generator=g)).clamp_(0, 1)
    rng = random.Random(seed)
    labels = []
    for _ in range(n):
        ln = min(16, 1 + int(rng.expovariate(1 / 3.8)))
        labels.append("".join(rng.choice(CHAR_SET[1:]) for _ in range(ln)))
    return x, labels


def encode(labels, char_to_index):
    # TrainNNPrep._call_model: train_nn_area.py:163-171
    y = torch.tensor([char_to_index[c] for l in labels for c in l], dtype=torch.int32)
    y_size = torch.tensor([len(l) for l in labels], dtype=torch.int32)
    return y, y_size


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md clocks line). NVML is polled from
    a thread every 5 ms (nvidia-smi -lms cannot start inside a 0.1-0.5 s region); nvidia-smi is the fallback."""
    REASONS = (("hw_slowdown", "nvmlClocksEventReasonHwSlowdown"), ("hw_thermal_slowdown", "nvmlClocksEventReasonHwThermalSlowdown"),
               ("sw_thermal_slowdown", "nvmlClocksEventReasonSwThermalSlowdown"), ("sw_power_cap", "nvmlClocksEventReasonSwPowerCap"))

    def __init__(self, index):
        self.index, self.sm, self.reasons, self.max_mhz, self.power = index, [], set(), None, []
        self._stop, self._thread, self.source = threading.Event(), None, None

    def _handle(self, nv):
        try:
            uuid = str(torch.cuda.get_device_properties(self.index).uuid)
            return nv.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
        except Exception:
            return nv.nvmlDeviceGetHandleByIndex(self.index)

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = self._handle(nv)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            masks = [(n, getattr(nv, a)) for n, a in self.REASONS]

            def poll():
                while not self._stop.is_set():
                    try:
                        self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                        r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                        self.reasons.update(n for n, m in masks if r & m)
                        self.power.append(nv.nvmlDeviceGetPowerUsage(h) / 1e3)
                    except Exception:
                        pass
                    time.sleep(0.005)

            self._thread = threading.Thread(target=poll, daemon=True)
            self._thread.start()
            self.source = "nvml"
        except Exception:
            self.source = None

    def stop(self):
        if self._thread is not None:
            self._stop.set()
            self._thread.join()
        if not self.sm:  # fallback: one nvidia-smi query right after the timed region
            try:
                q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                     "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=20).stdout.strip().split(",")
                self.sm, self.max_mhz = [float(out[0])], float(out[1])
                self.reasons = {n for (n, _), v in zip(self.REASONS, out[2:6]) if v.strip() == "Active"}
                self.source = "nvidia-smi (single sample after the timed region)"
            except Exception:
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock query unavailable"], "samples": 0}
        return {"sm_mhz": statistics.median(self.sm), "sm_min_mhz": min(self.sm), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.sm), "source": self.source,
                "power_w": statistics.median(self.power) if self.power else None}


# ------------------------------------------------------------------------------------------------------ reference arm
def reference_modules():
    """(UNet, CRNN, set_bn_eval, unet_forward, crnn_forward, kind, namespace): the UNMODIFIED reference from baseline/_ref
    (the copy that travels to the GPU box) or /root/reference when one is present - `kind` "reference" - else the oracle
    port over the mirror's parameter containers - `kind` "port"."""
    from oracle import refload
    if refload.available():
        ns = refload.load()
        return ns.model_unet.UNet, ns.model_crnn.CRNN, ns.utils.set_bn_eval, (lambda m, x: m(x)), (lambda m, x: m(x)), "reference", ns
    from oracle import nn_oracle
    from qeb_b200.mirror.models.model_crnn import CRNN
    from qeb_b200.mirror.models.model_unet import UNet
    from qeb_b200.mirror.utils import set_bn_eval
    return UNet, CRNN, set_bn_eval, nn_oracle.unet_forward, nn_oracle.crnn_forward, "port", None


def cpu_step_factory(batch, seed=42):
    """The same step on the host cores: the reference's own modules (or the oracle port), torch CPU CTC / MSE / Adam."""
    UNet, CRNN, set_bn_eval, unet_f, crnn_f, kind, _ = reference_modules()
    torch.manual_seed(seed)
    prep, crnn = UNet(), CRNN(len(CHAR_SET), False)
    crnn.register_backward_hook(crnn.backward_hook)
    ctc, mse = torch.nn.CTCLoss(), torch.nn.MSELoss()
    opt = torch.optim.Adam(prep.parameters(), lr=LR_PREP, weight_decay=0)
    c2i = {c: i for i, c in enumerate(CHAR_SET)}
    x, labels = synth_batch(batch, 7)

    def step():
        prep.train(); crnn.train(); crnn.apply(set_bn_eval)
        prep.zero_grad(); crnn.zero_grad()
        img = unet_f(prep, x)
        scores = crnn_f(crnn, img)
        y, y_size = encode(labels, c2i)
        pred_size = torch.tensor([scores.shape[0]] * batch, dtype=torch.int32)
        loss = ctc(scores, y, pred_size, y_size) + SCALAR * mse(img, torch.ones(img.shape))
        loss.backward()
        opt.step()
        return float(loss.detach())

    step.kind = kind
    step.what = ("unmodified reference modules (baseline/_ref)" if kind == "reference" else "oracle/nn_oracle.py") + \
        f" on torch {torch.__version__} CPU fp32"
    return step


def run_reference(args, rank):
    if rank != 0:
        return None
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    sample = BATCH if args.steps + args.warmup <= 30 else 16
    step = cpu_step_factory(sample)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    line = {"impl": "reference", "metric": "patches/sec per train step (UNet+CRNN+CTC)", "value": value, "unit": "patches/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.gpus),
            "cpu_baseline": {"value": value, "unit": "patches/s", "cores": threads, "kind": step.kind,
                             "sample": f"{args.steps} steps of {sample} patches, {step.what}"},
            "e2e": {"value": value, "unit": "patches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    return line


def workload_config(n_gpus):
    return {"workload": "configs[1]: train_nn_area phase B step - UNet(train BN) + CRNN surrogate (BN frozen) fwd/bwd + CTC(mean) + "
                        "1.0*MSE-to-white + Adam(lr 5e-5) on the UNet; 64 synthetic 32x128 patches per GPU, V=95, T=31",
            "batch_per_gpu": BATCH, "global_batch": BATCH * n_gpus, "patch": [H, W], "parallelism": f"dp{n_gpus}",
            "operand_precision": "tensor-core operands with an 11-bit significand: fp16 copies of the activations in the forward pass, fp16 copies of the gradients scaled by a power of two per tensor (from the previous step's maximum) in the conv backward passes, tf32 reads of fp32 in the LSTM / Linear backward and in a network's first backward call; fp32 accumulation / activations / gradients / parameters (reference: fp32)",
            "l2": "no explicit flush: a step streams ~1.4 GB of activations, > 126 MB L2",
            "launch": "forward+losses+backward (+ the gradient all-reduce when N > 1) replayed as one CUDA graph (qeb_b200.graphs.GraphedStep); Adam outside it"}


def tensor_peaks():
    """Measured dense peaks per operand kind: MEASURED_PEAKS.json (driver-written: bf16 burst / sustained) and, when present,
    profiles/r2_tensor_peaks.json (scripts/measure_peaks.py on the same pool: fp16 and TF32 by the same cuBLAS method)."""
    pk = {}
    for f in (os.path.join(ROOT, "profiles", "r2_tensor_peaks.json"), os.path.join(ROOT, "MEASURED_PEAKS.json")):
        try:
            pk.update(json.load(open(f)))
        except (OSError, ValueError):
            pass
    return pk


def fold_kinds(rep):
    """Profile tags carry the operand kind of a tensor-core launch ("tc_conv_fprop.f16" / ".tf32"): returns the report with the
    kinds folded into their family, and {family: {kind: record}}."""
    fam, kinds = {}, {}
    for k, v in rep.items():
        base, _, kind = k.partition(".")
        a = fam.setdefault(base, {"launches": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
        for f in a:
            a[f] += v[f]
        if kind:
            kinds.setdefault(base, {})[kind] = v
    return fam, kinds


def by_operand_kind(kinds, peaks, psteps=1):
    """Each operand kind of a family against ITS measured dense peak (TF32: profiles/r2_tensor_peaks.json, else half the bf16
    figure). kinds: raw profile records over `psteps` profiled steps."""
    out = {}
    for kind, v in kinds.items():
        if kind == "tf32":
            peak, src = peaks.get("tf32_tflops"), "measured TF32 cuBLAS burst (profiles/r2_tensor_peaks.json)"
            if not peak:
                peak, src = peaks.get("bf16_tflops", 1590.0) / 2, "half the measured bf16 burst (no TF32 measurement on file)"
        else:
            peak, src = peaks.get("fp16_tflops"), "measured fp16 cuBLAS burst (profiles/r2_tensor_peaks.json)"
            if not peak:
                peak, src = peaks.get("bf16_tflops", 1590.0), "measured bf16 burst (MEASURED_PEAKS.json)"
        ach = v["flops"] / v["ms"] / 1e9 if v["ms"] > 0 else 0.0
        out[kind] = {"achieved": ach, "peak": peak, "frac": ach / peak, "unit": "TFLOP/s", "launches_per_step": v["launches"] / psteps,
                     "ms_per_step": v["ms"] / psteps, "peak_source": src}
    return out


def parity_block():
    """Measured parity of this build's network gradients (scripts/grad_parity.py on the B200, committed under profiles/)."""
    try:
        p = json.load(open(os.path.join(ROOT, "profiles", "r2_grad_parity.json")))
        b, a = p["phase_b"], p["phase_a"]
        zero = ("crnn.convo.conv5.bias", "crnn.convo.conv6.bias")   # analytically zero gradients (bias before a train-mode BN)
        wa = max((v[0], k) for k, v in a["grads"].items() if k not in zero)
        return {"source": "profiles/r2_grad_parity.json (scripts/grad_parity.py --competitors, B=16, vs the fp32 CPU oracle)",
                "phase_b_loss_rel": b["loss_rel"], "phase_b_worst_unet_weight_grad_rel_l2": b["worst_rel_l2"],
                "phase_b_worst_tensor": b["worst_tensor"], "phase_b_min_cos": b["min_cos"],
                "phase_a_worst_crnn_weight_grad_rel_l2": wa[0], "phase_a_worst_tensor": wa[1],
                "same_metric_other_implementations_worst": b.get("competitors_worst"),
                "note": "forward operands carry 11-bit significands (tcgen05 has no fp32 operand kind): ReLU / BatchNorm "
                        "decisions of near-zero pre-activations flip, which bounds network-gradient parity on random-init nets; "
                        "cuDNN TF32 shows the same, see DESIGN.md section 5"}
    except (OSError, ValueError, KeyError):
        return None


def decode_gate_block():
    """Observed counts of the north_star decode gate (tests/test_trained_decode_gpu.py on the B200, committed under profiles/)."""
    try:
        g = json.load(open(os.path.join(ROOT, "profiles", "r2_decode_gate.json")))
        return {"source": "profiles/r2_decode_gate.json (tests/test_trained_decode_gpu.py)", **g}
    except (OSError, ValueError):
        return None


# ------------------------------------------------------------------------------------------------------ our arm
def main():
    # stdout carries exactly ONE line (the JSON result): everything libraries print while the run is in flight (NCCL's
    # version banner, warnings) is sent to stderr by pointing fd 1 at fd 2 until the result is ready
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        line = run()
    finally:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)
    if line is not None:
        print(json.dumps(line), flush=True)


WORKLOADS = ("prep_step", "jitter_step", "area_step", "cer_topk")


def parse_args(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="qeb", choices=["qeb", "reference"])
    ap.add_argument("--workload", default="prep_step", choices=WORKLOADS,
                    help="prep_step = BASELINE.json configs[1] (the headline, default); jitter_step = configs[2]; "
                         "area_step = configs[3]; cer_topk = configs[4] (bench_workloads.py)")
    ap.add_argument("--no-extras", action="store_true", help="default workload only: do not append the short runs of the other "
                                                             "three workloads under `workloads`")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-serial", action="store_true", help="e2e leg without input staging: H2D copies, label encoding and the replay in line on one stream")
    ap.add_argument("--skip-profile", action="store_true")
    ap.add_argument("--skip-eager", action="store_true", help="profiling runs: only the graph arm (variants are null)")
    args = ap.parse_args(argv)
    args.warmup = max(args.warmup, 3) if args.impl == "qeb" else max(args.warmup, 1)
    return args


def run():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        if args.workload != "prep_step":
            import bench_workloads
            return bench_workloads.run_reference(args, rank)
        return run_reference(args, rank)
    if args.workload != "prep_step":
        import bench_workloads
        return bench_workloads.run(args, rank, world, local_rank)

    import torch.distributed as dist
    import qeb_b200
    from qeb_b200 import _lib
    from qeb_b200.mirror import ctc as qctc
    from qeb_b200.mirror import dist as qdist
    from qeb_b200.mirror import train_ops
    from qeb_b200.mirror.models.model_crnn import CRNN
    from qeb_b200.mirror.models.model_unet import UNet
    from qeb_b200.mirror.utils import set_bn_eval

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the qeb hot path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)
    assert _lib.load().qeb_check_device() == 0, _lib.load().qeb_last_error()

    torch.manual_seed(42)
    prep, crnn = UNet().to(dev), CRNN(len(CHAR_SET), False).to(dev)
    crnn.register_backward_hook(crnn.backward_hook)          # train_nn_area.py:92
    ctc_loss = qctc.CTCLoss()                                 # train_nn_area.py:146
    opt = train_ops.Adam(prep.parameters(), lr=LR_PREP, weight_decay=0)  # train_nn_area.py:152-154
    c2i = {c: i for i, c in enumerate(CHAR_SET)}
    x_host, labels = synth_batch(BATCH, 7 + rank)
    x_pin = x_host.pin_memory()
    x_dev = x_host.to(dev)
    y, y_size = encode(labels, c2i)
    pred_size = torch.tensor([W // 4 - 1] * BATCH, dtype=torch.int32)
    packed = qctc.pack_targets(y, pred_size, y_size, dev)
    unet_params = [p for p in prep.parameters()]

    # data parallel: the 31 MB flat UNet gradient is averaged over the ranks once per step; the part of it the backward pass
    # finishes first (decoder, then bottleneck + encoder 4: 96 %) is reduced on a communication stream meanwhile
    overlap = qdist.BucketedAllReduce(prep, average=True)

    def allreduce_grads():
        overlap()

    def step_device():
        prep.train(); crnn.train(); crnn.apply(set_bn_eval)          # train_nn_area.py:277-279
        prep.zero_grad(set_to_none=True); crnn.zero_grad(set_to_none=True)
        img = prep(x_dev)                                             # :283
        scores = crnn(img)                                            # :284 (_call_model)
        loss = ctc_loss(scores, packed) + SCALAR * train_ops.mse_to_ones(img)   # :285 (_get_loss)
        loss.backward()                                               # :286
        allreduce_grads()
        opt.step()                                                    # :287
        return loss

    def step_e2e():
        prep.train(); crnn.train(); crnn.apply(set_bn_eval)
        prep.zero_grad(set_to_none=True); crnn.zero_grad(set_to_none=True)
        xv = x_pin.to(dev, non_blocking=True)                         # X_var = images.to(self.device)  :217
        img = prep(xv)
        scores = crnn(img)
        yy, yy_size = encode(labels, c2i)                             # host-side label encoding, as _call_model
        ps = torch.tensor([scores.shape[0]] * BATCH, dtype=torch.int32)
        loss = ctc_loss(scores, yy, ps, yy_size) + SCALAR * train_ops.mse_to_ones(img)
        loss.backward()
        allreduce_grads()
        opt.step()
        return loss.item()                                            # D2H read of the step's loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        # hand back a number, not the loss tensor: a live loss keeps its autograd graph - and the parameters' AccumulateGrad
        # nodes, which remember the (legacy default) stream of the eager steps - alive, and a later graph capture fails with
        # cudaErrorStreamCaptureImplicit when the engine syncs that stream
        return float(ms) / steps, (float(out.detach()) if torch.is_tensor(out) else out)

    # ---- eager arm first (it re-allocates .grad every step; the graph below pins them, so eager runs must precede it)
    ms_eager = ms_e2e_eager = ms_frozen = None
    if not args.skip_eager:
        for _ in range(args.warmup):
            step_device()
        ms_eager, _ = timed(step_device, args.steps)
        for _ in range(2):
            step_e2e()
        ms_e2e_eager, _ = timed(step_e2e, args.steps)
        # variant (reported beside the headline, not instead of it): the surrogate's parameter gradients are computed and
        # thrown away by the reference in this phase (train_nn_area.py:280,286 - only optimizer_prep steps); freezing them
        # with requires_grad_(False) is a one-line change on the caller's side that skips those kernels
        for p_ in crnn.parameters():
            p_.requires_grad_(False)
        for _ in range(2):
            step_device()
        ms_frozen, _ = timed(step_device, args.steps)
        for p_ in crnn.parameters():
            p_.requires_grad_(True)

    roofline, kernels = None, None
    if not args.skip_profile and not args.skip_eager:
        _lib.prof_enable(True)
        torch.cuda.synchronize()
        _lib.prof_report()  # drop anything recorded so far
        psteps = min(args.steps, 5)
        for _ in range(psteps):
            step_device()
        torch.cuda.synchronize()
        rep, kinds = fold_kinds(_lib.prof_report())
        _lib.prof_enable(False)
        total = sum(v["ms"] for v in rep.values())
        kernels = {k: {"launches_per_step": v["launches"] / psteps, "ms_per_step": v["ms"] / psteps, "share": v["ms"] / total,
                       "tflops": v["flops"] / v["ms"] / 1e9 if v["ms"] > 0 else 0.0,
                       "gbs": v["bytes"] / v["ms"] / 1e6 if v["ms"] > 0 else 0.0} for k, v in rep.items()}
        top = max(rep, key=lambda k: rep[k]["ms"])
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json"))).get(top)
        except (OSError, ValueError):
            pass
        v = rep[top]
        if top.startswith("tc_"):
            peak = peaks.get("bf16_tflops", 1590.0)
            ach = v["flops"] / v["ms"] / 1e9
            roofline = {"kernel": top, "bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                        "traffic": traffic, "peak_source": ("measured bf16 dense burst (MEASURED_PEAKS.json); the family is kind::f16 in the forward "
                                                            "pass and in the conv backward passes, kind::tf32 (nominally half the rate) in the LSTM / Linear backward"
                                                            if peaks else "fallback 1590 bf16")}
        else:
            peak = peaks.get("hbm_gbs", 6650.0)
            ach = v["bytes"] / v["ms"] / 1e6
            roofline = {"kernel": top, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                        "traffic": traffic, "peak_source": "measured copy bandwidth (MEASURED_PEAKS.json)" if peaks else "fallback 6650"}
        if top in kinds:
            roofline["by_operand_kind"] = by_operand_kind(kinds[top], tensor_peaks(), psteps)
        roofline["share_of_step"] = v["ms"] / total
        roofline["launches_per_step"] = v["launches"] / psteps
        roofline["avg_launch_ms"] = v["ms"] / v["launches"]

    # ---- headline arm: the same step captured once as a CUDA graph (qeb_b200.graphs) and replayed; the gradient
    # all-reduce and the Adam launch stay outside the graph
    from qeb_b200.graphs import GraphedStep, StaticTargets
    tg_static = StaticTargets(BATCH, W // 4 - 1, dev).load(y, pred_size, y_size)
    x_static = x_dev.clone()
    prep.train(); crnn.train(); crnn.apply(set_bn_eval)

    def fwd_bwd_local():
        img = prep(x_static)
        scores = crnn(img)
        loss = ctc_loss(scores, tg_static) + SCALAR * train_ops.mse_to_ones(img)
        loss.backward()
        return loss

    def fwd_bwd():
        loss = fwd_bwd_local()
        allreduce_grads()       # captured with the step: event wait, communication-stream fork / join and both NCCL calls
        return loss

    ms_no_allreduce = None
    if world > 1:   # the same step without the exchange, for the exposed all-reduce time (captured and timed FIRST: the
        g0 = GraphedStep(fwd_bwd_local, modules=[prep, crnn], warmup=3)   # headline graph below re-pins the gradients)

        def step_noar():
            loss = g0()
            opt.step()
            return loss

        for _ in range(args.warmup):
            step_noar()
        ms_no_allreduce, _ = timed(step_noar, args.steps)
        g0.close()

    gstep = GraphedStep(fwd_bwd, modules=[prep, crnn], warmup=3)

    def step_graph():
        loss = gstep()
        opt.step()
        return loss

    # End to end through the public API of the graphed step (qeb_b200.graphs): every step copies ITS batch from pinned host memory
    # and encodes ITS labels on the host (as _call_model does), and reads its loss back. BatchStager issues the host -> device
    # copies of the next step on a copy stream and the host encodes the next labels while the current replay runs (one batch ahead,
    # as a DataLoader is); --e2e-serial keeps everything in line on one stream (the round-1 arrangement).
    from qeb_b200.graphs import BatchStager
    stager = BatchStager(x_static, tg_static)

    def step_graph_e2e_serial():
        x_static.copy_(x_pin, non_blocking=True)                     # host batch -> HBM
        yy, yy_size = encode(labels, c2i)                             # host-side label encoding, as _call_model
        tg_static.load(yy, pred_size, yy_size)                        # one pinned staging buffer, one H2D copy
        loss = gstep()
        opt.step()
        return loss.item()                                            # D2H read of the step's loss

    def step_graph_e2e_staged():
        stager.commit()                                               # this step's batch + labels -> the static buffers (D2D)
        loss = gstep()
        opt.step()
        yy, yy_size = encode(labels, c2i)                             # the NEXT step's labels, encoded behind the replay
        stager.stage(x_pin, yy, pred_size, yy_size)                   # ... and its H2D copies on the copy stream
        return loss.item()                                            # D2H read of THIS step's loss

    step_graph_e2e = step_graph_e2e_serial if args.e2e_serial else step_graph_e2e_staged
    if not args.e2e_serial:
        yy0, yy0_size = encode(labels, c2i)
        stager.stage(x_pin, yy0, pred_size, yy0_size)

    for _ in range(args.warmup):
        step_graph()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_step, last_loss = timed(step_graph, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    launches = (gstep.launches + 1) * args.steps                      # captured library kernels per replay + Adam
    for _ in range(2):
        step_graph_e2e()
    ms_e2e, _ = timed(step_graph_e2e, args.steps)

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.skip_cpu_baseline:
        threads = os.cpu_count() or 1
        torch.set_num_threads(threads)
        cstep = cpu_step_factory(BATCH)
        cstep()
        t0, n = time.perf_counter(), 0
        while n < 3 or (time.perf_counter() - t0 < 12 and n < 12):
            cstep(); n += 1
        dt = time.perf_counter() - t0
        cpu_baseline = {"value": BATCH * n / dt, "unit": "patches/s", "cores": threads, "kind": cstep.kind,
                        "sample": f"{n} steps of {BATCH} patches ({dt:.1f} s), {cstep.what}"}

    line = None
    if rank == 0:
        value = BATCH * world / (ms_step / 1e3)
        e2e = BATCH * world / (ms_e2e / 1e3)
        h2d = x_pin.numel() * 4 + tg_static._host.numel() * 4   # image batch + the staged CTC targets / lengths
        line = {"metric": "patches/sec per train step (UNet+CRNN+CTC)", "value": value, "unit": "patches/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "fp16/tf32", "data": "synthetic", "config": workload_config(world),
                "e2e": {"value": e2e, "unit": "patches/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                        "staging": ("serial: H2D copies, host label encoding and the replay in line on one stream" if args.e2e_serial else
                                    "qeb_b200.graphs.BatchStager: every step copies its own batch and encodes its own labels, one batch "
                                    "ahead on a copy stream (behind the previous replay), and reads its own loss back")},
                "gpu_launches": launches, "gpu_launches_per_step": launches / args.steps, "clocks": clocks,
                "allreduce": None if world == 1 else {
                    "bytes": sum(p.numel() for p in unet_params) * 4, "op": "AVG", "collectives_per_step": 3,
                    "ms_per_step_without_exchange": ms_no_allreduce, "exposed_us": 1e3 * (ms_step - ms_no_allreduce),
                    "note": "decoder gradients (12.2 MB) and bottleneck + encoder-4 gradients (17.7 MB) all-reduced on a communication stream as soon as "
                            "they are final, while the rest of the backward runs (qeb_unet_backward_bucketed); the last 1.1 MB after it; all "
                            "inside the captured graph"},
                "tflops_algorithmic": GFLOP_PER_PATCH * value / 1e3, "loss": last_loss,
                "variants": None if args.skip_eager else {
                    "eager_modules": {"value": BATCH * world / (ms_eager / 1e3), "ms_per_step": ms_eager,
                                      "e2e_value": BATCH * world / (ms_e2e_eager / 1e3), "e2e_ms_per_step": ms_e2e_eager,
                                      "note": "the mirror modules called eagerly (no CUDA graph), as an unmodified trainer does"},
                    "surrogate_requires_grad_false": {"value": BATCH * world / (ms_frozen / 1e3), "ms_per_step": ms_frozen,
                                                      "note": "eager; skips the CRNN weight gradients the reference computes and discards"}},
                "roofline": roofline, "kernels": kernels, "cpu_baseline": cpu_baseline, "parity": parity_block(),
                "decode_gate": decode_gate_block()}
    # ---- the other BASELINE.json configs, short runs, appended so that they are part of the driver-run line
    # (single-GPU line only: at N > 1 an exception on one rank inside an extra would leave the others waiting in a collective;
    # the multi-GPU forms of the workloads are run explicitly with --workload)
    if not args.no_extras and world == 1:
        import bench_workloads
        extras = {}
        for wl in ("jitter_step", "area_step", "cer_topk"):
            sub = argparse.Namespace(**vars(args))
            sub.workload, sub.steps, sub.warmup = wl, max(3, min(args.steps, 10)), 3
            sub.skip_cpu_baseline = True if world > 1 else args.skip_cpu_baseline
            try:
                extras[wl] = bench_workloads.run(sub, rank, world, local_rank, own_process_group=False)
            except Exception as e:   # an extra must never take the headline down with it
                extras[wl] = {"error": f"{type(e).__name__}: {e}"}
        if line is not None:
            line["workloads"] = extras
    if world > 1:
        gstep.close()            # the captured graph holds NCCL collectives: it has to go before the communicator
        dist.destroy_process_group()
    return line if rank == 0 else None


if __name__ == "__main__":
    main()
