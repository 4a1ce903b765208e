#!/usr/bin/env python
"""Dense tensor-core peaks of this B200 by operand kind, measured the way MEASURED_PEAKS.json's bf16 figure is (cuBLAS GEMM,
CUDA events, burst = best of short runs, sustained = a 2 s loop): fp16, bf16 and TF32 (fp32 inputs, allow_tf32). The backward
contractions of the step are kind::tf32, the forward ones kind::f16: bench.py rates each against its own peak
(roofline.by_operand_kind). `python scripts/measure_peaks.py > profiles/r2_tensor_peaks.json`."""
import json
import time

import torch


def gemm_tflops(dtype, tf32, n=8192, burst_iters=20, sustain_s=2.0):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    a = torch.randn(n, n, device="cuda", dtype=dtype)
    b = torch.randn(n, n, device="cuda", dtype=dtype)
    c = torch.empty(n, n, device="cuda", dtype=dtype)
    for _ in range(5):
        torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    best = 0.0
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(burst_iters):
            torch.matmul(a, b, out=c)
        e1.record()
        torch.cuda.synchronize()
        best = max(best, 2.0 * n ** 3 * burst_iters / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    t0, it = time.perf_counter(), 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    while time.perf_counter() - t0 < sustain_s:
        for _ in range(10):
            torch.matmul(a, b, out=c)
        it += 10
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    return best, 2.0 * n ** 3 * it / (e0.elapsed_time(e1) * 1e-3) / 1e12


if __name__ == "__main__":
    out = {"device": torch.cuda.get_device_name(0), "method": "torch.matmul (cuBLAS) 8192^3, CUDA events; burst = best of 5 x 20 launches, sustained = 2 s loop"}
    for name, dt, tf32 in (("bf16", torch.bfloat16, False), ("fp16", torch.float16, False), ("tf32", torch.float32, True), ("fp32", torch.float32, False)):
        b, s = gemm_tflops(dt, tf32)
        out[name + "_tflops"] = b
        out[name + "_tflops_sustained"] = s
    torch.backends.cuda.matmul.allow_tf32 = False
    print(json.dumps(out, indent=1))
