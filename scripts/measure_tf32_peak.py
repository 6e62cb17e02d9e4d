#!/usr/bin/env python
"""TF32 tensor-core peak measured the way MEASURED_PEAKS.json measures bf16: torch.matmul (cuBLAS) on
8192^3 fp32 operands with TF32 allowed, best of 10 (burst) and back to back for 4 s (sustained), 2*N^3 flops.
The 3xTF32 feature GEMM executes three MMAs per useful one, so its tensor-pipe share is quoted against this."""
import json
import time

import torch


def main():
    torch.backends.cuda.matmul.allow_tf32 = True
    n = 8192
    a = torch.randn(n, n, device="cuda")
    b = torch.randn(n, n, device="cuda")
    c = torch.empty(n, n, device="cuda")
    for _ in range(3):
        torch.matmul(a, b, out=c)
    best = float("inf")
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b, out=c)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    burst = 2 * n ** 3 / (best * 1e-3) / 1e12
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0, k = time.perf_counter(), 0
    e0.record()
    while time.perf_counter() - t0 < 4.0:
        for _ in range(20):
            torch.matmul(a, b, out=c)
        k += 20
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    sustained = 2 * n ** 3 * k / (e0.elapsed_time(e1) * 1e-3) / 1e12
    print(json.dumps({"tf32_tflops": burst, "tf32_tflops_sustained": sustained, "gpu_name": torch.cuda.get_device_name(0),
                      "how": "torch.matmul fp32 8192^3 with torch.backends.cuda.matmul.allow_tf32 (cuBLAS TF32 "
                             "tensor-core path), 2*N^3 flops: best of 10 (burst) and back to back for 4 s (sustained)"}))


if __name__ == "__main__":
    main()
