"""Measure the FP64 roofline denominators on the box: cuBLAS DGEMM / ZGEMM throughput via torch (library
calls, used ONLY as the measured peak) and HBM copy bandwidth.  Writes gpurun_out/fp64_peak.json."""
import json
import os
import sys
import time

import torch


def bench(fn, n_iter=10):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(n_iter):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    return best


def main():
    out = {"gpu": torch.cuda.get_device_name(0)}
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    t = bench(lambda: torch.matmul(a, b), 5)
    out["dgemm_8192_tflops"] = 2 * n ** 3 / t / 1e12
    t0 = time.time()
    k = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    while time.time() - t0 < 3.0:
        torch.matmul(a, b)
        k += 1
        if k % 4 == 0:
            torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    out["dgemm_8192_tflops_sustained"] = 2 * n ** 3 * k / (e0.elapsed_time(e1) * 1e-3) / 1e12
    n = 4096
    a = torch.randn(n, n, dtype=torch.complex128, device="cuda")
    b = torch.randn(n, n, dtype=torch.complex128, device="cuda")
    t = bench(lambda: torch.matmul(a, b), 5)
    out["zgemm_4096_tflops"] = 8 * n ** 3 / t / 1e12
    for n in (512, 1024):
        a = torch.randn(64, n, n, dtype=torch.complex128, device="cuda")
        b = torch.randn(64, n, n, dtype=torch.complex128, device="cuda")
        t = bench(lambda: torch.matmul(a, b), 5)
        out[f"zgemm_batched64_{n}_tflops"] = 64 * 8 * n ** 3 / t / 1e12
    # library SVD for context (NOT used by the product): cuSOLVER batched complex128 SVD 512x512 x 6
    try:
        m = torch.randn(6, 512, 512, dtype=torch.complex128, device="cuda")
        t = bench(lambda: torch.linalg.svd(m, full_matrices=False), 2)
        out["cusolver_svd_6x512_ms"] = t * 1e3
    except Exception as ex:  # noqa
        out["cusolver_svd_6x512_ms"] = str(ex)[:100]
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/fp64_peak.json", "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
