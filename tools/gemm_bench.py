"""CUDA-event timings of the tcgen05 GEMM on the shapes the Conformer-S step uses (vs torch.matmul / cuBLAS as a yardstick)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tensorflowasr_b200 import kernels  # noqa: E402


def time_us(fn, iters=20, warmup=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def main():
    res = {}
    only = sys.argv[1] if len(sys.argv) > 1 else None
    shapes = [(8000, 576, 144), (8000, 144, 576), (8000, 144, 144), (8000, 432, 144), (8000, 144, 2880), (520000, 1000, 320), (520000, 320, 1000)]
    for M, N, K in shapes:
        a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
        b = torch.randn(N, K, device="cuda").to(torch.bfloat16)
        bias = torch.randn(N, device="cuda")
        out = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
        key = f"{M}x{N}x{K}"
        if only and only != key:
            continue
        t_plain = time_us(lambda: kernels.gemm(a, b, out=out))
        t_bias_swish = time_us(lambda: kernels.gemm(a, b, out=out, bias=bias, activation="swish"))
        t_torch = time_us(lambda: torch.matmul(a, b.t()))
        fl = 2.0 * M * N * K
        res[key] = dict(ours_us=t_plain, ours_bias_swish_us=t_bias_swish, cublas_us=t_torch, ours_tflops=fl / t_plain / 1e6, cublas_tflops=fl / t_torch / 1e6)
        print(key, res[key], flush=True)
    # weight-gradient shape (MN-major, split-K)
    for M, N, K in [(144, 576, 8000), (576, 144, 8000), (320, 1000, 520000)]:
        x = torch.randn(K, M, device="cuda").to(torch.bfloat16)
        dy = torch.randn(K, N, device="cuda").to(torch.bfloat16)
        acc = torch.zeros(M, N, device="cuda")
        key = f"wgrad_{M}x{N}x{K}"
        if only and only != key:
            continue
        for ks in (1, 8, 16, 32, 48):
            t = time_us(lambda: kernels.gemm(x, dy, mn_major=True, out=acc, accumulate=True, k_splits=ks))
            res[f"{key}_ks{ks}"] = dict(us=t, tflops=2.0 * M * N * K / t / 1e6)
            print(key, ks, res[f"{key}_ks{ks}"], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "gemm_bench.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
