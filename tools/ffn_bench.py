"""Microbenchmark: FFModule forward as one kernel (csrc/ffn_fused.cu) vs LayerNorm + two GEMM launches, Conformer-S shape (M = 8000, D = 144,
H = 576), timed as 32 dependent calls inside one CUDA graph (how the step runs them).  Prints JSON."""
import json
import sys

import torch

sys.path.insert(0, ".")
from tensorflowasr_b200 import kernels  # noqa: E402


def main():
    M, D, H = 8000, 144, 576
    dev = "cuda"
    torch.manual_seed(0)
    x = torch.randn(M, D, device=dev).to(torch.bfloat16)
    g, b = torch.ones(D, device=dev), torch.zeros(D, device=dev)
    w1t = (torch.randn(H, D, device=dev) / 12).to(torch.bfloat16)
    w2t = (torch.randn(D, H, device=dev) / 24).to(torch.bfloat16)
    b1, b2 = torch.zeros(H, device=dev), torch.zeros(D, device=dev)
    seed = torch.zeros(1, dtype=torch.int64, device=dev)

    def fused_chain(n):
        y = x
        for _ in range(n):
            y = kernels.ffn_fwd(y, g, b, w1t, b1, w2t, b2, 0.5, 0.1, seed, 1, 2)[0]
        return y

    def unfused_chain(n):
        y = x
        for _ in range(n):
            a, _, _ = kernels.layernorm_fwd(y, g, b, 1e-3)
            z = torch.empty((M, H), dtype=torch.bfloat16, device=dev)
            h = kernels.gemm(a, w1t, bias=b1, activation="swish", aux=z, dropout_p=0.1, dropout_seed=seed, dropout_salt=1)
            y = kernels.gemm(h, w2t, bias=b2, residual=y, residual_scale=0.5, dropout_p=0.1, dropout_seed=seed, dropout_salt=2)
        return y

    w2 = w2t.t().contiguous()  # W2 [H, D]
    w1 = w1t.t().contiguous()  # W1 [D, H]
    zz = torch.randn(M, H, device=dev).to(torch.bfloat16)

    def fused_bwd_chain(n):
        d = x
        for _ in range(n):
            d = kernels.ffn_bwd(d, zz, w2, w1, 0.5, 0.1, seed, 1, 2)[2]
        return d

    def unfused_bwd_chain(n):
        d = x
        for _ in range(n):
            dy = kernels.scale_mask_colsum(d, scale=0.5, dropout_p=0.1, seed=seed, salt=2)
            dz = kernels.gemm(dy, w2, mul=zz, mul_mode="swish_grad", dropout_p=0.1, dropout_seed=seed, dropout_salt=1)
            d = kernels.gemm(dz, w1)
        return d

    out = {}
    for name, fn in (("fused", fused_chain), ("unfused", unfused_chain), ("fused_bwd", fused_bwd_chain), ("unfused_bwd", unfused_bwd_chain)):
        n = 32
        fn(2)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            fn(n)
        for _ in range(3):
            graph.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            graph.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / (10 * n)
        out[name] = {"us_per_module": us, "tflops": 2 * 2 * M * D * H / us * 1e-6}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
