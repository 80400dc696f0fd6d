"""Per-launch GPU time of the small kernels of the step, measured back to back inside one CUDA graph (no host launch overhead):
the tcgen05 GEMM on the encoder's shapes (with the epilogues the step uses), cuBLAS on the same shapes, LN / row kernels."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tensorflowasr_b200 import kernels  # noqa: E402

REPS = 40


def graph_time_us(fn):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(REPS):
                fn()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (5 * REPS) * 1e3


def main():
    only = sys.argv[1] if len(sys.argv) > 1 else None
    res = {}
    dev = "cuda"
    M = 8000

    def rnd(*shape):
        return torch.randn(*shape, device=dev).to(torch.bfloat16)

    cases = []
    for name, (m, n, k) in {"ffn1": (M, 576, 144), "ffn2": (M, 144, 576), "qkv": (M, 432, 144), "out": (M, 144, 144), "glu": (M, 288, 144),
                            "pos": (15968, 144, 144), "sub2": (32 * 250 * 20, 144, 1296), "lin": (M, 144, 2880)}.items():
        a, b, bias = rnd(m, k), rnd(n, k), torch.randn(n, device=dev)
        out = torch.empty(m, n, dtype=torch.bfloat16, device=dev)
        aux = torch.empty(m, n, dtype=torch.bfloat16, device=dev)
        res_in = rnd(m, n)
        seed = torch.zeros(1, dtype=torch.int64, device=dev)
        cases.append((f"gemm_{name}_plain_{m}x{n}x{k}", lambda a=a, b=b, out=out: kernels.gemm(a, b, out=out)))
        cases.append((f"gemm_{name}_bias", lambda a=a, b=b, out=out, bias=bias: kernels.gemm(a, b, out=out, bias=bias)))
        if name == "ffn1":
            cases.append((f"gemm_{name}_bias_swish", lambda a=a, b=b, out=out, bias=bias: kernels.gemm(a, b, out=out, bias=bias, activation="swish")))
            cases.append((f"gemm_{name}_bias_aux", lambda a=a, b=b, out=out, bias=bias, aux=aux: kernels.gemm(a, b, out=out, bias=bias, aux=aux)))
            cases.append((f"gemm_{name}_bias_swish_aux", lambda a=a, b=b, out=out, bias=bias, aux=aux: kernels.gemm(a, b, out=out, bias=bias, activation="swish", aux=aux)))
        if name in ("ffn2", "out"):
            cases.append((f"gemm_{name}_bias_res", lambda a=a, b=b, out=out, bias=bias, r=res_in: kernels.gemm(a, b, out=out, bias=bias, residual=r, residual_scale=0.5)))
            cases.append((f"gemm_{name}_bias_drop", lambda a=a, b=b, out=out, bias=bias, seed=seed: kernels.gemm(
                a, b, out=out, bias=bias, dropout_p=0.1, dropout_seed=seed, dropout_salt=3)))
            cases.append((f"gemm_{name}_bias_drop_res", lambda a=a, b=b, out=out, bias=bias, r=res_in, seed=seed: kernels.gemm(
                a, b, out=out, bias=bias, residual=r, residual_scale=0.5, dropout_p=0.1, dropout_seed=seed, dropout_salt=3)))
        cases.append((f"cublas_{name}", lambda a=a, b=b: torch.matmul(a, b.t())))
    # weight gradients (MN-major, split-K, fp32 atomics)
    for name, (m, n, k) in {"wg_ffn1": (144, 576, M), "wg_ffn2": (576, 144, M), "wg_qkv": (144, 432, M), "wg_out": (144, 144, M)}.items():
        x, dy = rnd(k, m), rnd(k, n)
        acc = torch.zeros(m, n, device=dev)
        for ks in (4, 8, 16):
            cases.append((f"{name}_ks{ks}", lambda x=x, dy=dy, acc=acc, ks=ks: kernels.gemm(x, dy, mn_major=True, out=acc, accumulate=True, k_splits=ks)))
        cases.append((f"cublas_{name}", lambda x=x, dy=dy: torch.matmul(x.t(), dy)))
    # row kernels
    x = rnd(M, 144)
    g_, b_ = torch.ones(144, device=dev), torch.zeros(144, device=dev)
    y, mean, rstd = kernels.layernorm_fwd(x, g_, b_)
    dg, db = torch.zeros(144, device=dev), torch.zeros(144, device=dev)
    cases.append(("layernorm_fwd_8000x144", lambda: kernels.layernorm_fwd(x, g_, b_)))
    cases.append(("layernorm_bwd_8000x144", lambda: kernels.layernorm_bwd(x, x, mean, rstd, g_, dg, db)))
    seed = torch.zeros(1, dtype=torch.int64, device=dev)
    cs = torch.zeros(144, device=dev)
    cases.append(("scale_mask_colsum_8000x144", lambda: kernels.scale_mask_colsum(x, scale=0.5, dropout_p=0.1, seed=seed, salt=1, colsum=cs)))
    cases.append(("torch_add_8000x144 (elementwise yardstick)", lambda: torch.add(x, x)))
    z = torch.empty(1, device=dev)
    cases.append(("torch_fill_1 (empty-kernel yardstick)", lambda: z.fill_(1.0)))
    for name, fn in cases:
        if only and only not in name:
            continue
        try:
            res[name] = round(graph_time_us(fn), 2)
        except Exception as e:  # noqa: BLE001
            res[name] = f"error: {e}"
        print(name, res[name], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "graph_bench.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
