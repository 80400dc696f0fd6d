"""Microbenchmark: relative-position attention forward and backward (csrc/attention.cu, the mma.sync generation the train step uses at
T' <= 256) at the Conformer-S shape (B = 32, T' = 250, 4 heads x 36): 16 dependent calls per CUDA graph, device time per call.  Prints JSON.
    python tools/att_fwd_bwd_bench.py [B T N H]"""
import json
import sys

import torch

sys.path.insert(0, ".")
from tensorflowasr_b200 import kernels, ops  # noqa: E402


def timed_graph(fn, calls=16, reps=10):
    g = torch.cuda.CUDAGraph()
    fn()
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        for _ in range(calls):
            fn()
    for _ in range(3):
        g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (calls * reps)


def main():
    B, T, N, H = (int(a) for a in sys.argv[1:5]) if len(sys.argv) >= 5 else (32, 250, 4, 36)
    dev = "cuda"
    torch.manual_seed(0)
    NH, R = N * H, 2 * T - 1
    ops.set_attention_tc(False)
    qkv = (torch.randn(B * T, 3 * NH, device=dev) * 0.7).to(torch.bfloat16)
    pos = (torch.randn(B * R, NH, device=dev) * 0.7).to(torch.bfloat16)
    ub, vb = torch.randn(N, H, device=dev) * 0.3, torch.randn(N, H, device=dev) * 0.3
    length = torch.full((B,), T, dtype=torch.int32, device=dev)
    ctx, probs = kernels.rel_attention_fwd(qkv, pos, ub, vb, length, B, T, N, H)
    dctx = (torch.randn(B * T, NH, device=dev) * 0.1).to(torch.bfloat16)
    du, dvb = torch.zeros(N, H, device=dev), torch.zeros(N, H, device=dev)
    out = {"shape": [B, T, N, H]}
    out["fwd_us"] = timed_graph(lambda: kernels.rel_attention_fwd(qkv, pos, ub, vb, length, B, T, N, H))
    out["bwd_us"] = timed_graph(lambda: kernels.rel_attention_bwd(qkv, pos, ub, vb, length, B, T, N, H, probs, dctx, du, dvb, ctx=ctx))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
