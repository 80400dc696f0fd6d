"""Microbenchmark: relative-position attention forward, tcgen05 kernel (csrc/attention_tc.cu) vs the mma.sync generation (csrc/attention.cu),
Conformer-S shape (B = 32, T' = 250, 4 heads x 36) and the Conformer-L streaming decode shape (B = 8, T' = 750, 8 heads x 64, chunk 16 /
history 64); 16 dependent calls inside one CUDA graph.  Prints JSON."""
import json
import sys

import torch

sys.path.insert(0, ".")
from tensorflowasr_b200 import kernels, ops  # noqa: E402


def bench(B, T, N, H, chunk=None, history=None):
    dev = "cuda"
    torch.manual_seed(0)
    NH, R = N * H, 2 * T - 1
    qkv = (torch.randn(B * T, 3 * NH, device=dev) * 0.7).to(torch.bfloat16)
    pos = (torch.randn(B * R, NH, device=dev) * 0.7).to(torch.bfloat16)
    ub, vb = torch.randn(N, H, device=dev) * 0.3, torch.randn(N, H, device=dev) * 0.3
    length = torch.full((B,), T, dtype=torch.int32, device=dev)
    out = {}
    ref = None
    for name, flag in (("mma_sync", False), ("tcgen05", True)):
        ops.set_attention_tc(flag)
        try:
            ctx, probs = kernels.rel_attention_fwd(qkv, pos, ub, vb, length, B, T, N, H, chunk, history)
        except Exception as e:  # the mma.sync generation has no kernel for every shape
            out[name] = {"error": str(e)[:120]}
            continue
        torch.cuda.synchronize()
        if ref is None:
            ref = ctx.float()
        else:
            out["max_abs_diff_ctx"] = (ctx.float() - ref).abs().max().item()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(16):
                kernels.rel_attention_fwd(qkv, pos, ub, vb, length, B, T, N, H, chunk, history)
        for _ in range(3):
            g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        out[name] = {"us_per_call": e0.elapsed_time(e1) * 1e3 / 160}
    return out


if __name__ == "__main__":
    print(json.dumps({"conformer_s_B32_T250_4x36": bench(32, 250, 4, 36), "conformer_l_streaming_B8_T750_8x64": bench(8, 750, 8, 64, 16, 64)}))
