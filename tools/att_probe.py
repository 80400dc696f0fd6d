"""Relative-attention kernels on the Conformer-S shape (B=32, T'=250, 4 heads x 36): target for ncu captures
(`ncu --set full -k regex:rel_attention --launch-skip 3 -c 3`) and CUDA-event timing."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tensorflowasr_b200 import kernels  # noqa: E402

B, T, N, H = 32, 250, 4, 36
NH, R = N * H, 2 * T - 1
qkv = (torch.randn(B * T, 3 * NH, device="cuda") * 0.7).to(torch.bfloat16)
pos = (torch.randn(B * R, NH, device="cuda") * 0.7).to(torch.bfloat16)
u = torch.randn(N, H, device="cuda") * 0.2
v = torch.randn(N, H, device="cuda") * 0.2
length = torch.full((B,), T, device="cuda", dtype=torch.int32)
dctx = (torch.randn(B * T, NH, device="cuda") * 0.5).to(torch.bfloat16)
du = torch.zeros(N, H, device="cuda")
dv = torch.zeros(N, H, device="cuda")
for _ in range(2):
    ctx, probs = kernels.rel_attention_fwd(qkv, pos, u, v, length, B, T, N, H)
    kernels.rel_attention_bwd(qkv, pos, u, v, length, B, T, N, H, probs, dctx, du, dv, ctx=ctx)
torch.cuda.synchronize()
