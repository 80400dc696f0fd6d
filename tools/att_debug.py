"""Debug helper: run the tcgen05 attention forward on ONE shape (argv: B T N H [chunk history]) and compare with the mma.sync kernel."""
import sys

import torch

sys.path.insert(0, ".")
from tensorflowasr_b200 import kernels, ops  # noqa: E402

B, T, N, H = (int(v) for v in sys.argv[1:5])
chunk, hist = (int(sys.argv[5]), int(sys.argv[6])) if len(sys.argv) > 6 else (None, None)
torch.manual_seed(0)
NH, R = N * H, 2 * T - 1
qkv = (torch.randn(B * T, 3 * NH, device="cuda") * 0.7).to(torch.bfloat16)
pos = (torch.randn(B * R, NH, device="cuda") * 0.7).to(torch.bfloat16)
ub, vb = torch.randn(N, H, device="cuda") * 0.3, torch.randn(N, H, device="cuda") * 0.3
length = torch.full((B,), T, dtype=torch.int32, device="cuda")
ops.set_attention_tc(True)
ctx, probs = kernels.rel_attention_fwd(qkv, pos, ub, vb, length, B, T, N, H, chunk, hist)
torch.cuda.synchronize()
msg = f"tc ok {B, T, N, H, chunk, hist}"
try:
    ops.set_attention_tc(False)
    ctx2, probs2 = kernels.rel_attention_fwd(qkv, pos, ub, vb, length, B, T, N, H, chunk, hist)
    torch.cuda.synchronize()
    msg += f" max|ctx diff| {(ctx.float() - ctx2.float()).abs().max().item():.4f} max|probs diff| {(probs.float() - probs2.float()).abs().max().item():.4f}"
except Exception as e:
    msg += f" (mma.sync path unavailable: {str(e)[:60]})"
print(msg)
