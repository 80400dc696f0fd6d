"""Phase timeline of the fused FFN kernel (csrc/ffn_fused.cu), CTA 0, from globaltimer stamps: where the ~25 us of one tile go."""
import ctypes
import sys

import torch

sys.path.insert(0, ".")
from tensorflowasr_b200 import _lib, kernels  # noqa: E402

lib = _lib.load()
lib.tfasr_ffn_set_trace.argtypes = [ctypes.c_void_p]
lib.tfasr_ffn_set_trace.restype = None
M, D, H = 8000, 144, 576
dev = "cuda"
torch.manual_seed(0)
x = torch.randn(M, D, device=dev).to(torch.bfloat16)
g, b = torch.ones(D, device=dev), torch.zeros(D, device=dev)
w1t = (torch.randn(H, D, device=dev) / 12).to(torch.bfloat16)
w2t = (torch.randn(D, H, device=dev) / 24).to(torch.bfloat16)
b1, b2 = torch.zeros(H, device=dev), torch.zeros(D, device=dev)
seed = torch.zeros(1, dtype=torch.int64, device=dev)
zz = torch.randn(M, H, device=dev).to(torch.bfloat16)
for mode in ("fwd", "bwd"):
    tr = torch.zeros(128, dtype=torch.int64, device=dev)
    for it in range(3):
        if it == 2:
            lib.tfasr_ffn_set_trace(tr.data_ptr())
        if mode == "fwd":
            kernels.ffn_fwd(x, g, b, w1t, b1, w2t, b2, 0.5, 0.1, seed, 1, 2)
        else:
            kernels.ffn_bwd(x, zz, w2t.t().contiguous(), w1t.t().contiguous(), 0.5, 0.1, seed, 1, 2)
        torch.cuda.synchronize()
    lib.tfasr_ffn_set_trace(None)
    t = tr.cpu().tolist()
    t0 = t[0]
    rel = lambda i: (t[i] - t0) / 1e3 if t[i] else float("nan")  # noqa: E731
    print(f"== {mode}: setup->0; prologue done {rel(1):.2f} us; last G2 issued {rel(2):.2f}; acc2 ready {rel(3):.2f}; out stored {rel(4):.2f}")
    print(" c | W landed | G1 issue | acc1 ready(epi) | in regs | computed | buf free | written | published(w0) | h published(MMA)")
    for c in range(9):
        print(f" {c} | {rel(16 + c):8.2f} | {rel(28 + c):8.2f} | {rel(52 + c):8.2f} | {rel(88 + c):8.2f} | {rel(64 + c):8.2f} | {rel(76 + c):8.2f} | "
              f"{rel(100 + c):8.2f} | {rel(112 + c):8.2f} | {rel(40 + c):8.2f}")
