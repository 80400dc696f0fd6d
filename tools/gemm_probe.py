"""One FFN-1-shaped GEMM (M=8000, N=576, K=144, bias + swish + aux) launched a few times: target for `ncu --set full -k regex:gemm_bf16`."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tensorflowasr_b200 import kernels  # noqa: E402

m, n, k = 8000, 576, 144
a = torch.randn(m, k, device="cuda").to(torch.bfloat16)
b = torch.randn(n, k, device="cuda").to(torch.bfloat16)
bias = torch.randn(n, device="cuda")
out = torch.empty(m, n, dtype=torch.bfloat16, device="cuda")
aux = torch.empty_like(out)
for _ in range(6):
    kernels.gemm(a, b, out=out, bias=bias, activation="swish", aux=aux)
torch.cuda.synchronize()
