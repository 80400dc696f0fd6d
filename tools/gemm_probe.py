"""One encoder-shaped GEMM launched a few times: target for `ncu --set full -k regex:gemm_bf16 --launch-skip 4 -c 1`.
usage: gemm_probe.py [plain|swish_aux|drop_res|joint]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tensorflowasr_b200 import kernels  # noqa: E402

variant = sys.argv[1] if len(sys.argv) > 1 else "swish_aux"
m, n, k = (8000, 144, 576) if variant == "drop_res" else (8000, 576, 144)
if variant == "joint":  # the joint's vocabulary projection of the bench workload: [B T' (U+1), 320] x [320, 1000] + bias -> f32 logits
    m, n, k = 32 * 250 * 65, 1000, 320
a = torch.randn(m, k, device="cuda").to(torch.bfloat16)
b = torch.randn(n, k, device="cuda").to(torch.bfloat16)
bias = torch.randn(n, device="cuda")
out = torch.empty(m, n, dtype=torch.float32 if variant == "joint" else torch.bfloat16, device="cuda")
extra = {}
if variant == "joint":
    extra = dict(bias=bias)
elif variant == "swish_aux":
    extra = dict(bias=bias, activation="swish", aux=torch.empty_like(out))
elif variant == "drop_res":
    extra = dict(bias=bias, residual=torch.randn(m, n, device="cuda").to(torch.bfloat16), residual_scale=0.5, dropout_p=0.1,
                 dropout_seed=torch.zeros(1, dtype=torch.int64, device="cuda"), dropout_salt=3)
for _ in range(6):
    kernels.gemm(a, b, out=out, **extra)
torch.cuda.synchronize()
