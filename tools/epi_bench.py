"""Epilogue-variant timings of the tcgen05 GEMM at the FFN shapes."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tensorflowasr_b200 import kernels
from tools.gemm_bench import time_us

M, D, Hd = 8000, 144, 576
a = torch.randn(M, D, device="cuda").to(torch.bfloat16)
w1t = torch.randn(Hd, D, device="cuda").to(torch.bfloat16)
h = torch.randn(M, Hd, device="cuda").to(torch.bfloat16)
w2t = torch.randn(D, Hd, device="cuda").to(torch.bfloat16)
b1 = torch.randn(Hd, device="cuda"); b2 = torch.randn(D, device="cuda")
z = torch.empty(M, Hd, dtype=torch.bfloat16, device="cuda")
x = torch.randn(M, D, device="cuda").to(torch.bfloat16)
seed = torch.zeros(1, dtype=torch.int64, device="cuda")
cs = torch.zeros(Hd, device="cuda")
dy = torch.randn(M, D, device="cuda").to(torch.bfloat16)
w2 = torch.randn(Hd, D, device="cuda").to(torch.bfloat16)
acc = torch.zeros(D, Hd, device="cuda")
cases = {
 "fwd1 plain": lambda: kernels.gemm(a, w1t),
 "fwd1 bias": lambda: kernels.gemm(a, w1t, bias=b1),
 "fwd1 bias+swish": lambda: kernels.gemm(a, w1t, bias=b1, activation="swish"),
 "fwd1 bias+swish+aux": lambda: kernels.gemm(a, w1t, bias=b1, activation="swish", aux=z),
 "fwd1 bias+swish+aux+drop": lambda: kernels.gemm(a, w1t, bias=b1, activation="swish", aux=z, dropout_p=0.1, dropout_seed=seed, dropout_salt=3),
 "fwd2 plain": lambda: kernels.gemm(h, w2t),
 "fwd2 bias+res": lambda: kernels.gemm(h, w2t, bias=b2, residual=x, residual_scale=0.5),
 "fwd2 bias+res+drop": lambda: kernels.gemm(h, w2t, bias=b2, residual=x, residual_scale=0.5, dropout_p=0.1, dropout_seed=seed, dropout_salt=3),
 "dgrad2 plain": lambda: kernels.gemm(dy, w2),
 "dgrad2 mul": lambda: kernels.gemm(dy, w2, mul=z, mul_mode="swish_grad"),
 "dgrad2 mul+drop": lambda: kernels.gemm(dy, w2, mul=z, mul_mode="swish_grad", dropout_p=0.1, dropout_seed=seed, dropout_salt=3),
 "dgrad2 mul+drop+colsum": lambda: kernels.gemm(dy, w2, mul=z, mul_mode="swish_grad", dropout_p=0.1, dropout_seed=seed, dropout_salt=3, colsum=cs),
 "wgrad ks15": lambda: kernels.gemm(a, z, mn_major=True, out=acc, accumulate=True, k_splits=15),
 "ln_fwd": lambda: kernels.layernorm_fwd(x, b2, b2),
 "scale_mask_colsum": lambda: kernels.scale_mask_colsum(dy, scale=0.5, dropout_p=0.1, seed=seed, salt=3, colsum=b2),
}
for k, f in cases.items():
    print(f"{k:28s} {time_us(f):8.1f} us", flush=True)
