"""Thin Python launchers over the C ABI (include/tfasr_b200.h) for the dense / fused encoder kernels."""
from __future__ import annotations

import ctypes

import torch

from . import _lib

_ACT = {None: 0, "none": 0, "swish": 1, "relu": 2, "tanh": 3}
_MUL = {None: 0, "swish_grad": 1, "tanh_grad": 2, "mul": 3}


def gemm(a, b, *, out=None, out_dtype=torch.bfloat16, bias=None, activation=None, aux=None, residual=None, residual_scale=1.0,
         mul=None, mul_mode=None, colsum=None, glu=False, alpha=1.0, dropout_p=0.0, dropout_seed=0, mn_major=False):
    """C = epilogue(alpha * A . B^T) on the tcgen05 GEMM.

    default: a [M,K], b [N,K] (bf16, row-major).  mn_major=True: a [K,M], b [K,N] -> C [M,N] = a^T . b.
    """
    lib = _lib.load()
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16 and a.dim() == 2 and b.dim() == 2
    assert a.stride(1) == 1 and b.stride(1) == 1
    if mn_major:
        K, M = a.shape
        K2, N = b.shape
    else:
        M, K = a.shape
        N, K2 = b.shape
    assert K == K2, (a.shape, b.shape)
    n_out = N // 2 if glu else N
    if out is None:
        out = torch.empty((M, n_out), dtype=out_dtype, device=a.device)
    assert out.stride(1) == 1 and out.shape == (M, n_out)
    e = _lib.GemmEpilogue()
    e.bias = _lib.ptr(bias) if bias is not None else None
    e.aux = aux.data_ptr() if aux is not None else None
    e.ld_aux = aux.stride(0) if aux is not None else 0
    e.residual = residual.data_ptr() if residual is not None else None
    e.ld_residual = residual.stride(0) if residual is not None else 0
    e.mul = mul.data_ptr() if mul is not None else None
    e.ld_mul = mul.stride(0) if mul is not None else 0
    e.colsum = colsum.data_ptr() if colsum is not None else None
    e.c_f32 = int(out.dtype == torch.float32)
    e.activation = _ACT[activation]
    e.glu = int(bool(glu))
    e.mul_mode = _MUL[mul_mode]
    e.residual_scale = float(residual_scale)
    e.alpha = float(alpha)
    e.dropout_p = float(dropout_p)
    e.dropout_seed = int(dropout_seed) & 0xFFFFFFFFFFFFFFFF
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == N
    _lib.check(
        lib.tfasr_gemm_bf16(a.data_ptr(), b.data_ptr(), out.data_ptr(), M, N, K, a.stride(0), b.stride(0), out.stride(0),
                            int(mn_major), int(mn_major), ctypes.byref(e), _lib.stream_ptr(a.device)),
        "tfasr_gemm_bf16",
    )
    return out
