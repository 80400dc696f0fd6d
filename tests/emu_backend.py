"""Run the PRODUCT's Python stack on the CPU — TEST INFRASTRUCTURE ONLY (tests/test_conformer_emulated_cpu.py).

`install()` points `_lib.load()` at the emulator build of the kernel sources (tests/emu: EVERY hand-written kernel, the tcgen05 GEMM included,
bound with the product's own ctypes signatures) and lets the launchers take CPU tensors (host pointers, no stream).  Everything else —
fused.py, encoder.py, transducer.py, optim.py, features.py, the kernels.py launchers — runs unchanged.  `gemm` below is a torch restatement
of the GEMM's epilogue contract (bias, activation + aux copy, GLU, residual, backward multipliers, column sums, fp32 accumulate-into, row
softmax statistics), kept as an independent second opinion (`install(real_gemm=False)`)."""
from __future__ import annotations

import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "emu"))

BF = torch.bfloat16


def _act(z, name):
    if name == "swish":
        return z * torch.sigmoid(z)
    if name == "relu":
        return torch.relu(z)
    if name == "tanh":
        return torch.tanh(z)
    return z


def gemm(a, b, *, out=None, out_dtype=BF, bias=None, activation=None, aux=None, residual=None, residual_scale=1.0, mul=None, mul_mode=None,
         colsum=None, glu=False, alpha=1.0, dropout_p=0.0, dropout_seed=None, dropout_salt=0, mn_major=False, accumulate=False, k_splits=1,
         rowstats=None):
    """The epilogue order of csrc/gemm_kernel.cuh: alpha * acc + bias -> (row statistics) -> aux copy -> activation | GLU -> backward
    multiplier -> dropout -> residual -> column sums -> store / accumulate."""
    assert a.dtype == BF and b.dtype == BF and dropout_p == 0.0, "the CPU stand-in has no dropout stream: run with dropout 0"
    v = (a.float().t() @ b.float()) if mn_major else (a.float() @ b.float().t())
    v = v * alpha
    if bias is not None:
        v = v + bias.float().reshape(-1)
    if rowstats is not None:
        M, N = v.shape
        if out_dtype == BF and out is None or (out is not None and out.dtype == BF):
            v = v.to(BF).float()  # bf16 output: the statistics describe the stored values
        for s in range(rowstats.shape[1]):
            nb, half = s // 2, s % 2
            n_cols = min(256, N - nb * 256)
            n_groups = (n_cols + 31) // 32
            g_mid = (n_groups + 1) // 2
            g0, g1 = (0, g_mid) if half == 0 else (g_mid, n_groups)
            c0, c1 = nb * 256 + g0 * 32, nb * 256 + min(n_cols, g1 * 32)
            if c1 <= c0:
                rowstats[:, s, 0], rowstats[:, s, 1] = float("-inf"), 0.0
            else:
                m = v[:, c0:c1].max(1).values
                rowstats[:, s, 0] = m
                rowstats[:, s, 1] = torch.exp(v[:, c0:c1] - m[:, None]).sum(1)
    if aux is not None:
        aux.copy_(v.to(aux.dtype))
    if glu:
        v = v[:, 0::2] * torch.sigmoid(v[:, 1::2])
    else:
        v = _act(v, activation)
    if mul is not None:
        z = mul.float()
        if mul_mode == "swish_grad":
            sg = torch.sigmoid(z)
            v = v * sg * (1 + z * (1 - sg))
        elif mul_mode == "tanh_grad":
            v = v * (1 - z * z)
        else:
            v = v * z
    if residual is not None:
        v = residual.float() + residual_scale * v
    if colsum is not None:
        colsum.view(-1).add_(v.sum(0))
    if out is None:
        return v.to(out_dtype)
    assert out.shape == v.shape, (out.shape, v.shape)
    if accumulate:
        assert out.dtype == torch.float32
        out.add_(v)
    else:
        out.copy_(v.to(out.dtype))
    return out


class _EmuLib:
    """The emulator library with the product's ctypes signatures; the library-level entry points of api.cu answered here."""

    def __init__(self, lib, signatures):
        self._lib = lib
        self.calls = {}  # entry point -> number of calls (lets a test see which kernels a switch really routed to)
        for name, (res, args) in signatures.items():
            if hasattr(lib, name):
                fn = getattr(lib, name)
                fn.restype, fn.argtypes = res, args

    def __getattr__(self, name):
        if name == "tfasr_version":
            return lambda: 100
        if name == "tfasr_status_string":
            return lambda st: {0: b"ok", 1: b"invalid value", 2: b"execution failed", 3: b"workspace too small"}.get(int(st), b"unknown")
        if name in ("tfasr_launch_count", "tfasr_launch_count_reset"):
            return lambda *a: 0
        fn = getattr(self._lib, name)

        def counted(*a):
            self.calls[name] = self.calls.get(name, 0) + 1
            return fn(*a)

        return counted


def install(monkeypatch, real_gemm=True):
    """real_gemm=True: kernels.gemm stays the product launcher and runs the tcgen05 GEMM's own source on the emulator (tests/emu/emu_sm100.h);
    False: the torch restatement of its epilogue contract above (an independent second opinion)."""
    import checks_v1

    from tensorflowasr_b200 import _lib, kernels, ops

    emu = _EmuLib(checks_v1.load(), _lib.SIGNATURES)
    monkeypatch.setattr(_lib, "load", lambda: emu)
    monkeypatch.setattr(_lib, "stream_ptr", lambda device=None: None)
    monkeypatch.setattr(_lib, "require_cuda", lambda t, what: None)
    monkeypatch.setattr(_lib, "EMULATED", True)  # lets make_optimizer build the one-launch FlatAdamW over host tensors

    def ptr(t):
        if t is None:
            return None
        assert not t.is_cuda and t.is_contiguous(), "emulated kernels take contiguous host tensors"
        return t.data_ptr()

    monkeypatch.setattr(_lib, "ptr", ptr)

    def ptr_rows(t):
        if t is None:
            return None
        assert not t.is_cuda and t.dim() == 2 and t.stride(1) == 1, "emulated kernels take host tensors with unit column stride"
        return t.data_ptr()

    monkeypatch.setattr(_lib, "ptr_rows", ptr_rows)
    monkeypatch.setattr(torch.cuda, "is_current_stream_capturing", lambda: False)  # no CUDA driver here
    if not real_gemm:
        monkeypatch.setattr(kernels, "gemm", gemm)
    monkeypatch.setattr(ops, "fused_active", lambda x: ops._FUSED and ops.precision() == "bf16")
    return emu
