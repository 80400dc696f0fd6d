"""Phase timeline of the tcgen05 GEMM on the encoder's small shapes (debug build of the library with -DTFASR_GEMM_TRACE).

    python tools/gemm_trace.py build     # here: compiles tools/libtfasr_trace.so
    python tools/gemm_trace.py           # on the GPU box: prints per-phase times (ns since the first CTA started)
slots: 0 entry, 1 prologue done, 2 first operand stage landed, 3/4 MMA committed tile 0/1, 5/6 epilogue tile 0 start/end,
       7/8 epilogue tile 1 start/end, 9 stores drained, 10 exit
"""
import ctypes
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
TRACE_LIB = os.path.join(ROOT, "tools", "libtfasr_trace.so")


def build():
    from tensorflowasr_b200 import _build

    objs = []
    for src in sorted(glob.glob(os.path.join(_build.CSRC, "*.cu"))):
        obj = os.path.join("/tmp", "trace_" + os.path.basename(src)[:-3] + ".o")
        subprocess.run([_build._nvcc(), *_build.NVCC_FLAGS, "-DTFASR_GEMM_TRACE", "-c", src, "-o", obj], check=True)
        objs.append(obj)
    subprocess.run([_build._nvcc(), "-shared", "-o", TRACE_LIB, *objs, "-lcudart"], check=True)
    print(TRACE_LIB)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "build":
        return build()
    import numpy as np
    import torch

    from tensorflowasr_b200 import _lib, kernels

    _lib.LIB_PATH = TRACE_LIB
    lib = _lib.load()
    lib.tfasr_debug_set_gemm_trace.argtypes = [ctypes.c_void_p]
    trace = torch.zeros(148 * 16, dtype=torch.int64, device="cuda")
    names = ["entry", "prologue", "stage0", "mma_t0", "mma_t1", "epi0_beg", "epi0_end", "epi1_beg", "epi1_end", "drained", "exit"]
    for (m, n, k, kw) in [(8000, 576, 144, {}), (8000, 576, 144, {"bias": True, "swish_aux": True}), (8000, 144, 576, {"bias": True, "drop_res": True}),
                          (8000, 144, 144, {})]:
        a = torch.randn(m, k, device="cuda").to(torch.bfloat16)
        b = torch.randn(n, k, device="cuda").to(torch.bfloat16)
        bias = torch.randn(n, device="cuda") if kw.get("bias") else None
        out = torch.empty(m, n, dtype=torch.bfloat16, device="cuda")
        extra = {}
        if kw.get("swish_aux"):
            extra = dict(activation="swish", aux=torch.empty_like(out))
        if kw.get("drop_res"):
            extra = dict(residual=torch.randn(m, n, device="cuda").to(torch.bfloat16), residual_scale=0.5, dropout_p=0.1,
                         dropout_seed=torch.zeros(1, dtype=torch.int64, device="cuda"), dropout_salt=3)
        for _ in range(5):
            kernels.gemm(a, b, out=out, bias=bias, **extra)
        torch.cuda.synchronize()
        trace.zero_()
        lib.tfasr_debug_set_gemm_trace(trace.data_ptr())
        kernels.gemm(a, b, out=out, bias=bias, **extra)
        torch.cuda.synchronize()
        lib.tfasr_debug_set_gemm_trace(None)
        t = trace.cpu().numpy().reshape(148, 16).astype(np.int64)
        used = t[:, 0] > 0
        t0 = t[used, 0].min()
        print(f"\n{m}x{n}x{k} {kw}: CTAs {used.sum()}, kernel span {t[used, 10].max() - t0} ns")
        for i, nm in enumerate(names):
            col = t[used, i]
            col = col[col > 0] - t0
            if len(col):
                print(f"  {nm:9s} n={len(col):3d}  min {col.min():6d}  median {int(np.median(col)):6d}  max {col.max():6d}")


if __name__ == "__main__":
    main()
