"""Builds tests/emu's CPU emulation of the plain-SIMT kernels (g++, no CUDA): build/emu/libtfasr_emu[_asan].so.

Test infrastructure only — the same csrc/*.cu launcher files, compiled with -DTFASR_EMU_BUILD against tests/emu/cuda_emu.h, export the
same C ABI symbols but take HOST pointers and run every CUDA thread as an OS thread."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(REPO, "tensorflowasr_b200", "csrc")
OUT = os.path.join(REPO, "build", "emu")
SOURCES = ["seqconv.cu", "decode.cu", "rnnt_combine.cu", "conv1.cu", "rnnt_grad_v2.cu", "joint_v2.cu", "subsample_v2.cu", "dwconv_wgrad_v2.cu"]  # launcher files written against simt.cuh


def build(asan: bool = False, tsan: bool = False) -> str:
    gxx = shutil.which("g++")
    if gxx is None:
        raise RuntimeError("g++ not found")
    os.makedirs(OUT, exist_ok=True)
    lib = os.path.join(OUT, "libtfasr_emu_asan.so" if asan else ("libtfasr_emu_tsan.so" if tsan else "libtfasr_emu.so"))
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    deps = srcs + [os.path.join(CSRC, "simt.cuh"), os.path.join(CSRC, "seqconv_kernels.cuh"), os.path.join(CSRC, "decode_kernels.cuh"), os.path.join(CSRC, "rnnt_combine.cu"), os.path.join(CSRC, "conv1.cu"), os.path.join(CSRC, "rnnt_grad_v2.cu"), os.path.join(CSRC, "joint_v2.cu"), os.path.join(CSRC, "subsample_v2.cu"), os.path.join(CSRC, "dwconv_wgrad_v2.cu"), os.path.join(HERE, "cuda_emu.h"),
                   os.path.join(REPO, "include", "tfasr_b200.h"), os.path.abspath(__file__)]
    if os.path.exists(lib) and os.path.getmtime(lib) >= max(os.path.getmtime(d) for d in deps):
        return lib
    cmd = [gxx, "-std=c++20", "-O1", "-g", "-fPIC", "-shared", "-pthread", "-DTFASR_EMU_BUILD", "-I", HERE, "-I", os.path.join(REPO, "include")]
    if asan:
        cmd += ["-fsanitize=address", "-fno-omit-frame-pointer"]
    if tsan:
        cmd += ["-fsanitize=thread", "-fno-omit-frame-pointer"]
    for s in srcs:
        cmd += ["-x", "c++", s]
    cmd += ["-o", lib + ".tmp"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("emulation build failed:\n" + r.stdout + r.stderr)
    os.replace(lib + ".tmp", lib)
    return lib


def tsan_runtime() -> str | None:
    r = subprocess.run(["gcc", "-print-file-name=libtsan.so"], capture_output=True, text=True)
    p = r.stdout.strip()
    return os.path.realpath(p) if r.returncode == 0 and os.path.isabs(p) and os.path.exists(p) else None


def asan_runtime() -> str | None:
    r = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True)
    p = r.stdout.strip()
    return os.path.realpath(p) if r.returncode == 0 and os.path.isabs(p) and os.path.exists(p) else None


if __name__ == "__main__":
    print(build(), build(asan=True), build(tsan=True))
