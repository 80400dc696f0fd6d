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
SOURCES = ["seqconv.cu", "decode.cu", "rnnt_combine.cu", "conv1.cu", "rnnt_grad_v2.cu", "rnnt_lattice_v2.cu", "joint_v2.cu", "subsample_v2.cu", "dwconv_wgrad_v2.cu", "peer_allreduce.cu"]  # launcher files written against simt.cuh


V1_SOURCES = ["rnnt_loss.cu", "ctc_loss.cu", "logmel.cu", "fused_rows.cu", "conv_bn.cu", "dwconv.cu", "joint.cu", "param_ops.cu", "specaugment.cu", "attention.cu", "lstm.cu", "gemm.cu", "gemm_inst_64.cu", "gemm_inst_128.cu",
              "gemm_inst_160.cu", "gemm_inst_192.cu", "gemm_inst_256.cu", "ffn_fused.cu", "attention_tc.cu"]
V1_HEADERS = ["gemm_kernel.cuh"]  # converted next to the sources that include them
# first-generation files (written against common.cuh, <<<>>> launches): converted by cu2emu.py into build/emu/gen/*.cpp


def _generate_v1() -> list[str]:
    import cu2emu

    gen = os.path.join(OUT, "gen")
    os.makedirs(gen, exist_ok=True)
    outs = []
    for name in V1_HEADERS:
        dst = os.path.join(gen, name)
        text = cu2emu.convert(open(os.path.join(CSRC, name)).read(), name)
        if not os.path.exists(dst) or open(dst).read() != text:
            open(dst, "w").write(text)
    for name in V1_SOURCES:
        src = os.path.join(CSRC, name)
        dst = os.path.join(gen, name[:-3] + ".cpp")
        text = cu2emu.convert(open(src).read(), name)
        if not os.path.exists(dst) or open(dst).read() != text:
            open(dst, "w").write(text)
        outs.append(dst)
    return outs


def build(asan: bool = False, tsan: bool = False) -> str:
    """Compile every source to its own object (in parallel) and link build/emu/libtfasr_emu[_asan|_tsan].so; returns the path."""
    import concurrent.futures
    import sys

    gxx = shutil.which("g++")
    if gxx is None:
        raise RuntimeError("g++ not found")
    sys.path.insert(0, HERE)
    tag = "asan" if asan else ("tsan" if tsan else "plain")
    objdir = os.path.join(OUT, "obj_" + tag)
    os.makedirs(objdir, exist_ok=True)
    lib = os.path.join(OUT, "libtfasr_emu.so" if tag == "plain" else f"libtfasr_emu_{tag}.so")
    srcs = [os.path.join(CSRC, s) for s in SOURCES] + _generate_v1() + [os.path.join(HERE, "emu_support.cpp")]
    hdrs = [os.path.join(CSRC, h) for h in ("simt.cuh", "seqconv_kernels.cuh", "decode_kernels.cuh", "dropout.cuh")] + \
           [os.path.join(HERE, h) for h in ("cuda_emu.h", "emu_common.h", "emu_attention_mma.h", "emu_sm100.h", "cu2emu.py")] + \
           [os.path.join(CSRC, "gemm_kernel.cuh")] + [os.path.join(REPO, "include", "tfasr_b200.h"), os.path.abspath(__file__)]
    hdr_m = max(os.path.getmtime(h) for h in hdrs)
    flags = ["-std=c++20", "-O1", "-g", "-fPIC", "-pthread", "-DTFASR_EMU_BUILD", "-I", HERE, "-I", os.path.join(REPO, "include"), "-I", CSRC]
    if asan:
        flags += ["-fsanitize=address", "-fno-omit-frame-pointer"]
    if tsan:
        flags += ["-fsanitize=thread", "-fno-omit-frame-pointer"]
    jobs, objs = [], []
    for src in srcs:
        obj = os.path.join(objdir, os.path.splitext(os.path.basename(src))[0] + ".o")
        objs.append(obj)
        if not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_m):
            jobs.append([gxx, *flags, "-c", "-x", "c++", src, "-o", obj])

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("emulation build failed: " + " ".join(cmd) + "\n" + r.stdout + r.stderr)

    if jobs:
        with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(run, jobs))
    if jobs or not os.path.exists(lib) or any(os.path.getmtime(o) > os.path.getmtime(lib) for o in objs):
        link = [gxx, "-shared", "-pthread", "-o", lib + ".tmp", *objs]
        if asan:
            link.append("-fsanitize=address")
        if tsan:
            link.append("-fsanitize=thread")
        run(link)
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
