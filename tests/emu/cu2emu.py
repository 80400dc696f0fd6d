"""Source-to-source shim: a first-generation csrc/*.cu file (written against common.cuh, launching with <<<>>>) -> a .cpp that g++ compiles
against the SIMT emulator.  TEST INFRASTRUCTURE ONLY.  Rewrites, and nothing else:
  * `#include "common.cuh"`                           -> `#include "emu_common.h"`
  * `kernel<<<grid, block[, smem[, stream]]>>>(args);` -> `emu::launch(dim3(grid), dim3(block), smem, [&] { kernel(args); });`
  * `extern __shared__ [__align__(N)] T name[];`       -> `T* name = reinterpret_cast<T*>(emu::t_dyn_smem);`
  * kernels declared `__cluster_dims__(n, 1, 1)` launch through emu::launch_cluster (the n CTAs of a cluster run concurrently)
  * the tanh.approx inline PTX of joint.cu             -> tanhf
  * `#include <cuda_*.h>` lines dropped; `template <> __device__` loses the __device__ (it maps to `static` in the emulator)
Anything it does not understand (another `asm`, a launch it cannot parse) is an error, not a silent pass-through."""
from __future__ import annotations

import re


def _match_paren(s: str, i: int) -> int:
    """s[i] == '(' -> index of the matching ')'."""
    depth = 0
    for j in range(i, len(s)):
        if s[j] == "(":
            depth += 1
        elif s[j] == ")":
            depth -= 1
            if depth == 0:
                return j
    raise ValueError("unbalanced parentheses")


def _split_top(s: str) -> list[str]:
    out, depth, cur = [], 0, []
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append("".join(cur).strip())
            cur = []
        else:
            cur.append(ch)
    out.append("".join(cur).strip())
    return out


def _rewrite_launches(src: str, clusters: dict | None = None) -> str:
    clusters = clusters or {}
    out, pos = [], 0
    while True:
        i = src.find("<<<", pos)
        if i < 0:
            out.append(src[pos:])
            return "".join(out)
        # kernel expression: identifier (+ optional <template args>) right before <<<
        j = i
        if src[j - 1] == ">":  # template arguments
            depth = 0
            while True:
                j -= 1
                if src[j] == ">":
                    depth += 1
                elif src[j] == "<":
                    depth -= 1
                    if depth == 0:
                        break
        k = j
        while k > 0 and (src[k - 1].isalnum() or src[k - 1] in "_:"):
            k -= 1
        kernel = src[k:i]
        e = src.find(">>>", i)
        cfg = _split_top(src[i + 3:e])
        if not (2 <= len(cfg) <= 4):
            raise ValueError(f"cannot parse launch configuration of {kernel!r}: {cfg}")
        a0 = src.index("(", e)
        a1 = _match_paren(src, a0)
        semi = a1 + 1
        while src[semi].isspace():
            semi += 1
        smem = cfg[2] if len(cfg) > 2 else "0"
        out.append(src[pos:k])
        base = kernel.split("<")[0].split("::")[-1]
        if base in clusters:  # declared with __cluster_dims__(n, 1, 1): the n CTAs of a cluster run concurrently
            call = f"emu::launch_cluster(dim3({cfg[0]}), dim3({cfg[1]}), (size_t)({smem}), {clusters[base]}, [&] {{ {kernel}({src[a0 + 1:a1]}); }})"
        else:
            call = f"emu::launch(dim3({cfg[0]}), dim3({cfg[1]}), (size_t)({smem}), [&] {{ {kernel}({src[a0 + 1:a1]}); }})"
        if src[semi] == ";":
            out.append(call + ";")
            pos = semi + 1
        else:  # the launch is the tail of a macro body: the `;` comes from the macro's use
            out.append(call)
            pos = a1 + 1


def convert(src: str, name: str = "?") -> str:
    src = src.replace('#include "common.cuh"', '#include "emu_common.h"')
    src = src.replace('#include "attention_mma.cuh"', '#include "emu_attention_mma.h"')  # ldmatrix / mma.sync / cp.async on the emulator
    src = re.sub(r"#include\s+<(cuda[\w.]*|math_constants\.h)>\n", "", src)
    src = re.sub(r"template\s*<>\s*__device__", "template <>", src)  # an explicit specialisation cannot repeat the storage class
    src = re.sub(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?(\w+)\s+(\w+)\[\];", r"\1* \2 = reinterpret_cast<\1*>(emu::t_dyn_smem);", src)
    src = src.replace("#include <cooperative_groups.h>\n", "")  # emu_common.h carries the cluster slice of it
    clusters = {m.group(2): m.group(1).strip() for m in
                re.finditer(r"__cluster_dims__\(([^,)]+)[^)]*\)\s*(?:__launch_bounds__\([^)]*\)\s*)?(\w+)\s*\(", src)}
    src = re.sub(r'asm\("tanh\.approx\.f32 %0, %1;"\s*:\s*"=f"\((\w+)\)\s*:\s*"f"\(([^;]+)\)\);', r"\1 = tanhf(\2);", src)
    # ---- the tcgen05 GEMM (gemm_kernel.cuh / gemm.cu / gemm_inst_*.cu): sm100.cuh -> emu_sm100.h, the few other PTX lines by hand
    src = re.sub(r"__maxnreg__\(\d+\)", "", src)  # a register budget means nothing on the emulator
    src = src.replace('#include "sm100.cuh"', '#include "emu_sm100.h"')
    src = re.sub(r'asm volatile\("bar\.sync (\d+), (\d+);" ::: "memory"\);', r"emu::named_barrier_sync(\1, \2);", src)
    src = src.replace('asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));', "t = 0;")
    src = re.sub(r'const __half2 h = __floats2half2_rn\(0\.5f \* a, 0\.5f \* b\);.*?sb = fmaf\(0\.5f, tf\.y, 0\.5f\);',
                 "sa = fmaf(0.5f, tanhf(0.5f * a), 0.5f);\n  sb = fmaf(0.5f, tanhf(0.5f * b), 0.5f);", src, flags=re.S)
    # the host tensor-map encoders of gemm.cu (cuTensorMapEncodeTiled through the driver) are provided by emu_sm100.h
    src = re.sub(r"// -+ host: tensor maps\n.*?(?=\n// kernels are instantiated per tile width)", "", src, flags=re.S)
    if re.search(r"\basm\b", src):
        raise ValueError(f"{name}: inline PTX the shim does not know")
    src = _rewrite_launches(src, clusters)
    if "<<<" in src or "extern __shared__" in src:
        raise ValueError(f"{name}: unconverted CUDA syntax left")
    return src


if __name__ == "__main__":
    import sys

    print(convert(open(sys.argv[1]).read(), sys.argv[1]))
