#!/usr/bin/env bash
# Quick single-GPU check after a kernel change: the GEMM / head / fused-module parity tests and the headline bench (no CPU arm).
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_lse.py tests/test_gpu_fused.py tests/test_gpu_ffn.py -m gpu -q --tb=short 2>&1 | tail -8 > gpurun_out/quick_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
tail -3 gpurun_out/quick_tests.log
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_quick.json").read().strip().splitlines()[-1])
print("ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["ms_per_step"], 3), {k: round(v, 3) for k, v in d["kernel_ms_per_step"].items()},
      "frac", round(d["roofline"]["frac"], 3), "grad", round(d["roofline"]["gradient_kernel"]["frac"], 3), "joint", round(d["roofline_joint_gemm"]["frac"], 3))
PY
tail -2 gpurun_out/bench_quick.err
