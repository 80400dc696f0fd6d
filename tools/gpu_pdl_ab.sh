#!/usr/bin/env bash
# Programmatic dependent launch of the tcgen05 GEMM (TFASR_PDL=1): parity tests with it on, headline bench off / on.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
TFASR_PDL=1 timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_fused.py tests/test_gpu_lse.py -m gpu -q --tb=short 2>&1 | tail -8 > gpurun_out/pdl_tests.log
TFASR_PDL=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_pdl0.json 2> gpurun_out/bench_pdl0.err
TFASR_PDL=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_pdl1.json 2> gpurun_out/bench_pdl1.err
TFASR_PDL=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_pdl1b.json 2> gpurun_out/bench_pdl1b.err
tail -3 gpurun_out/pdl_tests.log
for f in pdl0 pdl1 pdl1b; do python - $f <<'PY'
import json, sys
f = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/bench_{f}.json").read().strip().splitlines()[-1])
    print(f, "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["ms_per_step"], 3))
except Exception as e:
    print(f, "FAILED", e); print(open(f"gpurun_out/bench_{f}.err").read()[-1200:])
PY
done
