#!/usr/bin/env bash
# Validation of programmatic dependent launch as the default (TFASR_PDL unset = on): full GPU suite, bench with it off and on, smoke.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -30 > gpurun_out/gpu_tests.log
echo "=== pytest exit ${PIPESTATUS[0]}" >> gpurun_out/gpu_tests.log
TFASR_PDL=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_pdl0.json 2> gpurun_out/bench_pdl0.err
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
tail -4 gpurun_out/gpu_tests.log; tail -1 gpurun_out/smoke.log | cut -c1-300
for f in pdl0 default; do python - $f <<'PY'
import json, sys
f = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/bench_{f}.json").read().strip().splitlines()[-1])
    print(f, "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["ms_per_step"], 3), {k: round(v, 3) for k, v in d["kernel_ms_per_step"].items()}, "launches", d["gpu_launches"])
except Exception as e:
    print(f, "FAILED", e); print(open(f"gpurun_out/bench_{f}.err").read()[-1500:])
PY
done
