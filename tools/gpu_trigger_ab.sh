#!/usr/bin/env bash
# Early trigger of programmatic dependents in the persistent tcgen05 kernels (TFASR_PDL_TRIGGER, default on): full GPU suite with it on,
# headline bench off / on, smoke.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -30 > gpurun_out/gpu_tests.log
echo "=== pytest exit ${PIPESTATUS[0]}" >> gpurun_out/gpu_tests.log
TFASR_PDL_TRIGGER=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_trig0.json 2> gpurun_out/bench_trig0.err
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
tail -4 gpurun_out/gpu_tests.log; tail -1 gpurun_out/smoke.log | cut -c1-300
for f in trig0 default; do python - $f <<'PY'
import json, sys
f = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/bench_{f}.json").read().strip().splitlines()[-1])
    print(f, "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["ms_per_step"], 3), "serial", round(d["e2e"]["unpipelined_ms_per_step"], 3))
except Exception as e:
    print(f, "FAILED", e); print(open(f"gpurun_out/bench_{f}.err").read()[-1500:])
PY
done
