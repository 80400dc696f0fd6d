#!/usr/bin/env bash
# One GPU box call: the GPU suite, the headline bench (full line incl. the CPU baseline) and the reference arm.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -30 > gpurun_out/gpu_tests.log
echo "=== pytest exit ${PIPESTATUS[0]}" >> gpurun_out/gpu_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
( time timeout 900 python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
tail -4 gpurun_out/gpu_tests.log; cat gpurun_out/bench_default.json | cut -c1-1500; cat gpurun_out/bench_reference.json | cut -c1-900; tail -3 gpurun_out/bench_reference.err; tail -2 gpurun_out/smoke.log
