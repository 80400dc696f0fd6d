#!/usr/bin/env bash
# One GPU box call for the bf16-logits head: the GPU suite at the defaults, the oracle parity tests with the switch on, the headline bench both
# ways, and an ncu capture of the head kernels with the switch on.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
NCU="ncu --clock-control none"
timeout 900 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -30 > gpurun_out/gpu_tests.log
echo "=== pytest exit ${PIPESTATUS[0]}" >> gpurun_out/gpu_tests.log
TFASR_BF16_LOGITS=1 timeout 600 python -m pytest tests/test_gpu_bf16_oracle.py tests/test_gpu_model.py tests/test_gpu_fused.py -m gpu -q -s --tb=short 2>&1 | tail -60 > gpurun_out/gpu_tests_bf16_logits.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
TFASR_BF16_LOGITS=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_bf16_logits.json 2> gpurun_out/bench_bf16_logits.err
TFASR_BF16_LOGITS=1 timeout 600 $NCU --profile-from-start off --set full --import-source on -k regex:'rnnt_' -c 4 -o gpurun_out/head_bf16_r02 python tools/profile_step.py --ncu > gpurun_out/ncu_head.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
tail -4 gpurun_out/gpu_tests.log; tail -25 gpurun_out/gpu_tests_bf16_logits.log; cut -c1-1800 gpurun_out/bench_default.json; cut -c1-1800 gpurun_out/bench_bf16_logits.json; tail -3 gpurun_out/bench_bf16_logits.err; tail -2 gpurun_out/smoke.log; tail -5 gpurun_out/ncu_head.log
