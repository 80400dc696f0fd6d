#!/usr/bin/env bash
# Attention A/B on one GPU: the previous build's attention object (build/libtfasr_b200_prev.so) against the current one, plus the attention parity tests.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_fused.py -m gpu -q --tb=short -k "attention" 2>&1 | tail -5 > gpurun_out/att_tests.log
TFASR_LIB_PATH=$PWD/build/libtfasr_b200_prev.so timeout 200 python tools/att_fwd_bwd_bench.py > gpurun_out/att_prev.json 2> gpurun_out/att_prev.err
timeout 200 python tools/att_fwd_bwd_bench.py > gpurun_out/att_new.json 2> gpurun_out/att_new.err
tail -2 gpurun_out/att_tests.log; echo prev; cat gpurun_out/att_prev.json; tail -2 gpurun_out/att_prev.err; echo new; cat gpurun_out/att_new.json; tail -2 gpurun_out/att_new.err
