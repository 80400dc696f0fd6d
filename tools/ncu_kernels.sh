#!/usr/bin/env bash
# ncu --set full captures of the new tcgen05 kernels (one launch each), brought back as .ncu-rep for `ncu -i ... --page raw/source` here.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ffn_kernel -s 2 -c 2 -o gpurun_out/ffn_r02 python tools/ffn_bench.py > gpurun_out/ncu_ffn.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rel_attention_fwd_tc -s 1 -c 1 -o gpurun_out/att_tc_r02 python tools/att_debug.py 32 250 4 36 > gpurun_out/ncu_att.log 2>&1
timeout 300 python -m pytest tests/test_gpu_bf16_oracle.py -m gpu -q -s --tb=short 2>&1 | tail -40 > gpurun_out/parity_full.log
ls -la gpurun_out; tail -30 gpurun_out/parity_full.log
