#!/usr/bin/env bash
# One GPU box call: the GPU suite, microbenchmarks and the headline bench with a few switch variants.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_call.sh'
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -q --tb=short -s 2>&1 | tail -150 > gpurun_out/gpu_tests.log
echo "=== pytest exit ${PIPESTATUS[0]}" >> gpurun_out/gpu_tests.log
timeout 200 python tools/ffn_bench.py > gpurun_out/ffn_bench.json 2> gpurun_out/ffn_bench.err
timeout 200 python tools/ffn_trace.py > gpurun_out/ffn_trace.txt 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
TFASR_FUSED_FFN=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_noffn.json 2> gpurun_out/bench_noffn.err
for f in gpurun_out/bench_*.json; do echo "$f: $(python -c "import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print(round(d['ms_per_step'],3), 'ms', round(d['value'],1), 'utt/s', 'rnnt frac', round((d.get('roofline') or {}).get('frac', 0), 3))" 2>/dev/null)"; done | tee gpurun_out/bench_variants.txt
tail -5 gpurun_out/gpu_tests.log; cat gpurun_out/ffn_bench.json; cat gpurun_out/ffn_trace.txt
