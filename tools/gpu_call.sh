#!/usr/bin/env bash
# One GPU box call: the GPU suite, the FFN microbenchmark and the headline bench with a few switch variants.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_call.sh'
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -q --tb=short -s 2>&1 | tail -150 > gpurun_out/gpu_tests.log
echo "=== pytest exit ${PIPESTATUS[0]}" >> gpurun_out/gpu_tests.log
timeout 200 python tools/ffn_bench.py > gpurun_out/ffn_bench.json 2> gpurun_out/ffn_bench.err
timeout 200 python tools/att_bench.py > gpurun_out/att_bench.json 2> gpurun_out/att_bench.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
TFASR_SIDE_PRED=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_nosidepred.json 2> gpurun_out/bench_nosidepred.err
TFASR_FUSED_SEQCONV=1 timeout 300 python -m pytest tests/test_gpu_seqconv.py tests/test_gpu_model.py -m gpu -q --tb=short 2>&1 | tail -15 > gpurun_out/seqconv_on.log
for c in ds2 contextnet decode; do timeout 400 python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/bench_cfg_$c.json 2> gpurun_out/bench_cfg_$c.err; tail -3 gpurun_out/bench_cfg_$c.err; done
TFASR_FUSED_FFN=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_noffn.json 2> gpurun_out/bench_noffn.err
TFASR_ATTENTION_TC=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_noatttc.json 2> gpurun_out/bench_noatttc.err
for f in gpurun_out/bench_*.json; do echo "$f: $(python -c "import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print(round(d['ms_per_step'],3), 'ms', round(d['value'],1), 'utt/s', 'rnnt frac', round((d.get('roofline') or {}).get('frac', 0), 3))" 2>/dev/null)"; done | tee gpurun_out/bench_variants.txt
tail -5 gpurun_out/gpu_tests.log; cat gpurun_out/ffn_bench.json gpurun_out/att_bench.json
