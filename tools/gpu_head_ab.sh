#!/usr/bin/env bash
# One GPU box call: GPU suite at the defaults, oracle parity with the new head switches on, headline bench A/B, ncu of the head kernels.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
NCU="ncu --clock-control none"
timeout 900 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -30 > gpurun_out/gpu_tests.log
echo "=== pytest exit ${PIPESTATUS[0]}" >> gpurun_out/gpu_tests.log
TFASR_BF16_LOGITS=1 TFASR_LATTICE_V2=1 timeout 600 python -m pytest tests/test_gpu_bf16_oracle.py tests/test_gpu_model.py tests/test_gpu_fused.py -m gpu -q -s --tb=short 2>&1 | tail -60 > gpurun_out/gpu_tests_new_head.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
TFASR_LATTICE_V2=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_lattice_v2.json 2> gpurun_out/bench_lattice_v2.err
TFASR_BF16_LOGITS=1 TFASR_LATTICE_V2=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_new_head.json 2> gpurun_out/bench_new_head.err
TFASR_BF16_LOGITS=1 TFASR_LATTICE_V2=1 timeout 600 $NCU --profile-from-start off --set full --import-source on -k regex:'rnnt_' -c 4 -o gpurun_out/head_new_r02 python tools/profile_step.py --ncu > gpurun_out/ncu_head.log 2>&1
tail -4 gpurun_out/gpu_tests.log; tail -12 gpurun_out/gpu_tests_new_head.log
for f in default lattice_v2 new_head; do python - "$f" <<'PY'
import json, sys
f = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/bench_{f}.json").read().strip().splitlines()[-1])
    print(f, "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["ms_per_step"], 3), "serial", round(d["e2e"]["unpipelined_ms_per_step"], 3),
          {k: round(v, 3) for k, v in d["kernel_ms_per_step"].items()}, "frac", round(d["roofline"]["frac"], 3))
except Exception as e:
    print(f, "FAILED", e); print(open(f"gpurun_out/bench_{f}.err").read()[-1500:])
PY
done
tail -3 gpurun_out/ncu_head.log
