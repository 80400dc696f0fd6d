#!/usr/bin/env bash
# First gpurun call of the next round: everything that was written without GPU time in round 1, in one box.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/round2_first_call.sh'
# Results land in gpurun_out/ (merged back): r2_unverified_tests.log, seqconv_bench.json, decode_bench.json, ncu launch lists.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
# 1. the GPU parity tests marked xfail(strict=False) in round 1, as real tests (--runxfail), each file in its own process so that a
#    sticky CUDA error in one cannot poison the other
for f in tests/test_gpu_zz_seqconv.py tests/test_gpu_zz_decode.py tests/test_gpu_zz_lse.py; do
  timeout 600 python -m pytest "$f" -m gpu --runxfail -q -x --tb=short 2>&1 | tail -40 >> gpurun_out/r2_unverified_tests.log
  echo "=== $f exit ${PIPESTATUS[0]}" >> gpurun_out/r2_unverified_tests.log
done
# 2. microbenchmarks of the new routes (each guarded: a failure keeps the rest)
timeout 600 python tools/seqconv_bench.py --out gpurun_out/seqconv_bench.json > gpurun_out/seqconv_bench.log 2>&1
timeout 300 python tools/decode_bench.py --out gpurun_out/decode_bench.json > gpurun_out/decode_bench.log 2>&1
# 3. ncu launch lists (durations only; per-launch times are cold-cache and serialised)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'seq_|se_|conv_im2col|conv_col2im' -c 200 --csv \
  --log-file gpurun_out/seqconv_launches_r02.csv python tools/seqconv_bench.py --batch 16 --seconds 10 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:greedy_decode -c 1 -o gpurun_out/greedy_decode_r02 \
  python tools/decode_bench.py --batch 32 --frames 60 > /dev/null 2>&1
tail -5 gpurun_out/r2_unverified_tests.log
# 4. fused log-sum-exp (gemm feature class 8): the headline bench with and without it
TFASR_FUSED_LSE=1 timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fused_lse.json 2> gpurun_out/bench_fused_lse.err
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
# 5. second-generation conv1 kernels (csrc/conv1.cu) in the headline step
TFASR_CONV1_V2=1 timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_conv1_v2.json 2> gpurun_out/bench_conv1_v2.err
# 6. second-generation RNN-T gradient kernel, alone and together with the fused log-sum-exp and conv1 v2 (the candidate new default)
TFASR_RNNT_GRAD_V2=1 timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_rnnt_grad_v2.json 2> gpurun_out/bench_rnnt_grad_v2.err
TFASR_JOINT_V2=1 timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_joint_v2.json 2> gpurun_out/bench_joint_v2.err
TFASR_IM2COL_V2=1 timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_im2col_v2.json 2> gpurun_out/bench_im2col_v2.err
TFASR_DWCONV_WGRAD_V2=1 timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_dwconv_wgrad_v2.json 2> gpurun_out/bench_dwconv_wgrad_v2.err
TFASR_POS_BATCHED=1 timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pos_batched.json 2> gpurun_out/bench_pos_batched.err
TFASR_POS_BATCHED=1 TFASR_RNNT_GRAD_V2=1 TFASR_FUSED_LSE=1 TFASR_CONV1_V2=1 TFASR_JOINT_V2=1 TFASR_IM2COL_V2=1 TFASR_DWCONV_WGRAD_V2=1 timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_all_v2.json 2> gpurun_out/bench_all_v2.err
# 7. launch list of one step with every second-generation kernel on (compare shares with profiles/launch_summary_r01b.txt)
TFASR_POS_BATCHED=1 TFASR_RNNT_GRAD_V2=1 TFASR_FUSED_LSE=1 TFASR_CONV1_V2=1 TFASR_JOINT_V2=1 TFASR_IM2COL_V2=1 TFASR_DWCONV_WGRAD_V2=1 timeout 600 \
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02_all_v2.csv \
  python tools/profile_step.py --ncu > /dev/null 2>&1
python tools/summarize_launches.py gpurun_out/launches_r02_all_v2.csv > gpurun_out/launch_summary_r02_all_v2.txt 2>&1 || true
for f in gpurun_out/bench_*.json; do echo "$f: $(python -c "import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print(round(d['ms_per_step'],3), 'ms', round(d['value'],1), 'utt/s', 'rnnt frac', round((d.get('roofline') or {}).get('frac', 0), 3))" 2>/dev/null)"; done | tee gpurun_out/bench_variants.txt
