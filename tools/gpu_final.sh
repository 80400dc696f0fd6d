#!/usr/bin/env bash
# Round-end GPU box call (1 GPU): GPU suite, headline bench (full line), reference arm, smoke, launch list of one step, ncu captures of the
# RNN-T head kernels + log-mel and of the joint GEMM with row statistics.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
NCU="ncu --clock-control none"
timeout 900 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -30 > gpurun_out/gpu_tests.log
echo "=== pytest exit ${PIPESTATUS[0]}" >> gpurun_out/gpu_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
( time timeout 900 python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
for c in contextnet decode ds2; do timeout 400 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err; done
timeout 600 $NCU --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_r02c.csv python tools/profile_step.py --ncu > /dev/null 2>&1
python tools/summarize_launches.py gpurun_out/launches_r02c.csv > gpurun_out/launch_summary_r02c.txt 2>&1
timeout 600 $NCU --profile-from-start off --set full --import-source on -k regex:'rnnt_' -c 3 -o gpurun_out/head_final_r02 python tools/profile_step.py --ncu > gpurun_out/ncu_head.log 2>&1
timeout 600 $NCU --profile-from-start off --set full --kernel-name-base demangled -k regex:'gemm_bf16_kernel<\(int\)256, \(bool\)0, \(bool\)0, \(unsigned int\)8>' -c 1 -o gpurun_out/joint_gemm_final_r02 python tools/profile_step.py --ncu > gpurun_out/ncu_joint.log 2>&1
timeout 400 $NCU --set full --import-source on -k regex:rel_attention --launch-skip 3 -c 3 -o gpurun_out/attention_r02 python tools/att_probe.py > gpurun_out/ncu_att.log 2>&1
tail -4 gpurun_out/gpu_tests.log; cut -c1-2500 gpurun_out/bench_default.json; cut -c1-900 gpurun_out/bench_reference.json; tail -4 gpurun_out/bench_reference.err; tail -2 gpurun_out/smoke.log
for c in contextnet decode ds2; do cut -c1-400 gpurun_out/bench_$c.json; done
head -12 gpurun_out/launch_summary_r02c.txt; tail -2 gpurun_out/ncu_head.log; tail -2 gpurun_out/ncu_joint.log
