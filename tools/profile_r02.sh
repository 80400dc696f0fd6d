#!/usr/bin/env bash
# Round-2 ncu evidence (1 GPU): launch list of one eager step + --set full captures of the kernels this round wrote or changed.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/profile_r02.sh'      (reports come back in gpurun_out/, summaries go to profiles/)
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
NCU="ncu --clock-control none"
timeout 600 $NCU --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_r02.csv python tools/profile_step.py --ncu > /dev/null 2>&1
python tools/summarize_launches.py gpurun_out/launches_r02.csv > gpurun_out/launch_summary_r02.txt 2>&1
timeout 400 $NCU --set full --import-source on -k regex:ffn_kernel -s 4 -c 2 -o gpurun_out/ffn_r02 python tools/ffn_trace.py > gpurun_out/ncu_ffn.log 2>&1
timeout 400 $NCU --set full --import-source on -k regex:rel_attention_fwd_tc -s 0 -c 1 -o gpurun_out/att_tc_r02 python tools/att_debug.py 32 250 4 36 > gpurun_out/ncu_att.log 2>&1
timeout 600 $NCU --profile-from-start off --set full --import-source on -k regex:'rnnt_|logmel' -c 4 -o gpurun_out/rnnt_logmel_r02 python tools/profile_step.py --ncu > gpurun_out/ncu_rnnt.log 2>&1
timeout 600 $NCU --profile-from-start off --set full -k regex:'gemm_bf16_kernel<256, 0, 0, 8' -c 1 -o gpurun_out/joint_gemm_r02 python tools/profile_step.py --ncu > gpurun_out/ncu_joint.log 2>&1
ls -la gpurun_out; head -30 gpurun_out/launch_summary_r02.txt
