#!/usr/bin/env bash
# 8-GPU check of the headline step: peer-memory sync-BN exchange (default) vs NCCL.   gpurun --gpus 8 --timeout 600 -- 'bash tools/scale8.sh'
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 240 $RUN --master-port 29520 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_n8_peer_bn.json 2> gpurun_out/bench_n8_peer_bn.err
TFASR_PEER_BN=0 timeout 240 $RUN --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_n8_nccl_bn.json 2> gpurun_out/bench_n8_nccl_bn.err
timeout 120 $RUN --master-port 29522 tools/peer_bn_check.py --out gpurun_out/peer_bn_check_n8.json > gpurun_out/peer_bn_check_n8.log 2>&1
for f in gpurun_out/bench_n8_*.json; do echo "$f: $(python -c "import json; d=json.loads(open('$f').read().strip().splitlines()[-1]); print(round(d['ms_per_step'],3), 'ms', round(d['value'],1), 'utt/s', d['config'].get('sync_bn_exchange'))" 2>/dev/null)"; done
cat gpurun_out/peer_bn_check_n8.json
