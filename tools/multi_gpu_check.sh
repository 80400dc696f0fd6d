#!/usr/bin/env bash
# Round 2, multi-GPU call: the peer-memory sync-BatchNorm exchange (csrc/peer_allreduce.cu), written without GPU time in round 1.
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 900 -- 'bash tools/multi_gpu_check.sh 2'
set -u
N=${1:-2}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
# 1. one GPU: the protocol with a world of one rank (eager + graph replay)
timeout 300 python -m pytest tests/test_gpu_peer.py tests/test_gpu_peer_multi.py -m gpu -q --tb=short 2>&1 | tail -20 > gpurun_out/peer_single.log
# 2. N GPUs: correctness vs NCCL, identical sums on every rank, latency of 36 dependent exchanges in one graph
timeout 300 $RUN --master-port 29517 tools/peer_bn_check.py --out gpurun_out/peer_bn_check_n$N.json > gpurun_out/peer_bn_check_n$N.log 2>&1
echo "peer_bn_check exit $?" >> gpurun_out/peer_bn_check_n$N.log
# 3. the headline step at N GPUs, NCCL exchange vs peer-memory exchange
timeout 400 $RUN --master-port 29518 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n${N}_nccl_bn.json 2> gpurun_out/bench_n${N}_nccl_bn.err
TFASR_PEER_BN=1 timeout 400 $RUN --master-port 29519 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n${N}_peer_bn.json 2> gpurun_out/bench_n${N}_peer_bn.err
tail -3 gpurun_out/peer_single.log gpurun_out/peer_bn_check_n$N.log
for f in gpurun_out/bench_n${N}_*.json; do echo "$f: $(python -c "import json; d=json.loads(open('$f').read().strip().splitlines()[-1]); print(round(d['ms_per_step'],3), 'ms', round(d['value'],1), 'utt/s')" 2>/dev/null)"; done
