#!/usr/bin/env bash
# Eight-GPU check of the round's final defaults: the headline bench at N = 8, started the way the driver starts it.   /usr/local/graft/bin/gpurun --gpus 8 --timeout 900 -- 'bash tools/gpu_n8.sh'
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 400 $RUN --master-port 29538 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err
echo "bench n8 exit $?" >> gpurun_out/bench_n8.err
cut -c1-1200 gpurun_out/bench_n8.json; tail -3 gpurun_out/bench_n8.err
