#!/usr/bin/env bash
# Four-GPU check of the round's final defaults: the headline bench at N = 4, started the way the driver starts it.   /usr/local/graft/bin/gpurun --gpus 4 --timeout 900 -- 'bash tools/gpu_n4.sh'
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 400 $RUN --master-port 29534 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err
echo "bench n4 exit $?" >> gpurun_out/bench_n4.err
cut -c1-1200 gpurun_out/bench_n4.json; tail -3 gpurun_out/bench_n4.err
