#!/usr/bin/env bash
# Two-GPU check of the round's final defaults: the 2-rank peer-exchange test, the headline bench at N = 2 (both arms start the way the driver
# starts them).   /usr/local/graft/bin/gpurun --gpus 2 --timeout 900 -- 'bash tools/gpu_n2.sh'
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 python -m pytest tests/test_gpu_peer_multi.py -m gpu -q --tb=short 2>&1 | tail -5 > gpurun_out/peer_multi.log
timeout 400 $RUN --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "bench n2 exit $?" >> gpurun_out/bench_n2.err
tail -2 gpurun_out/peer_multi.log; cut -c1-1200 gpurun_out/bench_n2.json; tail -3 gpurun_out/bench_n2.err
