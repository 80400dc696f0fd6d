#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for args in "1 250 1 64" "1 100 1 64" "1 250 1 36" "1 250 2 36" "2 19 4 36" "1 750 8 64 16 64" "1 130 1 32"; do
  echo "=== $args" >> gpurun_out/att_debug.log
  timeout 120 python tools/att_debug.py $args 2>&1 | tail -3 >> gpurun_out/att_debug.log
done
echo "=== sanitizer 1 250 2 36" >> gpurun_out/att_debug.log
timeout 300 compute-sanitizer --tool memcheck python tools/att_debug.py 1 250 2 36 2>&1 | grep -v "^$" | head -60 >> gpurun_out/att_debug.log
echo "=== sanitizer 1 100 1 64" >> gpurun_out/att_debug.log
timeout 300 compute-sanitizer --tool memcheck python tools/att_debug.py 1 100 1 64 2>&1 | grep -v "^$" | head -40 >> gpurun_out/att_debug.log
cat gpurun_out/att_debug.log | cut -c1-220
