#!/bin/bash
# GPU session 1 (round 2): latency micro-benchmarks, the new exact-DP kernel (parity first, then timing), full suite
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/s1_gpu.txt
timeout 60 tools/ubench/lat > gpurun_out/s1_ubench_lat.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "exact or config" > gpurun_out/s1_exact_tests.log 2>&1
echo "exact tests rc=$?" >> gpurun_out/s1_exact_tests.log
for cfg in exact1 exact3; do
  for opt in "--prune 1 --lag 3" "--prune 1 --lag 4" "--prune 0"; do
    timeout 300 python tools/workloads.py $cfg --reps 3 $opt >> gpurun_out/s1_exact_timing.jsonl 2>> gpurun_out/s1_exact_timing.err
  done
done
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s1_gpu_tests.log 2>&1
echo "gpu tests rc=$?" >> gpurun_out/s1_gpu_tests.log
tail -3 gpurun_out/s1_exact_tests.log; cat gpurun_out/s1_exact_timing.jsonl; tail -3 gpurun_out/s1_gpu_tests.log; cat gpurun_out/s1_ubench_lat.txt
