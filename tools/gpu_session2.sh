#!/bin/bash
# K3 phase profile + exact parity tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-s4}
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "exact or config" > gpurun_out/${T}_exact_tests.log 2>&1
echo "exact tests rc=$?" >> gpurun_out/${T}_exact_tests.log
export PASIO_XD_PROF=1
for cfg in exact1 exact3; do
  for opt in "--prune 1 --lag 3" "--prune 1 --lag 4" "--prune 1 --lag 3 --ring 0"; do
    timeout 300 python tools/workloads.py $cfg --reps 2 $opt >> gpurun_out/${T}_exact_timing.jsonl 2>> gpurun_out/${T}_exact_prof.txt
  done
done
grep -h "E  \|rc=" gpurun_out/${T}_exact_tests.log | head; tail -2 gpurun_out/${T}_exact_tests.log
