#!/bin/bash
# full GPU validation: all tests, bench (1 GPU), K3 timing
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-v1}
export PASIO_XD_PROF=1
for cfg in exact1 exact3; do
  timeout 300 python tools/workloads.py $cfg --reps 2 --prune 1 --lag 3 >> gpurun_out/${T}_exact_timing.jsonl 2>> gpurun_out/${T}_exact_prof.txt
done
unset PASIO_XD_PROF
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_gpu_tests.log 2>&1
echo "gpu tests rc=$?" >> gpurun_out/${T}_gpu_tests.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
echo "bench rc=$?" >> gpurun_out/${T}_bench.err
tail -5 gpurun_out/${T}_gpu_tests.log; tail -3 gpurun_out/${T}_bench.err; head -c 600 gpurun_out/${T}_bench.json
