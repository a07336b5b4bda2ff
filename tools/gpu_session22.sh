#!/bin/bash
# K3 N tasks: parity tests + timing with and without
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-x1}
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "exact or config or pruned" > gpurun_out/${T}_gpu_tests.log 2>&1
echo "gpu tests rc=$?" >> gpurun_out/${T}_gpu_tests.log
tail -3 gpurun_out/${T}_gpu_tests.log
for nb in 1 0; do
for cfg in exact1 exact3; do
  PASIO_XD_PROF=1 timeout 300 python tools/workloads.py $cfg --reps 3 --nblock $nb >> gpurun_out/${T}_exact_nb$nb.jsonl 2>> gpurun_out/${T}_exact_prof_nb$nb.txt
done
done
cat gpurun_out/${T}_exact_nb1.jsonl gpurun_out/${T}_exact_nb0.jsonl | cut -c1-300
tail -4 gpurun_out/${T}_exact_prof_nb1.txt
