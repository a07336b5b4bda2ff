#!/bin/bash
# K3 block sweeps: parity tests + timing with and without
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-s5}
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "exact or config or pruned" > gpurun_out/${T}_gpu_tests.log 2>&1
echo "gpu tests rc=$?" >> gpurun_out/${T}_gpu_tests.log
tail -3 gpurun_out/${T}_gpu_tests.log
for sw in 1 0; do
for cfg in exact1 exact3; do
  PASIO_XD_SWEEPS=$sw PASIO_XD_PROF=1 timeout 300 python tools/workloads.py $cfg --reps 3 >> gpurun_out/${T}_exact_sw$sw.jsonl 2>> gpurun_out/${T}_exact_prof_sw$sw.txt
done
done
cat gpurun_out/${T}_exact_sw1.jsonl gpurun_out/${T}_exact_sw0.jsonl | cut -c1-250
tail -8 gpurun_out/${T}_exact_prof_sw1.txt
