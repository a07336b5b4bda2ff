#!/bin/bash
# K3 instruction-cache experiment: parity tests + timing / phase cycles of configs 1 and 3
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-i1}
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "exact or config or pruned" > gpurun_out/${T}_gpu_tests.log 2>&1
echo "gpu tests rc=$?" >> gpurun_out/${T}_gpu_tests.log
tail -3 gpurun_out/${T}_gpu_tests.log
for cfg in exact1 exact3; do
  PASIO_XD_PROF=1 timeout 300 python tools/workloads.py $cfg --reps 3 >> gpurun_out/${T}_exact.jsonl 2>> gpurun_out/${T}_exact_prof.txt
done
python - <<PY
import json
for l in open('gpurun_out/${T}_exact.jsonl'):
    d = json.loads(l); print(d['workload'][:7], 'kernel %.2f ms' % d['kernel_ms'])
PY
tail -4 gpurun_out/${T}_exact_prof.txt
