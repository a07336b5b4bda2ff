#!/bin/bash
# K3: merged far results + bulk copies; parity tests, then timing / per-step-type cycles for (lag, N blocks) = (3,1), (4,2)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-i6}
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "exact or config or pruned" > gpurun_out/${T}_gpu_tests.log 2>&1
echo "gpu tests rc=$?" >> gpurun_out/${T}_gpu_tests.log
tail -2 gpurun_out/${T}_gpu_tests.log
for geo in "5 3"; do
set -- $geo
for prof in 1 0; do
for cfg in exact1 exact3; do
  if [ $prof = 1 ]; then export PASIO_XD_PROF=1; else unset PASIO_XD_PROF; fi
  timeout 300 python tools/workloads.py $cfg --reps 3 --lag $1 --nblock $2 >> gpurun_out/${T}_exact_$1$2_prof$prof.jsonl 2>> gpurun_out/${T}_exact_prof_$1$2.txt
done
python - <<PY
import json
for l in open('gpurun_out/${T}_exact_$1$2_prof$prof.jsonl'):
    d = json.loads(l); print('lag $1 nblock $2 prof $prof', d['workload'][:7], 'kernel %.2f ms' % d['kernel_ms'], 'evaluated %.4f' % d['evaluated_frac'])
PY
done
tail -5 gpurun_out/${T}_exact_prof_$1$2.txt | grep "step type" | cut -c1-500
done
