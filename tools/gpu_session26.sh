#!/bin/bash
# K3 variants x (lag, N blocks) on one box, with the per-step-type counters
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-i3}
PASIO_B200_LIB=$PWD/build_variants/libe1.so timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "exact or config or pruned" > gpurun_out/${T}_gpu_tests_e1.log 2>&1
echo "e1 gpu tests rc=$?" >> gpurun_out/${T}_gpu_tests_e1.log
tail -2 gpurun_out/${T}_gpu_tests_e1.log
for v in e0 e1; do
for geo in "3 1" "4 2"; do
set -- $geo
for cfg in exact1 exact3; do
  PASIO_B200_LIB=$PWD/build_variants/lib$v.so PASIO_XD_PROF=1 timeout 300 python tools/workloads.py $cfg --reps 3 --lag $1 --nblock $2 >> gpurun_out/${T}_exact_${v}_$1$2.jsonl 2>> gpurun_out/${T}_exact_prof_${v}_$1$2.txt
done
python - <<PY
import json
for l in open('gpurun_out/${T}_exact_${v}_$1$2.jsonl'):
    d = json.loads(l); print('$v lag $1 nblock $2', d['workload'][:7], 'kernel %.2f ms' % d['kernel_ms'], 'evaluated %.4f' % d['evaluated_frac'])
PY
tail -5 gpurun_out/${T}_exact_prof_${v}_$1$2.txt | grep "step type" | cut -c1-600
done
done
