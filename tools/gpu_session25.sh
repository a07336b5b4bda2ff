#!/bin/bash
# K3 variants (tools/build_variants.sh) on one box: parity tests on the candidates, timing / phase cycles of configs 1 and 3
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-i2}
for v in ${PARITY_VARIANTS:-v3 v4}; do
  PASIO_B200_LIB=$PWD/build_variants/lib$v.so timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "exact or config or pruned" > gpurun_out/${T}_gpu_tests_$v.log 2>&1
  echo "$v gpu tests rc=$?" >> gpurun_out/${T}_gpu_tests_$v.log
  tail -2 gpurun_out/${T}_gpu_tests_$v.log
done
for v in ${VARIANTS:-v0 v1 v2 v3 v4}; do
for cfg in exact1 exact3; do
  PASIO_B200_LIB=$PWD/build_variants/lib$v.so PASIO_XD_PROF=1 timeout 300 python tools/workloads.py $cfg --reps 3 >> gpurun_out/${T}_exact_$v.jsonl 2>> gpurun_out/${T}_exact_prof_$v.txt
done
python - <<PY
import json
for l in open('gpurun_out/${T}_exact_$v.jsonl'):
    d = json.loads(l); print('$v', d['workload'][:7], 'kernel %.2f ms' % d['kernel_ms'])
PY
tail -4 gpurun_out/${T}_exact_prof_$v.txt | head -1 | cut -c1-400
done
