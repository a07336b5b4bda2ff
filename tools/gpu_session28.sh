#!/bin/bash
# K3 variants (tools/build_variants.sh) on one box: timing / per-step-type cycles of configs 1 and 3
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-j1}
for v in ${VARIANTS:-p0 p1}; do
for prof in 1 0; do
for cfg in exact1 exact3; do
  if [ $prof = 1 ]; then export PASIO_XD_PROF=1; else unset PASIO_XD_PROF; fi
  PASIO_B200_LIB=$PWD/build_variants/lib$v.so timeout 300 python tools/workloads.py $cfg --reps 3 $WL_ARGS >> gpurun_out/${T}_exact_${v}_prof$prof.jsonl 2>> gpurun_out/${T}_exact_prof_$v.txt
done
python - <<PY
import json
for l in open('gpurun_out/${T}_exact_${v}_prof$prof.jsonl'):
    d = json.loads(l); print('$v prof $prof', d['workload'][:7], 'kernel %.2f ms' % d['kernel_ms'], 'evaluated %.4f' % d['evaluated_frac'])
PY
done
tail -5 gpurun_out/${T}_exact_prof_$v.txt | grep "step type" | cut -c1-420
done
