#!/bin/bash
# K3 variants, timing only (no counters)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-k3}
for rep in 1 2; do
for v in ${VARIANTS:-c0 c1}; do
for cfg in exact1 exact3; do
  PASIO_B200_LIB=$PWD/build_variants/lib$v.so timeout 300 python tools/workloads.py $cfg --reps 5 >> gpurun_out/${T}_exact_${v}.jsonl 2>/dev/null
done
done
done
python - <<PY
import json
for v in '${VARIANTS:-c0 c1}'.split():
    for l in open('gpurun_out/${T}_exact_%s.jsonl' % v):
        d = json.loads(l); print(v, d['workload'][:7], 'kernel %.3f ms' % d['kernel_ms'])
PY
