#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-n2}
for m in 1 0; do
PASIO_B200_UPLOAD_NARROW=$m timeout 900 python bench.py --genome-scale 0 --skip-exact > gpurun_out/${T}_bench_narrow$m.json 2> gpurun_out/${T}_bench_narrow$m.err
echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open('gpurun_out/${T}_bench_narrow$m.json').read().strip().splitlines()[-1])
print('PASIO_B200_UPLOAD_NARROW=$m', {k: d[k] for k in ('value', 'ms_per_step') if k in d}, {k: d['e2e'][k] for k in ('value', 'ms_per_step', 'h2d_wire_bytes_per_step', 'ms_each_step', 'device_ms_per_step')})
PY
done
