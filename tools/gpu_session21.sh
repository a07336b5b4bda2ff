#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-h3}
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "narrowed or load_and_round" > gpurun_out/${T}_gpu_tests.log 2>&1
echo "gpu tests rc=$?" >> gpurun_out/${T}_gpu_tests.log
tail -4 gpurun_out/${T}_gpu_tests.log
for m in 1 0; do
PASIO_B200_UPLOAD_NARROW=$m timeout 900 python bench.py --genome-scale 0 --skip-exact > gpurun_out/${T}_bench_narrow$m.json 2> gpurun_out/${T}_bench_narrow$m.err
python - <<PY
import json
d = json.loads(open('gpurun_out/${T}_bench_narrow$m.json').read().strip().splitlines()[-1])
print('PASIO_B200_UPLOAD_NARROW=$m', {k: d[k] for k in ('value', 'ms_per_step') if k in d}, {k: d['e2e'][k] for k in ('value', 'ms_per_step', 'h2d_wire_bytes_per_step', 'ms_each_step', 'device_ms_per_step')})
PY
done
