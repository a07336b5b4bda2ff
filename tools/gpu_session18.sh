#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-g1}
timeout 1200 python -m pytest tests/test_gpu_api.py tests/test_gpu_parity.py -m gpu -x -q -k "logfac or load_and_round or pipeline or lmm or bedgraph" > gpurun_out/${T}_gpu_tests.log 2>&1
echo "gpu tests rc=$?" >> gpurun_out/${T}_gpu_tests.log
tail -5 gpurun_out/${T}_gpu_tests.log
timeout 900 python bench.py --skip-exact > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open('gpurun_out/${T}_bench.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'parity') if k in d}, {k: d['e2e'][k] for k in ('value', 'ms_per_step', 'h2d_wire_bytes_per_step', 'ms_each_step', 'device_ms_per_step')})
g = d['genome']; print('genome', g['value'], g['ms_per_step'], g['e2e'])
PY
