#!/bin/bash
# all GPU tests, default bench, smoke
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-z4}
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_gpu_tests.log 2>&1
echo "gpu tests rc=$?" >> gpurun_out/${T}_gpu_tests.log
timeout 1200 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
echo "bench rc=$?" >> gpurun_out/${T}_bench.err
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${T}_smoke.log 2>&1
tail -3 gpurun_out/${T}_gpu_tests.log; tail -2 gpurun_out/${T}_bench.err; tail -1 gpurun_out/${T}_smoke.log
python - <<PY
import json
d = json.loads(open('gpurun_out/${T}_bench.json').read().strip().splitlines()[-1])
print('ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'frac', d['roofline']['frac'], 'avg_launch_ms', d['roofline']['avg_launch_ms'], 'parity', d['parity']['splits_equal'])
print({c: d['exact_dp'][c]['kernel_ms'] for c in d['exact_dp']})
PY
