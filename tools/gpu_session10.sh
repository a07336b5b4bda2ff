#!/bin/bash
# new scan kernel + speculative window chain: tests, per-round times, bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-s1}
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_gpu_tests.log 2>&1
echo "gpu tests rc=$?" >> gpurun_out/${T}_gpu_tests.log
tail -4 gpurun_out/${T}_gpu_tests.log
PASIO_WD_SPECULATE=1 timeout 300 python tools/prune_stats.py 248956422 > gpurun_out/${T}_rounds_spec1.txt 2>&1
PASIO_WD_SPECULATE=0 timeout 300 python tools/prune_stats.py 248956422 > gpurun_out/${T}_rounds_spec0.txt 2>&1
cat gpurun_out/${T}_rounds_spec1.txt | tail -12
timeout 900 python bench.py --genome-scale 0 --skip-exact > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open('gpurun_out/${T}_bench.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'e2e', 'roofline', 'parity') if k in d})
print(d.get('kernel_ms_per_step'))
PY
