#!/bin/bash
# scan with prefetch: tests + time; then a PROF build of the window kernel: phase shares per round
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-s2}
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -m gpu -x -q > gpurun_out/${T}_gpu_tests.log 2>&1
echo "gpu tests rc=$?" >> gpurun_out/${T}_gpu_tests.log
tail -3 gpurun_out/${T}_gpu_tests.log
timeout 900 python bench.py --genome-scale 0 --skip-exact > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open('gpurun_out/${T}_bench.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'parity') if k in d})
print(d.get('kernel_ms_per_step')); print(d['roofline']['scan_kernel'])
PY
( cd pasio_b200/csrc && rm -f window_dp.o && make PROF=1 > /dev/null 2>&1 )
for spec in 1 0; do
PASIO_WD_SPECULATE=$spec timeout 300 python tools/prune_stats.py 248956422 > gpurun_out/${T}_prof_rounds_spec$spec.txt 2>&1
done
grep -c wd_prof gpurun_out/${T}_prof_rounds_spec1.txt
tail -30 gpurun_out/${T}_prof_rounds_spec1.txt
