#!/bin/bash
# window DP restructure: tests, per-round times, PROF build phase shares
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-s3}
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -m gpu -x -q > gpurun_out/${T}_gpu_tests.log 2>&1
echo "gpu tests rc=$?" >> gpurun_out/${T}_gpu_tests.log
tail -3 gpurun_out/${T}_gpu_tests.log
for spec in 1 0; do
PASIO_WD_SPECULATE=$spec timeout 300 python tools/prune_stats.py 248956422 > gpurun_out/${T}_rounds_spec$spec.txt 2>&1
done
cat gpurun_out/${T}_rounds_spec1.txt | tail -11
grep "round 9\|round 2:" gpurun_out/${T}_rounds_spec0.txt
( cd pasio_b200/csrc && rm -f window_dp.o && make PROF=1 > /dev/null 2>&1 )
PASIO_WD_SPECULATE=1 timeout 300 python tools/prune_stats.py 248956422 > gpurun_out/${T}_prof_rounds_spec1.txt 2>&1
grep -B2 "^round 9\|^round 3" gpurun_out/${T}_prof_rounds_spec1.txt
