#!/bin/bash
# window DP variant check: quick parity tests, per-round times, PROF build phase shares
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-s4}
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -m gpu -x -q -k "round or window or pipeline or fixture or chr or config" > gpurun_out/${T}_gpu_tests.log 2>&1
echo "gpu tests rc=$?" >> gpurun_out/${T}_gpu_tests.log
tail -3 gpurun_out/${T}_gpu_tests.log
timeout 300 python tools/prune_stats.py 248956422 > gpurun_out/${T}_rounds.txt 2>&1
cat gpurun_out/${T}_rounds.txt | tail -11
( cd pasio_b200/csrc && rm -f window_dp.o && make PROF=1 > /dev/null 2>&1 )
timeout 300 python tools/prune_stats.py 248956422 > gpurun_out/${T}_prof_rounds.txt 2>&1
grep -B2 "^round 9\|^round 3" gpurun_out/${T}_prof_rounds.txt
