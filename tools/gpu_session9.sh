#!/bin/bash
# K3 ncu captures (after a plain run exited 0) + the new tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-k1}
timeout 900 python -m pytest tests/test_gpu_api.py -m gpu -x -q -k "ends or mutated or refuses or regularized or logfac" > gpurun_out/${T}_new_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/${T}_new_tests.log
for c in 1 3; do
  timeout 300 python tools/workloads.py exact$c --reps 0 > gpurun_out/${T}_plain_exact$c.json 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:exact_pruned_kernel -c 1 -f -o gpurun_out/${T}_k3_config$c \
      python tools/workloads.py exact$c --reps 0 > gpurun_out/${T}_ncu_k3c$c.log 2>&1
  python tools/ncu_summary.py gpurun_out/${T}_k3_config$c.ncu-rep > gpurun_out/${T}_k3_config${c}_summary.txt 2>&1
done
tail -3 gpurun_out/${T}_new_tests.log; tail -3 gpurun_out/${T}_ncu_k3c1.log; ls -la gpurun_out/${T}_*ncu-rep
