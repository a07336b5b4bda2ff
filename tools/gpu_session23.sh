#!/bin/bash
# closing validation of round 2: all GPU tests, default bench, smoke; then (plain runs first) the K3 ncu captures of configs 1 and 3
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-z1}
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_gpu_tests.log 2>&1
echo "gpu tests rc=$?" >> gpurun_out/${T}_gpu_tests.log
timeout 1200 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
echo "bench rc=$?" >> gpurun_out/${T}_bench.err
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${T}_smoke.log 2>&1
for c in 1 3; do
  timeout 300 python tools/workloads.py exact$c --reps 0 > gpurun_out/${T}_plain_exact$c.json 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:exact_pruned_kernel -c 1 -f -o gpurun_out/${T}_k3_config$c \
      python tools/workloads.py exact$c --reps 0 > gpurun_out/${T}_ncu_k3c$c.log 2>&1
  python tools/ncu_summary.py gpurun_out/${T}_k3_config$c.ncu-rep > gpurun_out/${T}_k3_config${c}_summary.txt 2>&1
done
tail -4 gpurun_out/${T}_gpu_tests.log; tail -2 gpurun_out/${T}_bench.err; head -c 400 gpurun_out/${T}_bench.json; tail -2 gpurun_out/${T}_smoke.log; head -3 gpurun_out/${T}_k3_config1_summary.txt | cut -c1-200
