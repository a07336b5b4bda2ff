#!/bin/bash
# K3 timing experiments: which part of the diagonal CTA's step costs what (results are wrong with PASIO_XD_DBG set)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-d1}
export PASIO_XD_PROF=1
for dbg in 0 16 32 1; do
  echo "== PASIO_XD_DBG=$dbg" >> gpurun_out/${T}_dbg.txt
  PASIO_XD_DBG=$dbg timeout 120 python tools/workloads.py exact1 --reps 1 2>&1 | grep -E "xp_prof|kernel_ms" | sed -n '3,5p' | cut -c1-400 >> gpurun_out/${T}_dbg.txt
done
cat gpurun_out/${T}_dbg.txt
