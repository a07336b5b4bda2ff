#!/bin/bash
# Round-2 measurement batch: all GPU tests, bench (1 GPU), then -- only after those exited 0 without ncu -- the ncu launch
# list of the bench command and full captures of the kernels the judge asked for (K3 on configs 1 and 3, K1, window kernels).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-v2}
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_gpu_tests.log 2>&1
echo "gpu tests rc=$?" >> gpurun_out/${T}_gpu_tests.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
rc=$?
echo "bench rc=$rc" >> gpurun_out/${T}_bench.err
PASIO_WD_LGS=1 timeout 300 python tools/prune_stats.py 248956422 > gpurun_out/${T}_rounds_lgs1.txt 2>&1
PASIO_WD_LGS=0 timeout 300 python tools/prune_stats.py 248956422 > gpurun_out/${T}_rounds_lgs0.txt 2>&1
for cfg in exact1 exact3; do
  PASIO_XD_PROF=1 timeout 300 python tools/workloads.py $cfg --reps 2 >> gpurun_out/${T}_exact_timing.jsonl 2>> gpurun_out/${T}_exact_prof.txt
done
tail -4 gpurun_out/${T}_gpu_tests.log; tail -3 gpurun_out/${T}_bench.err; head -c 400 gpurun_out/${T}_bench.json
if [ "$2" = "ncu" ] && [ $rc -eq 0 ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:exact_pruned_kernel -c 1 -f -o gpurun_out/${T}_k3_config1 \
      python tools/workloads.py exact1 --reps 0 > gpurun_out/${T}_ncu_k3c1.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:exact_pruned_kernel -c 1 -f -o gpurun_out/${T}_k3_config3 \
      python tools/workloads.py exact3 --reps 0 > gpurun_out/${T}_ncu_k3c3.log 2>&1
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${T}_launches_bench.csv \
      python bench.py --steps 2 --warmup 3 --genome-scale 0 --skip-exact > gpurun_out/${T}_ncu_launches.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"window_dp|scan_counts" -s 30 -c 12 -f -o gpurun_out/${T}_wdp_scan \
      python bench.py --steps 2 --warmup 3 --genome-scale 0 --skip-exact > gpurun_out/${T}_ncu_wdp.log 2>&1
  for f in k3_config1 k3_config3 wdp_scan; do
    python tools/ncu_summary.py gpurun_out/${T}_${f}.ncu-rep > gpurun_out/${T}_${f}_summary.txt 2>&1
  done
  ls -la gpurun_out/${T}_*ncu-rep
fi
