#!/bin/bash
# chr1-scale parity against the OpenMP oracle + the ncu evidence of the round (after plain runs exited 0)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-e1}
nproc > gpurun_out/${T}_nproc.txt
timeout 1500 python tools/parity_chr1.py > gpurun_out/${T}_parity_chr1.json 2> gpurun_out/${T}_parity_chr1.err
echo "parity rc=$?" >> gpurun_out/${T}_parity_chr1.err
timeout 300 python tools/workloads.py exact1 --reps 1 > gpurun_out/${T}_plain_exact1.json 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:exact_pruned_kernel -c 1 -f -o gpurun_out/${T}_k3_config1 \
    python tools/workloads.py exact1 --reps 0 > gpurun_out/${T}_ncu_k3c1.log 2>&1
timeout 300 python tools/workloads.py exact3 --reps 1 > gpurun_out/${T}_plain_exact3.json 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:exact_pruned_kernel -c 1 -f -o gpurun_out/${T}_k3_config3 \
    python tools/workloads.py exact3 --reps 0 > gpurun_out/${T}_ncu_k3c3.log 2>&1
timeout 900 python bench.py --steps 2 --warmup 3 --genome-scale 0 --skip-exact > gpurun_out/${T}_plain_bench.json 2> gpurun_out/${T}_plain_bench.err && {
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${T}_launches_bench.csv \
      python bench.py --steps 2 --warmup 3 --genome-scale 0 --skip-exact > gpurun_out/${T}_ncu_launches.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"window_dp|scan_counts" -s 30 -c 12 -f -o gpurun_out/${T}_wdp_scan \
      python bench.py --steps 2 --warmup 3 --genome-scale 0 --skip-exact > gpurun_out/${T}_ncu_wdp.log 2>&1
}
for f in k3_config1 k3_config3 wdp_scan; do
  python tools/ncu_summary.py gpurun_out/${T}_${f}.ncu-rep > gpurun_out/${T}_${f}_summary.txt 2>&1
done
tail -2 gpurun_out/${T}_parity_chr1.err; head -c 600 gpurun_out/${T}_parity_chr1.json; ls -la gpurun_out/${T}_*ncu-rep
