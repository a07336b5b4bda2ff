#!/bin/bash
# round-2 closing batch on one GPU: all GPU tests, the default bench, smoke, then (after those exited 0) the ncu launch list and
# a full capture of the window / scan kernels of the same bench command; plus a host memory-bandwidth probe
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-f2}
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_gpu_tests.log 2>&1
echo "gpu tests rc=$?" >> gpurun_out/${T}_gpu_tests.log
timeout 1200 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
rc=$?
echo "bench rc=$rc" >> gpurun_out/${T}_bench.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${T}_smoke.log 2>&1
( g++ -O3 -std=c++17 -pthread tools/ubench/host_narrow.cpp -o /tmp/host_narrow && for t in 1 4 8 16; do /tmp/host_narrow $t | tail -1; done ) > gpurun_out/${T}_host_narrow.txt 2>&1
if [ $rc -eq 0 ]; then
  timeout 900 python bench.py --steps 2 --warmup 3 --genome-scale 0 --skip-exact > gpurun_out/${T}_plain_bench.json 2> gpurun_out/${T}_plain_bench.err && {
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${T}_launches_bench.csv \
      python bench.py --steps 2 --warmup 3 --genome-scale 0 --skip-exact > gpurun_out/${T}_ncu_launches.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"window_dp|scan_counts" -s 30 -c 12 -f -o gpurun_out/${T}_wdp_scan \
      python bench.py --steps 2 --warmup 3 --genome-scale 0 --skip-exact > gpurun_out/${T}_ncu_wdp.log 2>&1
  python tools/ncu_summary.py gpurun_out/${T}_wdp_scan.ncu-rep > gpurun_out/${T}_wdp_scan_summary.txt 2>&1
  }
fi
tail -4 gpurun_out/${T}_gpu_tests.log; tail -2 gpurun_out/${T}_bench.err; head -c 500 gpurun_out/${T}_bench.json; cat gpurun_out/${T}_smoke.log | tail -2; cat gpurun_out/${T}_host_narrow.txt; head -c 400 gpurun_out/${T}_bench_reference.json
timeout 1500 python tools/config5_stdin.py > gpurun_out/${T}_config5_stdin_text.json 2> gpurun_out/${T}_config5_stdin.err
timeout 900 python tools/cli_e2e.py > gpurun_out/${T}_cli_e2e.log 2>&1
cat gpurun_out/${T}_config5_stdin_text.json; tail -4 gpurun_out/${T}_cli_e2e.log
