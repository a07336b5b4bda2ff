# Measurement batch of a round (run under gpurun): bench, per-round times, ncu launch list and full capture.
set -x
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err || exit 1
python tools/prune_stats.py 248956422 > gpurun_out/ps_chr1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_final2.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_l2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:window_dp -s 42 -c 10 -o gpurun_out/prof_wdp_final2 -f python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_f2.log 2>&1
tail -2 gpurun_out/ncu_f2.log
