for t in test_stat_split test_log_marginal test_suffixes_scores test_scorer_attributes test_scorer_asserts test_benchmark_shapes test_object_route test_segments_with_scores test_split_bedgraph test_slidingwindow_algorithm test_regularized test_not_constant_and test_reducers_vs test_constant_reducers; do
  python -m pytest tests/test_gpu_api.py -x -q -k "$t or deep" > gpurun_out/bis_$t.log 2>&1; echo "$t rc=$?"
done
