#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-h1}
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -m gpu -x -q -k "load or batch or config or devices" > gpurun_out/${T}_gpu_tests.log 2>&1
echo "gpu tests rc=$?" >> gpurun_out/${T}_gpu_tests.log
tail -4 gpurun_out/${T}_gpu_tests.log
timeout 900 python tools/workloads.py transcripts > gpurun_out/${T}_config5_transcripts_1gpu.json 2> gpurun_out/${T}_transcripts.err
timeout 900 python tools/workloads.py genome > gpurun_out/${T}_config4_genome_1gpu.json 2> gpurun_out/${T}_genome.err
cat gpurun_out/${T}_config5_transcripts_1gpu.json gpurun_out/${T}_config4_genome_1gpu.json
