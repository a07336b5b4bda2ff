#!/bin/bash
# evidence runs of the other BASELINE configs on one GPU (every number DESIGN.md section 7 quotes has a file under profiles/)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-w1}
timeout 900 python tools/workloads.py genome > gpurun_out/${T}_config4_genome_1gpu.json 2> gpurun_out/${T}_genome.err
timeout 900 python tools/workloads.py transcripts > gpurun_out/${T}_config5_transcripts_1gpu.json 2> gpurun_out/${T}_transcripts.err
timeout 1500 python tools/config5_stdin.py > gpurun_out/${T}_config5_stdin_text.json 2> gpurun_out/${T}_config5_stdin.err
timeout 900 python tools/cli_e2e.py > gpurun_out/${T}_cli_e2e.log 2>&1
timeout 600 python tools/cli_stream.py 20000 > gpurun_out/${T}_cli_stream_20000.log 2>&1
cat gpurun_out/${T}_config4_genome_1gpu.json gpurun_out/${T}_config5_transcripts_1gpu.json gpurun_out/${T}_config5_stdin_text.json; tail -3 gpurun_out/${T}_cli_e2e.log gpurun_out/${T}_cli_stream_20000.log; tail -2 gpurun_out/${T}_*.err
