#!/bin/bash
# multi-GPU bench exactly as the driver launches it (torchrun, one rank per GPU over NCCL for the barrier / timing reduce)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}; T=${2:-m$N}
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
echo "bench rc=$?" >> gpurun_out/${T}_bench.err
tail -3 gpurun_out/${T}_bench.err; head -c 300 gpurun_out/${T}_bench.json
