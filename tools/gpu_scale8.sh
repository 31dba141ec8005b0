#!/bin/bash
# 8-GPU headline bench as the driver launches it (only the NS line: 8x box time).
TAG=$1; N=${2:-8}
mkdir -p gpurun_out
timeout -k 10 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29527 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu > gpurun_out/${TAG}_bench_n${N}.json 2> gpurun_out/${TAG}_bench_n${N}.err
cut -c1-1200 gpurun_out/${TAG}_bench_n${N}.json; tail -3 gpurun_out/${TAG}_bench_n${N}.err | cut -c1-300
