#!/bin/bash
# N-GPU bench as the driver launches it.  Usage: gpurun --gpus N -- 'bash tools/gpu_scale.sh <tag> N'
TAG=$1; N=$2
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_n${N}.json 2> gpurun_out/${TAG}_bench_n${N}.err
cat gpurun_out/${TAG}_bench_n${N}.json | cut -c1-1500; tail -3 gpurun_out/${TAG}_bench_n${N}.err
timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --config c5 --gpus $N --steps 5 --warmup 3 > gpurun_out/${TAG}_c5_n${N}.json 2> gpurun_out/${TAG}_c5_n${N}.err
cat gpurun_out/${TAG}_c5_n${N}.json | cut -c1-1500; tail -3 gpurun_out/${TAG}_c5_n${N}.err
timeout -k 10 300 python -m pytest tests/test_dist_nccl.py -m gpu -x -q 2>&1 | tail -3
