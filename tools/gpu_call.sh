#!/bin/bash
# the command of one gpurun call of round 2 (kept in a file so that retries send the current tree)
TAG=$1
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_sim_flat.py tests/test_gpu_pike_search.py tests/test_config.py -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
tail -4 gpurun_out/${TAG}_pytest.log
bash tools/gpu_exp.sh ${TAG} "" "-DCGX_UNROLL=1" "-DCGX_UNROLL=9" "-DCGX_K=4 -DCGX_WARPS=20" "-DCGX_K=16 -DCGX_WARPS=8 -DCGX_UNROLL=1"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:cgx_flat_jit -s 4 -c 1 -f -o gpurun_out/${TAG}_scan_full python bench.py --steps 1 --passes 1 --warmup 3 --no-e2e --no-cpu --no-parity > gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out | tail -3
