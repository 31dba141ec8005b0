#!/bin/bash
# the command of one gpurun call of round 2 (kept in a file so that retries send the current tree)
TAG=$1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sim_flat.py tests/test_gpu_pikevm.py -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
tail -4 gpurun_out/${TAG}_pytest.log
AB_PATS=2 bash tools/gpu_exp.sh ${TAG} "" "-DCGX_K=4 -DCGX_WARPS=19 -DCGX_CAP=160 -DCGX_UNROLL=5" "-DCGX_K=4 -DCGX_WARPS=23 -DCGX_CAP=160 -DCGX_UNROLL=5" "-DCGX_K=8 -DCGX_WARPS=15 -DCGX_UNROLL=9" "-DCGX_K=4 -DCGX_WARPS=19 -DCGX_CAP=320 -DCGX_UNROLL=5" "-DCGX_K=4 -DCGX_WARPS=19 -DCGX_CAP=128 -DCGX_UNROLL=5"
timeout 600 python tools/run_configs.py > gpurun_out/${TAG}_configs.jsonl 2> gpurun_out/${TAG}_configs.err
cut -c1-330 gpurun_out/${TAG}_configs.jsonl; tail -3 gpurun_out/${TAG}_configs.err
ls -la gpurun_out | tail -3
