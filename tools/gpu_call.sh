#!/bin/bash
# the command of one gpurun call of round 2 (kept in a file so that retries send the current tree)
TAG=$1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sim_teddy.py tests/test_sim_flat.py tests/test_gpu_teddy.py -m gpu -x -q -k "not 8gb and not 4gb" > gpurun_out/${TAG}_pytest.log 2>&1
tail -4 gpurun_out/${TAG}_pytest.log
AB_PATS=1 bash tools/gpu_exp.sh ${TAG} "" "-DCGX_RBITS=0" "" "-DCGX_RBITS=0"
AB_STEPS=10 AB_PATS=5 timeout -k 10 300 python tools/ab_flat.py 1 > gpurun_out/${TAG}_ab.jsonl 2> gpurun_out/${TAG}_ab.err
cut -c1-300 gpurun_out/${TAG}_ab.jsonl; tail -3 gpurun_out/${TAG}_ab.err
ls -la gpurun_out | tail -3
