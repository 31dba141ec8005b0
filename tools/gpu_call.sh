#!/bin/bash
# the command of one gpurun call of round 2 (kept in a file so that retries send the current tree)
TAG=$1
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_sim_flat.py tests/test_gpu_dfa.py -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
tail -4 gpurun_out/${TAG}_pytest.log
AB_PATS=4 timeout -k 10 300 python tools/ab_flat.py 16 > gpurun_out/${TAG}_ab.jsonl 2> gpurun_out/${TAG}_ab.err
cut -c1-400 gpurun_out/${TAG}_ab.jsonl; tail -3 gpurun_out/${TAG}_ab.err
AB_PATS=2 bash tools/gpu_exp.sh ${TAG} "" "-DCGX_ROT=0" "-DCGX_WARPS=15" "-DCGX_WARPS=16" "-DCGX_K=4 -DCGX_WARPS=24 -DCGX_CAP=160 -DCGX_UNROLL=1" "-DCGX_K=4 -DCGX_WARPS=20 -DCGX_CAP=160" "-DCGX_EXP_NOLB"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:cgx_flat_jit -s 4 -c 1 -f -o gpurun_out/${TAG}_scan_full python bench.py --steps 1 --passes 1 --warmup 3 --no-e2e --no-cpu --no-parity > gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out | tail -3
