#!/bin/bash
TAG=$1
mkdir -p gpurun_out
python tools/dbg_teddy.py 2>&1 | tail -8
timeout 900 python -m pytest tests/test_sim_teddy.py tests/test_gpu_teddy.py tests/test_gpu_large.py -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
tail -4 gpurun_out/${TAG}_pytest.log
CFG_ONLY=C3,C5 timeout 600 python tools/run_configs.py > gpurun_out/${TAG}_configs.jsonl 2> gpurun_out/${TAG}_configs.err
cut -c1-330 gpurun_out/${TAG}_configs.jsonl; tail -3 gpurun_out/${TAG}_configs.err
CFG_ONLY=C3 CFG_SCALE=0.25 timeout 300 ncu --set full --clock-control none --import-source on -k regex:scan_dfa_kernel -s 3 -c 1 -f -o gpurun_out/${TAG}_teddy_full python tools/run_configs.py > gpurun_out/${TAG}_ncu_teddy.log 2>&1
