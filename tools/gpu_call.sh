#!/bin/bash
# the command of one gpurun call of round 2 (kept in a file so that retries send the current tree)
TAG=$1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sim_flat.py tests/test_gpu_teddy.py -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
tail -4 gpurun_out/${TAG}_pytest.log
AB_STEPS=10 AB_PATS=5 timeout -k 10 300 python tools/ab_flat.py 16 > gpurun_out/${TAG}_ab.jsonl 2> gpurun_out/${TAG}_ab.err
cut -c1-400 gpurun_out/${TAG}_ab.jsonl; tail -3 gpurun_out/${TAG}_ab.err
bash tools/gpu_round2.sh ${TAG} bench c5 ncu
