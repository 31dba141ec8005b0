#!/bin/bash
# the command of one gpurun call of round 2 (kept in a file so that retries send the current tree)
TAG=$1
mkdir -p gpurun_out
AB_STEPS=10 AB_PATS=5 timeout -k 10 300 python tools/ab_flat.py 16 > gpurun_out/${TAG}_ab.jsonl 2> gpurun_out/${TAG}_ab.err
cut -c1-300 gpurun_out/${TAG}_ab.jsonl; tail -3 gpurun_out/${TAG}_ab.err
TEST_TIMEOUT=1500 bash tools/gpu_round2.sh ${TAG} tests bench c5 configs ncu
