#!/bin/bash
# the command of one gpurun call of round 2 (kept in a file so that retries send the current tree)
TAG=$1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_simd.py tests/test_gpu_dfa.py -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
tail -12 gpurun_out/${TAG}_pytest.log
