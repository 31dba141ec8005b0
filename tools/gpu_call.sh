#!/bin/bash
TAG=$1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sim_flat.py -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
tail -3 gpurun_out/${TAG}_pytest.log
AB_DENSE_GIB=6 AB_PATS=5 bash tools/gpu_exp.sh ${TAG} "" "-DCGX_RUN=1" "-DCGX_RUN=8"
