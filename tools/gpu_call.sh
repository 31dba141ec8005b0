#!/bin/bash
# the call of the moment: captures kernel forms + pipelined host submatch
TAG=${1:-r06}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pikevm.py -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1
tail -25 gpurun_out/${TAG}_pytest.log | cut -c1-600
