#!/bin/bash
TAG=${1:-r16}
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_teddy.py -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1
tail -8 gpurun_out/${TAG}_pytest.log | cut -c1-600
