#!/bin/bash
TAG=${1:-r10}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_wrappers.py -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1
tail -40 gpurun_out/${TAG}_pytest.log | cut -c1-500
