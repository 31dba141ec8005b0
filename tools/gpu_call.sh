#!/bin/bash
TAG=${1:-r11}
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_ref_patterns.py tests/test_cpp_header.py -m gpu -q --durations=3 > gpurun_out/${TAG}_pytest.log 2>&1
tail -30 gpurun_out/${TAG}_pytest.log | cut -c1-1500
