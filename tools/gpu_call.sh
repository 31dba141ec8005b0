#!/bin/bash
TAG=${1:-r09}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pikevm.py tests/test_gpu_dfa.py -m gpu -q -k "pikevm or nullable or pipelined or shards" > gpurun_out/${TAG}_pytest.log 2>&1
tail -25 gpurun_out/${TAG}_pytest.log | cut -c1-700
