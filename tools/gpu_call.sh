#!/bin/bash
# the command of one gpurun call of round 2 (kept in a file so that retries send the current tree)
TAG=$1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_large.py tests/test_sim_teddy.py tests/test_gpu_teddy.py -m gpu -x -q -k "not 8gb and not 4gb and not beyond_4gib" > gpurun_out/${TAG}_pytest.log 2>&1
tail -6 gpurun_out/${TAG}_pytest.log
CFG_ONLY=C3,C5 timeout 600 python tools/run_configs.py > gpurun_out/${TAG}_configs.jsonl 2> gpurun_out/${TAG}_configs.err
cut -c1-330 gpurun_out/${TAG}_configs.jsonl; tail -3 gpurun_out/${TAG}_configs.err
timeout -k 10 500 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
ls -la gpurun_out | tail -3
