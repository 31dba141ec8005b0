#!/bin/bash
# One gpurun call's worth of evidence (round 2).  Usage:
#   gpurun --timeout 1500 -- 'bash tools/gpu_round2.sh <tag> [steps...]'
# steps: tests large exp bench c5 ncu (default: all but ncu).  Output: gpurun_out/<tag>_*
TAG=${1:-r02}; shift
STEPS=${@:-tests exp bench c5}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
has() { [[ " $STEPS " == *" $1 "* ]]; }

if has tests; then
  echo "== pytest -m gpu"
  timeout -k 10 ${TEST_TIMEOUT:-900} python -m pytest tests -m gpu -x -q ${PYTEST_ARGS} > $OUT/${TAG}_pytest.log 2>&1
  echo "pytest exit $?" | tee -a $OUT/${TAG}_pytest.log
  tail -5 $OUT/${TAG}_pytest.log
fi
if has exp; then
  echo "== experiments"
  bash tools/micro/exp.sh ${EXP_VARIANTS:-"" "-DCGX_PAIR=1"} > $OUT/${TAG}_exp.txt 2>&1
  cat $OUT/${TAG}_exp.txt
fi
if has bench; then
  echo "== bench"
  timeout -k 10 600 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
  cat $OUT/${TAG}_bench.json; tail -3 $OUT/${TAG}_bench.err
  timeout -k 10 300 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/${TAG}_bench_ref.json 2>> $OUT/${TAG}_bench.err
  cat $OUT/${TAG}_bench_ref.json
fi
if has c5; then
  echo "== bench --config c5 (one shard)"
  timeout -k 10 600 python bench.py --config c5 --steps 5 --warmup 3 > $OUT/${TAG}_c5.json 2> $OUT/${TAG}_c5.err
  cat $OUT/${TAG}_c5.json; tail -3 $OUT/${TAG}_c5.err
fi
if has configs; then
  echo "== other configs"
  timeout -k 10 900 python tools/run_configs.py > $OUT/${TAG}_configs.jsonl 2> $OUT/${TAG}_configs.err
  cat $OUT/${TAG}_configs.jsonl; tail -3 $OUT/${TAG}_configs.err
fi
if has ncu; then
  echo "== ncu launch list (bench command)"
  timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $OUT/${TAG}_launches.csv python bench.py --steps 1 --passes 2 --warmup 3 --no-e2e --no-cpu --no-parity > $OUT/${TAG}_ncu_bench.log 2>&1
  echo "== ncu --set full (dominant kernel, 16 GiB)"
  timeout -k 10 400 ncu --set full --clock-control none --import-source on -k regex:'cgx_flat_jit|scan_flat_kernel|scan_dfa_kernel|cgx_' \
    -s 4 -c 1 -f -o $OUT/${TAG}_scan_full python bench.py --steps 1 --passes 1 --warmup 3 --no-e2e --no-cpu --no-parity > $OUT/${TAG}_ncu_full.log 2>&1
fi
ls -la $OUT | tail -20
