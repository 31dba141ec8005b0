#!/bin/bash
# Short iteration call: pipe micro-benchmark, A/B of the flat kernels (identity + oracle window
# check at full size), one ncu --set full capture of the specialised kernel.
#   gpurun --timeout 600 -- 'bash tools/gpu_quick.sh <tag>'
TAG=${1:-q}
OUT=gpurun_out
mkdir -p $OUT
if [ -n "$PIPES" ] && [ -x tools/micro/pipes.bin ]; then timeout 120 tools/micro/pipes.bin > $OUT/${TAG}_pipes.txt 2>&1; cat $OUT/${TAG}_pipes.txt; fi
if [ -n "$TILES2" ]; then echo "== CGX_TILES=2"; CGX_TILES=2 AB_ARMS=jit timeout -k 10 200 python tools/ab_flat.py 16 2>/dev/null | cut -c1-260; fi
timeout -k 10 300 python tools/ab_flat.py 16 > $OUT/${TAG}_ab.jsonl 2> $OUT/${TAG}_ab.err
cat $OUT/${TAG}_ab.jsonl; tail -3 $OUT/${TAG}_ab.err
timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:'cgx_flat_jit' \
  -s 3 -c 1 -f -o $OUT/${TAG}_scan_full python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > $OUT/${TAG}_ncu_full.log 2>&1
tail -2 $OUT/${TAG}_ncu_full.log | cut -c1-300
