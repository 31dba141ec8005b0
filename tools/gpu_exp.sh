#!/bin/bash
# A/B of build-time variants of the specialised bitstream kernel: one line per pattern and variant.
#   bash tools/gpu_exp.sh <tag> "<defs 1>" "<defs 2>" ...
TAG=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  echo "== defs=$v"
  CGX_JIT_DEFS="$v" AB_STEPS=${AB_STEPS:-20} AB_ARMS=jit AB_PATS=${AB_PATS:-2} timeout -k 10 200 python tools/ab_flat.py 16 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l); print('   ', d['pattern'][:18], d['bitstream']['GBps'], d['bitstream']['matches'], d['bitstream'].get('serial_replays'))
    except Exception: print('   ?', l[:300].rstrip())
"
done 2>&1 | tee gpurun_out/${TAG}_exp.txt
