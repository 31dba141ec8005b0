# A/B of build-time variants of the specialised kernel (CGX_JIT_DEFS), 16 GiB IP scan + the other flat patterns
#   bash tools/micro/exp.sh "" "-DCGX_PAIR=1" "-DCGX_WARPS=13 -DCGX_CTAS=2"      (one line of GB/s per pattern and variant)
mkdir -p gpurun_out
run() { # defs
  echo "== defs=$1"
  CGX_JIT_DEFS="$1" AB_ARMS=jit timeout -k 10 120 python tools/ab_flat.py 16 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l); print('   ', d['pattern'][:18], d['bitstream']['GBps'], d['bitstream']['matches'])
    except Exception: print('   ?', l[:200].rstrip())
"
}
for v in "$@"; do run "$v"; done
