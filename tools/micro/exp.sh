mkdir -p gpurun_out
run() { # tiles defs
  echo "== tiles=$1 defs=$2"
  CGX_TILES=$1 CGX_JIT_DEFS="$2" AB_ARMS=jit timeout -k 10 120 python tools/ab_flat.py 16 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('   ', d['pattern'][:18], d['bitstream']['GBps'])
"
}
run 1 ""
run 1 "-DCGX_OWN_MASK=0"
run 1 "-DCGX_EMIT_MERGED=0"
run 1 "-DCGX_TAIL_NOINLINE=0"
run 1 "-DCGX_OWN_MASK=0 -DCGX_EMIT_MERGED=0 -DCGX_TAIL_NOINLINE=0 -DCGX_IDLE_NS=2000"
