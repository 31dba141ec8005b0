#!/bin/bash
# One call's worth of confirmation on a B200 (gpurun --timeout 1500 -- 'bash tools/gpu_call.sh <tag>'):
# the whole GPU suite, smoke, the bench line, the reference arm, config 5 and the other configs.
TAG=${1:-r02x}
mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
TEST_TIMEOUT=1500 bash tools/gpu_round2.sh $TAG tests bench c5 configs
