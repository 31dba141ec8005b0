#!/bin/bash
# the call of the moment: whole GPU suite, four workers on the one GPU (the tests are host-bound)
TAG=${1:-r12}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -n 4 --durations=8 > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/${TAG}_pytest.log
tail -40 gpurun_out/${TAG}_pytest.log | cut -c1-400
