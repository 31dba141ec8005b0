#!/bin/bash
# the call of the moment: the bench line on the final tree
TAG=${1:-r17}
mkdir -p gpurun_out
timeout 100 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
cut -c1-260 gpurun_out/${TAG}_bench.json; tail -2 gpurun_out/${TAG}_bench.err | cut -c1-300
