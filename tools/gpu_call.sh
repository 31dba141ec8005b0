#!/bin/bash
TAG=$1
mkdir -p gpurun_out
AB_PATS=2 bash tools/gpu_exp.sh ${TAG} "" "-DCGX_FASTA=0" "-DCGX_FAST_UNROLL=2" ""
