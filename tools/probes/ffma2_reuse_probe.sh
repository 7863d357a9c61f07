#!/bin/bash
# build (works without a GPU) and, with "run", execute on the GPU box.  Output: build/probes/
set -e
cd "$(dirname "$0")"
OUT=../../build/probes; mkdir -p $OUT
F="-gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -std=c++17 -lineinfo"
nvcc $F -Xptxas -O1 -DPSFX=O1 -cubin -o $OUT/ffma2_kernels_O1.cubin ffma2_kernels.cu
nvcc $F -Xptxas -O3 -DPSFX=O3 -cubin -o $OUT/ffma2_kernels_O3.cubin ffma2_kernels.cu
nvcc $F -o $OUT/ffma2_reuse_probe ffma2_reuse_probe.cu -lcuda
if [ "$1" = "run" ]; then $OUT/ffma2_reuse_probe $OUT; fi
