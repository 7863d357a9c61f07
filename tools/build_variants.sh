#!/bin/bash
# development: build engine variants into build/ (NIQ_VARIANT bits, ptxas opt level) for tools/engine_probe.py
set -e
cd "$(dirname "$0")/.."
F="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo --fmad=false -std=c++17 -shared -Xcompiler -fPIC"
for spec in "$@"; do
  v=${spec%%:*}; o=${spec##*:}
  nvcc $F -Xptxas -O$o -DNIQ_VARIANT=$v -o build/libniq_v${v}_O${o}.so neural-implicit-queries_b200/csrc/niq_api.cu &
done
wait
