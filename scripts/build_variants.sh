#!/bin/bash
# Build A/B variants of libxsb200.so (selected at run time with XSB200_GPU_LIB).
# usage: build_variants.sh NAME "-DFLAGS" [NAME "-DFLAGS" ...]   -> xsbench_b200/variants/libxsb200_NAME.so
set -e
cd "$(dirname "$0")/.."
mkdir -p xsbench_b200/variants
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC -Iinclude $flags \
       -Xptxas -v -shared -o xsbench_b200/variants/libxsb200_$name.so xsbench_b200/csrc/xs_gpu.cu -cudart static -ldl 2> /tmp/ptxas_$name.log || { cat /tmp/ptxas_$name.log; exit 1; }
  echo "== $name ($flags)"; grep -A2 "onesweep_pass_kernel" /tmp/ptxas_$name.log | grep -E "registers|spill" | head -3
done
