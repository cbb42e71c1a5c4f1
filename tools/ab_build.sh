#!/bin/bash
# Build kernel variants for A/B timing: tools/ab_build.sh name "-DPTK_MIN_BLOCKS=8 ..."
set -e
cd "$(dirname "$0")/.."
mkdir -p build/ab
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -Xptxas -v $2 \
  -o build/ab/libptcuda_$1.so pathtracer_ocl_b200/csrc/ptcuda.cu 2>&1 | grep -E "error|trace_kernelIf|registers|spill" | grep -A2 "IfLi0" | head -4
