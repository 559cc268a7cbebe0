#!/bin/sh
# Builds the C-ABI kernel library in-tree (sm_100a only).
set -e
cd "$(dirname "$0")"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
  -Xcompiler -fPIC,-fvisibility=hidden -shared -Iinclude "$@" \
  torchdistx_b200/csrc/kernels/tdx_init_kernels.cu -o torchdistx_b200/libtdx_init.so
