#!/bin/sh
# Builds A/B variants of libtdx_init.so (knobs of the table kernel) + the C++ harness.
# usage: benchmarks/build_variants.sh "tag:-Dflags" ...
set -e
cd "$(dirname "$0")/.."
mkdir -p benchmarks/_variants
for spec in "$@"; do
  tag=${spec%%:*}; flags=${spec#*:}
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
    -Xcompiler -fPIC,-fvisibility=hidden -shared -Iinclude $flags \
    torchdistx_b200/csrc/kernels/tdx_init_kernels.cu -o benchmarks/_variants/libtdx_$tag.so &
done
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -Iinclude benchmarks/variant_bench.cu -o benchmarks/variant_bench -ldl &
wait
ls -la benchmarks/_variants
