#!/bin/sh
# Builds A/B variants of libtdx_init.so (knobs of the table kernel) + the C++ harness.
# usage: benchmarks/build_variants.sh "tag:-Dflags" ...
# knobs (tdx_init_kernels.cu): TDX_LUT_ELEMS_A / _B (looked-up elements of even / odd vectors, normal),
#   TDX_LUT_UNIFORM_ELEMS, TDX_LUT_GROUP (vectors per NaN test), TDX_LUT_PACK (0 IMAD, 1 PRMT),
#   TDX_LUT_VECS (vectors per thread and tile), TDX_LUT_MAX_CHUNK_LOG2 (largest grab),
#   TDX_LUT_PKEYS (round keys as kernel parameters), TDX_UNIFORM16_PACKED, TDX_VECS,
#   TDX_EXPERIMENTAL_ALGOS (Philox-7 / Box-Muller-16 kernels for benchmarks/kernel_sweep.py)
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
