// How fast can one B200 run Philox4x32-10 at all?  The RNG kernels are not HBM-bound; this
// microbenchmark measures the ceiling the generator itself sets: N Philox blocks per thread,
// results folded into one word (so nothing is optimised away), one 4-byte store per thread.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -Itorchdistx_b200/csrc/kernels \
//        benchmarks/philox_rate.cu -o benchmarks/philox_rate && benchmarks/philox_rate
//
// Output: blocks/s, the equivalent output bandwidth at 16 bytes per block (8 bf16 elements), and
// issue slots per block (from the SASS count given on the command line, optional).
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "philox.cuh"

using tdx::philox4x32;

template <int ILP, int MODE>
__global__ void __launch_bounds__(256) rate_kernel(uint32_t* out, uint32_t k0, uint32_t k1, int iters) {
  uint32_t acc = 0;
  uint32_t ctr = blockIdx.x * blockDim.x + threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      const uint4 w = philox4x32<10>(make_uint4(ctr, 0u, 7u, 0x80000000u), k0, k1);
      ctr += 0x10000u;
      if (MODE == 0) {
        acc ^= w.x ^ w.y ^ w.z ^ w.w;  // 2 LOP3
      } else if (MODE == 1) {  // + the 8 half-word extractions of a 16-bit generator (PRMT) and 8 adds
        const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          acc += __byte_perm(ws[q], 0u, 0x4410);
          acc += __byte_perm(ws[q], 0u, 0x4432);
        }
      } else {  // + 8 float conversions (PRMT magic, FADD, FFMA) as in the uniform generator
        const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
        float f = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          f += fmaf(__uint_as_float(__byte_perm(ws[q], 0x4b000000u, 0x7410)) - 8388608.0f, 1e-5f, 0.25f);
          f += fmaf(__uint_as_float(__byte_perm(ws[q], 0x4b000000u, 0x7432)) - 8388608.0f, 1e-5f, 0.25f);
        }
        acc ^= __float_as_uint(f);
      }
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int ILP, int MODE>
void run(const char* name, uint32_t* out, int ctas_per_sm) {
  const int iters = 2048 / ILP;
  const int grid = 148 * ctas_per_sm;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  rate_kernel<ILP, MODE><<<grid, 256>>>(out, 1234u, 5678u, iters);
  cudaDeviceSynchronize();
  float best = 1e9f;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(e0);
    rate_kernel<ILP, MODE><<<grid, 256>>>(out, 1234u, 5678u, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  const double blocks = double(grid) * 256.0 * iters * ILP;
  const double bps = blocks / (best * 1e-3);
  printf("{\"bench\": \"%s\", \"ilp\": %d, \"ctas_per_sm\": %d, \"ms\": %.4f, \"philox_blocks_per_s\": %.4g, "
         "\"equiv_GBs_at_16B_per_block\": %.1f, \"cycles_per_warp_block_per_smsp\": %.1f}\n",
         name, ILP, ctas_per_sm, best, bps, bps * 16 / 1e9,
         (best * 1e-3 * 1.965e9) / (blocks / 32.0 / (148.0 * 4.0)));
}

int main() {
  uint32_t* out;
  cudaMalloc(&out, 148 * 8 * 256 * 4);
  for (int c : {4, 8}) {
    run<1, 0>("philox_only", out, c);
    run<2, 0>("philox_only", out, c);
    run<4, 0>("philox_only", out, c);
    run<4, 1>("philox+8prmt+8iadd", out, c);
    run<4, 2>("philox+8(prmt,fadd,ffma,fadd)", out, c);
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { fprintf(stderr, "%s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
