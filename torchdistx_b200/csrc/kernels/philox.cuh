// Philox4x32-R counter-based generator, device side.
//
// Algorithm: Salmon et al., "Parallel Random Numbers: As Easy as 1, 2, 3" (SC'11);
// the same generator curand/ATen use on CUDA
// ($TORCH/include/ATen/core/PhiloxRNGEngine.h, curand_philox4x32_x.h).
// CPU restatement: oracle/tdx_oracle.c (philox4x32).  Known-answer vectors:
// tests/golden/philox_kat.json.
//
// Cost on sm_100a: 2 IMAD.WIDE.U32 (quarter rate: 4 pipe / 2 dispatch cycles) + 2 LOP3 (2 cycles) per round (round keys are
// loop-invariant per descriptor and hoisted by the compiler), i.e. 80 dispatch
// cycles per warp and 128 random bits at R = 10 (benchmarks/philox_rate.cu: 84 measured).
#pragma once
#include <cstdint>

namespace tdx {

constexpr uint32_t kPhiloxM0 = 0xD2511F53u;
constexpr uint32_t kPhiloxM1 = 0xCD9E8D57u;
constexpr uint32_t kPhiloxW0 = 0x9E3779B9u;
constexpr uint32_t kPhiloxW1 = 0xBB67AE85u;

template <int ROUNDS>
__device__ __forceinline__ uint4 philox4x32(uint4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < ROUNDS; ++r) {
    // mul.wide.u32 == one IMAD.WIDE.U32 (a plain 64-bit C++ multiply costs an extra add per product)
    unsigned long long p0, p1;
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(p0) : "r"(c.x), "r"(kPhiloxM0));
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(p1) : "r"(c.z), "r"(kPhiloxM1));
    const uint32_t hi0 = static_cast<uint32_t>(p0 >> 32), lo0 = static_cast<uint32_t>(p0);
    const uint32_t hi1 = static_cast<uint32_t>(p1 >> 32), lo1 = static_cast<uint32_t>(p1);
    c = make_uint4(hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0);
    k0 += kPhiloxW0;
    k1 += kPhiloxW1;
  }
  return c;
}

}  // namespace tdx
