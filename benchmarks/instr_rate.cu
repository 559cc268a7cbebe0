// Issue cost of the integer instructions Philox is made of, one B200 SM sub-partition at a time:
// cycles per warp-instruction for long independent streams (8 chains per thread, 8 warps per
// scheduler), measured with clock64 inside the kernel.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 benchmarks/instr_rate.cu -o benchmarks/instr_rate
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

#define CHAINS 8
#define UNROLL 16

template <int OP>
__global__ void __launch_bounds__(1024) k(uint32_t* out, long long* cyc, uint32_t m_reg, int iters) {
  uint32_t x[CHAINS], y[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) { x[c] = threadIdx.x * 7919u + c; y[c] = c * 31u + 1u; }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
#pragma unroll
      for (int c = 0; c < CHAINS; ++c) {
        if (OP == 0) {  // IMAD.WIDE.U32 with immediate multiplier, hi and lo both used
          unsigned long long p;
          asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(x[c]), "r"(0xD2511F53u));
          x[c] = static_cast<uint32_t>(p >> 32);
          y[c] = static_cast<uint32_t>(p);
        } else if (OP == 1) {  // IMAD.WIDE.U32 with the multiplier in a register
          unsigned long long p;
          asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(x[c]), "r"(m_reg));
          x[c] = static_cast<uint32_t>(p >> 32);
          y[c] = static_cast<uint32_t>(p);
        } else if (OP == 2) {  // mul.hi only
          asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(x[c]) : "r"(x[c]), "r"(0xD2511F53u));
        } else if (OP == 3) {  // mul.lo only
          asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(x[c]) : "r"(x[c]), "r"(0xD2511F53u));
        } else if (OP == 4) {  // LOP3 (3-input xor)
          asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(x[c]) : "r"(x[c]), "r"(y[c]), "r"(m_reg));
        } else if (OP == 5) {  // mad.wide with 64-bit addend
          unsigned long long p = (static_cast<unsigned long long>(x[c]) << 32) | y[c];
          asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(p) : "r"(x[c]), "r"(0xD2511F53u));
          x[c] = static_cast<uint32_t>(p >> 32);
          y[c] = static_cast<uint32_t>(p);
        } else if (OP == 6) {  // PRMT
          asm volatile("prmt.b32 %0, %1, %2, 0x4321;" : "=r"(x[c]) : "r"(x[c]), "r"(y[c]));
        } else if (OP == 7) {  // FFMA 3-reg
          float f = __uint_as_float(x[c]);
          asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f) : "f"(__uint_as_float(y[c])), "f"(__uint_as_float(m_reg)));
          x[c] = __float_as_uint(f);
        } else if (OP == 8) {  // mul.wide.u16 (16x16 -> 32)
          asm volatile("{ .reg .u16 a, b; mov.b32 {a, b}, %1; mul.wide.u16 %0, a, b; }" : "=r"(x[c]) : "r"(x[c]));
        } else if (OP == 10) {  // IMAD (32-bit lo), register multiplier
          asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(m_reg), "r"(y[c]));
        } else if (OP == 11) {  // IDP.2A
          asm volatile("dp2a.lo.u32.u32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(m_reg), "r"(y[c]));
        } else if (OP == 12) {  // HFMA2.BF16
          asm volatile("fma.rn.bf16x2 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(m_reg), "r"(y[c]));
        } else if (OP == 13) {  // F2FP pack
          asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(x[c]) : "f"(__uint_as_float(x[c])), "f"(__uint_as_float(y[c])));
        } else if (OP == 14) {  // IADD3
          asm volatile("add.u32 %0, %0, %1;" : "+r"(x[c]) : "r"(y[c]));
        } else if (OP == 15) {  // MUFU.LG2
          float f = __uint_as_float(x[c]);
          asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(f));
          x[c] = __float_as_uint(f);
        } else if (OP == 16) {  // FMNMX
          float f = __uint_as_float(x[c]);
          asm volatile("min.f32 %0, %0, %1;" : "+f"(f) : "f"(__uint_as_float(y[c])));
          x[c] = __float_as_uint(f);
        } else if (OP == 17) {  // packed bf16 min
          asm volatile("min.bf16x2 %0, %0, %1;" : "+r"(x[c]) : "r"(y[c]));
        } else if (OP == 18) {  // FFMA + IMAD.WIDE pair: does FFMA hide behind the 4-cycle IMAD.WIDE?
          unsigned long long p;
          asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(x[c]), "r"(0xD2511F53u));
          float f = __uint_as_float(y[c]);
          asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f) : "f"(__uint_as_float(m_reg)), "f"(1.5f));
          x[c] = static_cast<uint32_t>(p >> 32) ^ static_cast<uint32_t>(p);
          y[c] = __float_as_uint(f);
        } else if (OP == 19) {  // IDP + IMAD.WIDE pair: same pipe?
          unsigned long long p;
          asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(x[c]), "r"(0xD2511F53u));
          asm volatile("dp2a.lo.u32.u32 %0, %0, %1, %2;" : "+r"(y[c]) : "r"(m_reg), "r"(static_cast<uint32_t>(p)));
          x[c] = static_cast<uint32_t>(p >> 32);
        } else if (OP == 20) {  // IMAD.WIDE + 2 LOP3 (does the ALU pipe keep up two-for-one?)
          unsigned long long p;
          asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(x[c]), "r"(0xD2511F53u));
          asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(x[c]) : "r"(static_cast<uint32_t>(p >> 32)), "r"(y[c]), "r"(m_reg));
          asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(y[c]) : "r"(static_cast<uint32_t>(p)), "r"(y[c]), "r"(m_reg));
        } else if (OP == 21) {  // I2FP (u32 -> f32)
          float f;
          asm volatile("cvt.rn.f32.u32 %0, %1;" : "=f"(f) : "r"(x[c]));
          x[c] = __float_as_uint(f);
        } else if (OP == 22) {  // 16x16->32 via mul24? : mul.lo on 16-bit halves using mad.wide.u16
          asm volatile("{ .reg .u16 a, b; mov.b32 {a, b}, %1; mad.wide.u16 %0, a, b, %2; }" : "=r"(x[c]) : "r"(x[c]), "r"(y[c]));
        } else if (OP == 9) {  // IMAD.WIDE imm + LOP3 alternating (Philox's mix)
          unsigned long long p;
          asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(x[c]), "r"(0xD2511F53u));
          asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(x[c]) : "r"(static_cast<uint32_t>(p >> 32)), "r"(y[c]), "r"(m_reg));
          y[c] = static_cast<uint32_t>(p);
        }
      }
    }
  }
  const long long t1 = clock64();
  uint32_t acc = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) acc ^= x[c] ^ y[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, int per_iter_instrs) {
  uint32_t* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 200;
  k<OP><<<148, 1024>>>(out, cyc, 0xCD9E8D57u, iters);
  k<OP><<<148, 1024>>>(out, cyc, 0xCD9E8D57u, iters);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
  // per scheduler: 8 warps, each issues iters*UNROLL*CHAINS*per_iter_instrs instructions
  const double instrs = 8.0 * iters * UNROLL * CHAINS * per_iter_instrs;
  printf("{\"op\": \"%s\", \"cycles_per_warp_instr_per_smsp\": %.3f}\n", name, avg / instrs);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<0>("IMAD.WIDE.U32 imm", 1);
  run<1>("IMAD.WIDE.U32 reg", 1);
  run<2>("mul.hi.u32 imm", 1);
  run<3>("mul.lo.u32 imm", 1);
  run<4>("LOP3 (3 regs)", 1);
  run<5>("mad.wide.u32 imm + 64-bit addend", 1);
  run<6>("PRMT", 1);
  run<7>("FFMA 3-reg", 1);
  run<8>("mul.wide.u16", 1);
  run<9>("IMAD.WIDE imm + LOP3 pair", 2);
  run<10>("IMAD lo (reg multiplier, addend)", 1);
  run<11>("IDP.2A", 1);
  run<12>("HFMA2.BF16", 1);
  run<13>("F2FP.BF16 pack", 1);
  run<14>("IADD3", 1);
  run<15>("MUFU.LG2", 1);
  run<16>("FMNMX", 1);
  run<17>("HMNMX2.BF16", 1);
  run<18>("IMAD.WIDE + FFMA pair", 2);
  run<19>("IMAD.WIDE + IDP pair", 2);
  run<20>("IMAD.WIDE + 2 LOP3", 3);
  run<21>("I2FP.F32.U32", 1);
  run<22>("mad.wide.u16", 1);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { fprintf(stderr, "%s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
