// libtdx_init.so -- fused parameter-initialisation kernels for B200 (sm_100a).
//
// What this replaces: the reference materialises a deferred tensor by replaying
// every recorded aten op through the dispatcher, one stock ATen kernel per op,
// dead (overwritten) ops included
// (reference src/cc/torchdistx/deferred_init.cc:506-528 OpNode::materialize,
//  :256-272 Op::materialize, :218-220 handle.callBoxed).
// Here one TdxInitDesc carries the folded expression of a whole tensor and a
// persistent kernel streams it to HBM exactly once: Philox counter-based RNG in
// registers, 128-bit coalesced stores, zero reads, no intermediates.
//
// Roofline: pure HBM write.  Algorithmic bytes = elem_count * itemsize per descriptor.  What
// actually bounds the RNG kernels on B200 is the dispatch port (benchmarks/instr_rate.cu):
// IMAD.WIDE holds its pipe 4 cycles per warp instruction and, like every other half-rate
// instruction (LOP3, PRMT, IMAD, IDP, F2FP, HFMA2), the dispatch port 2; Philox4x32-10 is 20 + 20 of
// them = 80 cycles per 128 random bits, i.e. 7.0 TB/s-equivalent at one block per 16-byte vector
// with nothing else in the loop (benchmarks/philox_rate.cu).  So 16-bit outputs draw 16 random bits
// per element (8 elements per Philox block), large 16-bit descriptors turn half-words into outputs
// through a shared-memory table (tdx_lut16_kernel), counter.y is hoisted per tile, and the fp32
// generators spend one I2FP per element.  See DESIGN.md section 4.
//
// Scheduling: one kernel launch per kernel family (<= a handful per module, not one per tensor).
// 256-thread kernels: grid = #SM x resident CTAs, CTAs grab runs of tiles of the family's global
// tile space from an atomic counter (grab size chosen per launch).  Table kernel: one 1024-thread
// CTA per SM, grabs from a host-built, descriptor-pure work list.  Small and large tensors, shards
// and ragged tails balance across all 148 SMs either way; every kernel zeroes its work counter
// when its last CTA leaves.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "philox.cuh"
#include "tdx_init.h"

namespace tdx {

#ifndef TDX_VECS
#define TDX_VECS 4
#endif
#ifndef TDX_UNIFORM16_PACKED
#define TDX_UNIFORM16_PACKED 1
#endif
constexpr int kThreads = 256;
constexpr int kVecsPerThread = TDX_VECS;
constexpr int kTileVecs = kThreads * kVecsPerThread;  // 1024 x 16 B = 16 KiB per tile
constexpr int kTilesPerChunk = 64 / TDX_VECS;          // 256 KiB per work grab
#ifndef TDX_STATIC_TILES
#define TDX_STATIC_TILES 8
#endif
// launches of up to this many tiles per resident CTA are scheduled statically (GroupArgs::static_grabs)
constexpr unsigned long long kStaticTilesPerCta = TDX_STATIC_TILES;

// ---------------------------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float mufu_lg2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float mufu_sqrt(float x) {
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float mufu_sin(float x) {
  float y;
  asm("sin.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float mufu_cos(float x) {
  float y;
  asm("cos.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 128-bit streaming store: the data is written once and never re-read by this kernel.
__device__ __forceinline__ void store_vec(void* p, uint4 v) {
  // .cs (evict-first): measured 2 % ahead of the default, .wt and L1::no_allocate forms
  asm volatile("st.global.cs.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w)
               : "memory");
}

template <class Out>
struct OutTraits;
template <>
struct OutTraits<float> {
  static constexpr int kEpv = 4;
  static constexpr int kDtype = TDX_F32;
  __device__ static __forceinline__ float round_through(float v) { return v; }
  __device__ static __forceinline__ void store_one(void* dst, uint64_t i, float v) {
    static_cast<float*>(dst)[i] = v;
  }
  __device__ static __forceinline__ uint4 pack(const float (&v)[4]) {
    return make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]),
                      __float_as_uint(v[3]));
  }
  // largest representable value strictly below `to`
  __device__ static __forceinline__ float prev(float to) {
    uint32_t b = __float_as_uint(to);
    if (to > 0.f) return __uint_as_float(b - 1);
    if (to < 0.f) return __uint_as_float(b + 1);
    return __uint_as_float(0x80000001u);
  }
};
template <>
struct OutTraits<__nv_bfloat16> {
  static constexpr int kEpv = 8;
  static constexpr int kDtype = TDX_BF16;
  __device__ static __forceinline__ float round_through(float v) {
    return __bfloat162float(__float2bfloat16_rn(v));
  }
  __device__ static __forceinline__ void store_one(void* dst, uint64_t i, float v) {
    static_cast<__nv_bfloat16*>(dst)[i] = __float2bfloat16_rn(v);
  }
  __device__ static __forceinline__ uint4 pack(const float (&v)[8]) {
    uint4 r;
    __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]);
    __nv_bfloat162 b = __floats2bfloat162_rn(v[2], v[3]);
    __nv_bfloat162 c = __floats2bfloat162_rn(v[4], v[5]);
    __nv_bfloat162 d = __floats2bfloat162_rn(v[6], v[7]);
    r.x = *reinterpret_cast<uint32_t*>(&a);
    r.y = *reinterpret_cast<uint32_t*>(&b);
    r.z = *reinterpret_cast<uint32_t*>(&c);
    r.w = *reinterpret_cast<uint32_t*>(&d);
    return r;
  }
  __device__ static __forceinline__ uint32_t pack2(float a, float b) {
    __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&p);
  }
  __device__ static __forceinline__ uint32_t min2(uint32_t a, uint32_t b) {
    const __nv_bfloat162 m = __hmin2(*reinterpret_cast<const __nv_bfloat162*>(&a),
                                     *reinterpret_cast<const __nv_bfloat162*>(&b));
    return *reinterpret_cast<const uint32_t*>(&m);
  }
  __device__ static __forceinline__ float prev(float to) {  // `to` is exactly a bf16 value
    uint32_t b = __float_as_uint(to);
    if (to > 0.f) return __uint_as_float(b - 0x10000u);
    if (to < 0.f) return __uint_as_float(b + 0x10000u);
    return __uint_as_float(0x80010000u);
  }
};
template <>
struct OutTraits<__half> {
  static constexpr int kEpv = 8;
  static constexpr int kDtype = TDX_F16;
  __device__ static __forceinline__ float round_through(float v) {
    return __half2float(__float2half_rn(v));
  }
  __device__ static __forceinline__ void store_one(void* dst, uint64_t i, float v) {
    static_cast<__half*>(dst)[i] = __float2half_rn(v);
  }
  __device__ static __forceinline__ uint4 pack(const float (&v)[8]) {
    uint4 r;
    __half2 a = __floats2half2_rn(v[0], v[1]);
    __half2 b = __floats2half2_rn(v[2], v[3]);
    __half2 c = __floats2half2_rn(v[4], v[5]);
    __half2 d = __floats2half2_rn(v[6], v[7]);
    r.x = *reinterpret_cast<uint32_t*>(&a);
    r.y = *reinterpret_cast<uint32_t*>(&b);
    r.z = *reinterpret_cast<uint32_t*>(&c);
    r.w = *reinterpret_cast<uint32_t*>(&d);
    return r;
  }
  __device__ static __forceinline__ uint32_t pack2(float a, float b) {
    __half2 p = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&p);
  }
  __device__ static __forceinline__ uint32_t min2(uint32_t a, uint32_t b) {
    const __half2 m = __hmin2(*reinterpret_cast<const __half2*>(&a), *reinterpret_cast<const __half2*>(&b));
    return *reinterpret_cast<const uint32_t*>(&m);
  }
  __device__ static __forceinline__ float prev(float to) {  // `to` is exactly an fp16 value
    unsigned short h = __half_as_ushort(__float2half_rn(to));
    if (to > 0.f) return __half2float(__ushort_as_half(static_cast<unsigned short>(h - 1)));
    if (to < 0.f) return __half2float(__ushort_as_half(static_cast<unsigned short>(h + 1)));
    return __half2float(__ushort_as_half(static_cast<unsigned short>(0x8001)));
  }
};

// ---------------------------------------------------------------------------------------------
// epilogue: the in-place elementwise ops recorded after the source op
// ---------------------------------------------------------------------------------------------
struct EpiParams {
  uint32_t n;
  uint32_t flags;
  uint32_t op[TDX_MAX_EPI];
  float a[TDX_MAX_EPI];
  float b[TDX_MAX_EPI];
};

__device__ __forceinline__ EpiParams load_epi(const TdxInitDesc& d) {
  EpiParams e;
  e.n = d.n_epi;
  e.flags = d.reserved;
#pragma unroll
  for (int i = 0; i < TDX_MAX_EPI; ++i) {
    e.op[i] = d.epi[i].op;
    e.a[i] = d.epi[i].a;
    e.b[i] = d.epi[i].b;
  }
  return e;
}

template <class Out>
__device__ __forceinline__ float apply_epi(const EpiParams& e, float v) {
  using T = OutTraits<Out>;
  if (!(e.flags & TDX_FLAG_SRC_NOROUND)) v = T::round_through(v);
#pragma unroll
  for (int i = 0; i < TDX_MAX_EPI; ++i) {
    if (i < static_cast<int>(e.n)) {
      switch (e.op[i] & 0xffu) {
        case TDX_EPI_MUL: v = v * e.a[i]; break;
        case TDX_EPI_ADD: v = v + e.a[i]; break;
        case TDX_EPI_ERFINV: v = erfinvf(v); break;
        case TDX_EPI_CLAMP: v = fminf(fmaxf(v, e.a[i]), e.b[i]); break;
        case TDX_EPI_RPOW: v = powf(e.a[i], v); break;
        case TDX_EPI_RECIP: v = 1.0f / v; break;
        default: break;
      }
      if (!(e.op[i] & TDX_EPI_NOROUND)) v = T::round_through(v);
    }
  }
  return v;
}

// ---------------------------------------------------------------------------------------------
// generators.  Each produces the EPV final values (as fp32, before the rounding of the store) of
// global vector `gv`, i.e. of global elements [gv*EPV, gv*EPV+EPV).
// ---------------------------------------------------------------------------------------------
struct PhiloxCtx {
  uint32_t k0, k1;  // key = seed
  uint32_t cz, cw;  // counter.z/.w = stream id (generator offset), bit 31 of w set
};
__device__ __forceinline__ PhiloxCtx load_philox(const TdxInitDesc& d) {
  PhiloxCtx c;
  c.k0 = static_cast<uint32_t>(d.philox_seed);
  c.k1 = static_cast<uint32_t>(d.philox_seed >> 32);
  c.cz = static_cast<uint32_t>(d.philox_offset);
  c.cw = static_cast<uint32_t>(d.philox_offset >> 32) | 0x80000000u;
  return c;
}
// Counter layout (normative, include/tdx_init.h): (off_lo, blk_hi, off_hi | flags, blk_lo).  The word
// that differs from thread to thread (blk_lo) sits in position 3, which a Philox round only XORs and
// moves: both products of round 1, one of round 2 and one of round 3 then depend on per-tile
// uniform values only and are hoisted out of the vector loop -- 16 IMAD.WIDE + 18 LOP3 per block
// instead of 20 + 20 (IMAD.WIDE is quarter rate on sm_100: the instruction that bounds these
// kernels).  Every bit of blk_lo still passes through nine multiplying rounds (Philox4x32-7 is the
// published Crush-resistant minimum).
template <int R>
__device__ __forceinline__ uint4 philox_block(const PhiloxCtx& c, uint64_t blk, uint32_t wflag = 0) {
  return philox4x32<R>(
      make_uint4(c.cz, static_cast<uint32_t>(blk >> 32), c.cw | wflag, static_cast<uint32_t>(blk)),
      c.k0, c.k1);
}

// half-word `e` (0..7) of a Philox block as a float holding 2^23 + k, k in [0, 65535]
__device__ __forceinline__ float halfword_as_magic(const uint4& w, int e) {
  const uint32_t word = (e >> 1) == 0 ? w.x : (e >> 1) == 1 ? w.y : (e >> 1) == 2 ? w.z : w.w;
  // one PRMT each: {0x4b, 0x00, k_hi, k_lo}
  const uint32_t bits = __byte_perm(word, 0x4b000000u, (e & 1) ? 0x7432 : 0x7410);
  return __uint_as_float(bits);
}

// half-word `e` of a Philox block as the float k (one I2F.U16 with a half-word operand select: no
// byte-permute constant to keep in a register, which the table kernel -- 64 registers per thread --
// re-materialised for every vector)
__device__ __forceinline__ float halfword_as_float(const uint4& w, int e) {
  const uint32_t word = (e >> 1) == 0 ? w.x : (e >> 1) == 1 ? w.y : (e >> 1) == 2 ? w.z : w.w;
  unsigned short lo, hi;
  asm("mov.b32 {%0, %1}, %2;" : "=h"(lo), "=h"(hi) : "r"(word));
  float f;
  asm("cvt.rn.f32.u16 %0, %1;" : "=f"(f) : "h"((e & 1) ? hi : lo));
  return f;
}

// ---- uniform, 16 random bits per element (bf16 / fp16 outputs) -----------------------------
template <class Out, int R, bool EPI>
struct GenUniform16 {
  using OutT = Out;
  static constexpr int kEpv = 8;
  struct Params {
    PhiloxCtx ph;
    float from, scale, to_prev;
    EpiParams epi;
  };
  __device__ static __forceinline__ Params setup(const TdxInitDesc& d) {
    Params p;
    p.ph = load_philox(d);
    const float from = static_cast<float>(d.p0), to = static_cast<float>(d.p1);
    p.from = from;
    p.scale = (to - from) * 1.52587890625e-05f;  // * 2^-16, exact
    p.to_prev = (to > from) ? OutTraits<Out>::prev(to) : to;
    if (EPI) p.epi = load_epi(d);
#if TDX_UNIFORM16_PACKED
    // the three scalars are the same for every thread; out of a warp reduction they live in uniform
    // registers and the per-element FFMA reads two vector registers instead of three
    p.from = __uint_as_float(__reduce_or_sync(0xffffffffu, __float_as_uint(p.from)));
    p.scale = __uint_as_float(__reduce_or_sync(0xffffffffu, __float_as_uint(p.scale)));
    p.to_prev = __uint_as_float(__reduce_or_sync(0xffffffffu, __float_as_uint(p.to_prev)));
#endif
    return p;
  }
  __device__ static __forceinline__ void gen(const Params& p, uint64_t gv, float (&v)[8]) {
    const uint4 w = philox_block<R>(p.ph, gv);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float k = halfword_as_magic(w, e) - 8388608.0f;  // exact
      float x = fminf(fmaf(k, p.scale, p.from), p.to_prev);
      v[e] = EPI ? apply_epi<Out>(p.epi, x) : x;
    }
  }
  // Whole-vector form for the hot path (no epilogue): the clamp below `to` is applied to the packed
  // pairs (4 packed mins instead of 8 scalar ones).  Rounding is monotone and to_prev is a value of
  // the output type (checked by the caller: packed_ok), so min(round(x), to_prev) == round(min(x,
  // to_prev)): the same bits as gen().
  __device__ static __forceinline__ bool packed_ok(const Params& p) {
    return !EPI && OutTraits<Out>::round_through(p.to_prev) == p.to_prev;
  }
  __device__ static __forceinline__ uint4 gen_vec(const Params& p, uint64_t gv) {
    using T = OutTraits<Out>;
    const uint4 w = philox_block<R>(p.ph, gv);
    const uint32_t lim = T::pack2(p.to_prev, p.to_prev);
    uint32_t r[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float a = fmaf(halfword_as_magic(w, 2 * q) - 8388608.0f, p.scale, p.from);
      const float b = fmaf(halfword_as_magic(w, 2 * q + 1) - 8388608.0f, p.scale, p.from);
      r[q] = T::min2(T::pack2(a, b), lim);
    }
    return make_uint4(r[0], r[1], r[2], r[3]);
  }
};
template <class Gen>
struct HasGenVec { static constexpr bool value = false; };
template <class Out, int R>
struct HasGenVec<GenUniform16<Out, R, false>> { static constexpr bool value = TDX_UNIFORM16_PACKED != 0; };

// ---- uniform, 24-bit mantissa from 32 random bits per element -------------------------------
// fp32 outputs, and the "wide" form of 16-bit outputs (TDX_ALGO_WIDE32): exactly what
// `x.uniform_()` on an fp32 tensor followed by `.to(bf16)` produces -- fp32 arithmetic, fp32
// bounds, one rounding at the store (epilogue steps flagged NOROUND stay in fp32 too).
template <class Out, int R, bool EPI>
struct GenUniform32 {
  using OutT = Out;
  static constexpr int kEpv = OutTraits<Out>::kEpv;  // 4 (f32) or 8 (16-bit outputs: 2 blocks)
  struct Params {
    PhiloxCtx ph;
    float from, scale, to_prev;
    EpiParams epi;
  };
  __device__ static __forceinline__ Params setup(const TdxInitDesc& d) {
    Params p;
    p.ph = load_philox(d);
    const float from = static_cast<float>(d.p0), to = static_cast<float>(d.p1);
    p.from = from;
    p.scale = (to - from) * 2.3283064365386963e-10f;  // * 2^-32, exact
    p.to_prev = (to > from) ? OutTraits<float>::prev(to) : to;
    if (EPI) p.epi = load_epi(d);
    return p;
  }
  __device__ static __forceinline__ void gen(const Params& p, uint64_t gv, float (&v)[kEpv]) {
#pragma unroll
    for (int b = 0; b < kEpv / 4; ++b) {
      const uint4 w = philox_block<R>(p.ph, gv * (kEpv / 4) + b);
      const uint32_t x[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        // all 32 bits, rounded to the nearest fp32 (one I2FP; a shift to 24 exact bits would cost as
        // much again on the dispatch port); the top half-bin rounds to 2^32 and is clamped below `to`
        const float k = __uint2float_rn(x[e]);
        float r = fminf(fmaf(k, p.scale, p.from), p.to_prev);
        v[b * 4 + e] = EPI ? apply_epi<Out>(p.epi, r) : r;
      }
    }
  }
};

// ---- normal: Box-Muller on 2 x 32 random bits per pair -------------------------------------
// (x, y) -> radius from x (32-bit resolution, (0,1]), angle from the top 23 bits of y.
// Returns the two UNSCALED directions; the caller multiplies the radius by std once per pair.
__device__ __forceinline__ void box_muller32(uint32_t x, uint32_t y, float& r, float& c, float& s) {
  const float u1 = fmaf(__uint2float_rn(x), 2.3283064365386963e-10f, 1.1641532182693481e-10f);
  // r = sqrt(-2 ln u1) = sqrt(-2 ln2 * lg2 u1)
  r = mufu_sqrt(-1.3862943611198906f * mufu_lg2(u1));
  // angle: y as a SIGNED 32-bit integer times pi / 2^31, in [-pi, pi] (one I2FP + one FMUL)
  const float ang = __int2float_rn(static_cast<int>(y)) * 1.4629180792671596e-09f;
  c = mufu_cos(ang);
  s = mufu_sin(ang);
}

template <class Out, int R, bool EPI>
struct GenNormalBM32 {
  using OutT = Out;
  static constexpr int kEpv = OutTraits<Out>::kEpv;  // 4 (f32) or 8 (16-bit outputs: 2 blocks)
  struct Params {
    PhiloxCtx ph;
    float mean, std;
    EpiParams epi;
  };
  __device__ static __forceinline__ Params setup(const TdxInitDesc& d) {
    Params p;
    p.ph = load_philox(d);
    p.mean = static_cast<float>(d.p0);
    p.std = static_cast<float>(d.p1);
    if (EPI) p.epi = load_epi(d);
    return p;
  }
  __device__ static __forceinline__ void gen(const Params& p, uint64_t gv, float (&v)[kEpv]) {
#pragma unroll
    for (int b = 0; b < kEpv / 4; ++b) {
      const uint4 w = philox_block<R>(p.ph, gv * (kEpv / 4) + b);
      float r0, r1, d[4];
      box_muller32(w.x, w.y, r0, d[0], d[1]);
      box_muller32(w.z, w.w, r1, d[2], d[3]);
      const float rs[2] = {r0 * p.std, r1 * p.std};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float r = fmaf(rs[e >> 1], d[e], p.mean);
        v[b * 4 + e] = EPI ? apply_epi<Out>(p.epi, r) : r;
      }
    }
  }
};

// ---- normal: inverse CDF on 16 random bits per element (bf16 / fp16 outputs) ----------------
// k in [1, 65535]: x = k/32768 - 1 (a grid symmetric about 0), z = sqrt(2) erfinv(x)
//                  = x * P(lg2(1 - x^2)),  P = degree-5 minimax fit, |rel err| < 6.5e-5 on the grid.
// k == 0 (probability 2^-16): the two half-bins beyond the last grid points, i.e. the tails
//                  u < 2^-17 and u > 1 - 2^-17; resolved with 32 more Philox bits (sign + 31-bit
//                  position inside the tail), reaching |z| ~ 7.9 sigma.
__device__ __noinline__ float icdf16_tail(uint32_t word) {
  const float sgn = (word & 0x80000000u) ? 1.0f : -1.0f;
  // p in (0, 2^-17): lower-tail probability
  const float p = (static_cast<float>(word & 0x7fffffffu) + 0.5f) * 4.656612873077393e-10f *
                  7.62939453125e-06f;
  return sgn * -normcdfinvf(p);
}

template <class Out, int R, bool EPI>
struct GenNormalICDF16 {
  using OutT = Out;
  static constexpr int kEpv = 8;
  struct Params {
    PhiloxCtx ph;
    float mean, std;
    float c0, c1, c2, c3, c4, c5;  // std * P coefficients
    EpiParams epi;
  };
  __device__ static __forceinline__ Params setup(const TdxInitDesc& d) {
    Params p;
    p.ph = load_philox(d);
    p.mean = static_cast<float>(d.p0);
    p.std = static_cast<float>(d.p1);
    p.c0 = p.std * 0x1.40de66p+0f;
    p.c1 = p.std * -0x1.d03266p-3f;
    p.c2 = p.std * 0x1.26ef76p-7f;
    p.c3 = p.std * 0x1.bfaecap-10f;
    p.c4 = p.std * 0x1.9c35c0p-14f;
    p.c5 = p.std * 0x1.152c90p-19f;
    if (EPI) p.epi = load_epi(d);
    return p;
  }
  // cold path (k == 0), kept out of line so the hot loop stays inside the instruction cache;
  // scalar in / scalar out so that nothing of the hot path is forced into local memory
  __device__ static __noinline__ float refine_one(uint32_t k0, uint32_t k1, uint32_t cz, uint32_t cw,
                                                  uint64_t gv, int e) {
    PhiloxCtx c{k0, k1, cz, cw};
    const uint4 t = philox_block<R>(c, gv, e < 4 ? 0x40000000u : 0x20000000u);
    const int s = e & 3;
    return icdf16_tail(s == 0 ? t.x : s == 1 ? t.y : s == 2 ? t.z : t.w);
  }
  // one element: `magic` = 2^23 + k as a float; returns the value, `t` = 1 - x^2 (0 iff k == 0)
  __device__ static __forceinline__ float element_l(const Params& p, float magic, float& t, float& l) {
    return element_x(p, fmaf(magic, 3.0517578125e-05f, -257.0f), t, l);  // k/32768 - 1 (exact)
  }
  // ... from k itself as a float: the same x, exactly (both forms are exact)
  __device__ static __forceinline__ float element_kf(const Params& p, float kf) {
    float t, l;
    return element_x(p, fmaf(kf, 3.0517578125e-05f, -1.0f), t, l);
  }
  __device__ static __forceinline__ float element_x(const Params& p, float x, float& t, float& l) {
    t = fmaf(-x, x, 1.0f);
    l = mufu_lg2(t);
    float q = fmaf(p.c5, l, p.c4);
    q = fmaf(q, l, p.c3);
    q = fmaf(q, l, p.c2);
    q = fmaf(q, l, p.c1);
    q = fmaf(q, l, p.c0);
    return fmaf(q, x, p.mean);
  }
  __device__ static __forceinline__ float element(const Params& p, float magic, float& t) {
    float l;
    return element_l(p, magic, t, l);
  }
  __device__ static __forceinline__ void gen(const Params& p, uint64_t gv, float (&v)[8]) {
    const uint4 w = philox_block<R>(p.ph, gv);
    float tmin = 1.0f;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float t;
      v[e] = element(p, halfword_as_magic(w, e), t);
      tmin = fminf(tmin, t);
    }
    if (tmin <= 0.0f) {  // some k == 0 in this vector: ~1.2e-4 of the vectors
      const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        if (((ws[e >> 1] >> (16 * (e & 1))) & 0xffffu) == 0)
          v[e] = fmaf(refine_one(p.ph.k0, p.ph.k1, p.ph.cz, p.ph.cw, gv, e), p.std, p.mean);
      }
    }
    if (EPI) {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = apply_epi<Out>(p.epi, v[e]);
    }
  }
};

// ---- normal: Box-Muller on 16-bit pairs (experimental; no tail refinement) ------------------
template <class Out, int R, bool EPI>
struct GenNormalBM16 {
  using OutT = Out;
  static constexpr int kEpv = 8;
  struct Params {
    PhiloxCtx ph;
    float mean, std;
    EpiParams epi;
  };
  __device__ static __forceinline__ Params setup(const TdxInitDesc& d) {
    Params p;
    p.ph = load_philox(d);
    p.mean = static_cast<float>(d.p0);
    p.std = static_cast<float>(d.p1);
    if (EPI) p.epi = load_epi(d);
    return p;
  }
  __device__ static __forceinline__ void gen(const Params& p, uint64_t gv, float (&v)[8]) {
    const uint4 w = philox_block<R>(p.ph, gv);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float a = halfword_as_magic(w, 2 * j);      // 2^23 + k1
      const float b = halfword_as_magic(w, 2 * j + 1);  // 2^23 + k2
      const float u1 = fmaf(a - 8388608.0f, 1.52587890625e-05f, 7.62939453125e-06f);  // (k1+.5)/65536
      const float r = mufu_sqrt(-1.3862943611198906f * mufu_lg2(u1)) * p.std;
      const float ang = fmaf(b, 9.5873799242852573e-05f, -804.24771931898707f - 3.1415926535f);
      v[2 * j] = fmaf(r, mufu_cos(ang), p.mean);
      v[2 * j + 1] = fmaf(r, mufu_sin(ang), p.mean);
    }
    if (EPI) {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = apply_epi<Out>(p.epi, v[e]);
    }
  }
};

// ---------------------------------------------------------------------------------------------
// persistent scheduler shared by all kernels
// ---------------------------------------------------------------------------------------------
struct GroupArgs {
  const unsigned long long* tile_prefix;  // [n_desc + 1], exclusive prefix sum of tiles per desc
  const TdxInitDesc* descs;               // [n_desc]
  unsigned long long total_tiles;
  unsigned int* counter;  // chunk counter: zero at launch; the last CTA to leave zeroes it again
  unsigned int* done;     // CTAs that have left
  uint32_t n_desc;
  // Philox round keys of the group's seed when every descriptor of the group shares it
  // (seed_shared != 0): kernel parameters live in the constant bank, so the 20 keys are operands of
  // the Philox LOP3s directly -- no registers, no per-vector key arithmetic.
  uint32_t seed_shared;
  uint32_t tiles_per_chunk;    // 256-thread kernels: tiles per work grab (host-chosen per launch)
  uint32_t n_chunks;           // table kernel: host-built work list (see build_plan)
  const uint4* chunks;         // {descriptor, tiles, first tile in descriptor (lo, hi)}
  // table kernel: the first n_static entries of the list are pre-assigned -- CTA b owns entries
  // [seg_off[b], seg_off[b + 1]) -- and the rest is handed out by the counter (for_each_listed_chunk)
  uint32_t n_static;
  const uint32_t* seg_off;     // [gridDim.x + 1]
  // 256-thread kernels, launches of at most a few tiles per resident CTA: grid = number of grabs,
  // CTA b takes grab b -- no work counter, no exit protocol (a 1 MB tensor is then one kernel
  // launch's latency plus its stores, like a stock elementwise kernel)
  uint32_t static_grabs;
  uint32_t rk[20];
  // Single-descriptor launches (a lone tensor: materialize_tensor, the kernel sweep) carry their
  // descriptor in the kernel parameters: no dependent global loads stand between the launch and
  // the first store.
  uint32_t inline_desc;  // `one` is the group's (only) descriptor
  TdxInitDesc one;
};

// the descriptor `di` of a group (kernels take GroupArgs as a __grid_constant__ parameter, so the
// address of the inline copy is a constant-bank address)
__device__ __forceinline__ const TdxInitDesc* desc_of(const GroupArgs& g, uint32_t di) {
  return g.inline_desc ? &g.one : g.descs + di;
}

// Every CTA calls this once, after its last (failed) grab: the last one to arrive puts the two
// counters back to zero, so a plan can be launched again without a memset in between.
__device__ __forceinline__ void leave_grid(const GroupArgs& g) {
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(g.done, 1u) == gridDim.x - 1u) {
      *g.counter = 0u;
      *g.done = 0u;
      __threadfence();
    }
  }
}

// Calls f(desc_index, first_tile_in_desc, n_tiles) for runs of consecutive tiles of one descriptor.
// A work grab is g.tiles_per_chunk tiles: 256 KiB for launches that fill the machine, down to one
// 16 KiB tile for small ones (a 1 MB tensor is 64 single-tile grabs on 64 SMs, not 4 grabs of 16
// tiles on 4).
template <class F>
__device__ __forceinline__ void run_tiles(const GroupArgs& g, unsigned long long t, unsigned long long last, F&& f) {
  if (g.n_desc == 1) {  // one descriptor owns every tile
    f(0u, t, last - t);
    return;
  }
  uint32_t lo = 0, hi = g.n_desc;  // find d with prefix[d] <= t < prefix[d+1]
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (__ldg(g.tile_prefix + mid) <= t) lo = mid; else hi = mid;
  }
  uint32_t d = lo;
  while (t < last) {
    unsigned long long dend = __ldg(g.tile_prefix + d + 1);
    while (dend <= t) dend = __ldg(g.tile_prefix + (++d) + 1);  // skip empty descriptors
    const unsigned long long stop = min(dend, last);
    f(d, t - __ldg(g.tile_prefix + d), stop - t);
    t = stop;
  }
}

template <class F>
__device__ __forceinline__ void for_each_tile_run(const GroupArgs& g, F&& f) {
  const unsigned int tpc = g.tiles_per_chunk;
  if (g.static_grabs) {  // (uniform across the grid: a kernel parameter)
    const unsigned long long t = static_cast<unsigned long long>(blockIdx.x) * tpc;
    if (t < g.total_tiles) run_tiles(g, t, min(t + tpc, g.total_tiles), f);
    return;
  }
  __shared__ unsigned int s_chunk;
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) s_chunk = atomicAdd(g.counter, 1u);
    __syncthreads();
    const unsigned long long t = static_cast<unsigned long long>(s_chunk) * tpc;
    if (t >= g.total_tiles) {
      leave_grid(g);
      return;
    }
    run_tiles(g, t, min(t + tpc, g.total_tiles), f);
  }
}

// Scheduler of the table kernel: a host-built work list instead of (chunk index -> binary search).
// Measured on the table kernel (4 GiB, one descriptor, round 1): every grab cost ~2 us -- 32 warps
// drain into the barrier, then three dependent trips through a memory system that is saturated with
// stores (work counter, list entry, descriptor: ~0.7 us each) before the first new store:
// 0.631 / 0.667 / 0.686 / 0.697 of the HBM roof with 0.5 / 1 / 2 / 4 MiB grabs.  So
//  * the host cuts every descriptor into grabs of its own (guided sizes: up to 4 MiB while there is
//    plenty of work left, down to one 256 KiB tile at the end, for balance);
//  * all but the last ~1/8 of the work is PRE-ASSIGNED: the host deals the large grabs out to the
//    CTAs (least loaded first) and CTA b walks its own segment of the list -- no counter, and the
//    index of the next grab is known when the current one starts.  Only the tail is handed out
//    dynamically, which is what keeps the CTAs finishing together;
//  * the NEXT grab is brought into shared memory while the current one is being written: thread 0
//    starts an asynchronous copy (cp.async: no destination register, nothing to stall on) of the
//    16-byte list entry when a grab starts and, one tile later, prefetches the 128-byte descriptor
//    it names into the SM's L1.  After the barrier at the end of the grab everything the next one
//    needs is a shared-memory read or an L1 hit away.  All of thread 0's scheduling state lives in shared memory: the hot
//    loop has no register to spare (64 per thread at 1024 threads).
#ifdef TDX_LUT_TIMELINE
// Measurement build only (benchmarks/lut_timeline.py): per-CTA timestamps of the table kernel.
// 16 slots per CTA: 0 enter, 1 first grab known, 2 first table built, 3 last grab done, 4 grabs,
// 5 ns thread 0 spent in barriers at grab ends, 6 ns from a grab's barrier to its first tile,
// 7 exit, 8 tiles, 9 table builds, 10 ns in table builds, 11 ns in prefetch_start/finish.
__device__ unsigned long long* g_lut_timeline = nullptr;
__device__ __forceinline__ unsigned long long tl_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define TL_SET(i, v) do { if (threadIdx.x == 0 && g_lut_timeline) g_lut_timeline[blockIdx.x * 16 + (i)] = (v); } while (0)
#define TL_ADD(i, v) do { if (threadIdx.x == 0 && g_lut_timeline) g_lut_timeline[blockIdx.x * 16 + (i)] += (v); } while (0)
#else
#define TL_SET(i, v) do { } while (0)
#define TL_ADD(i, v) do { } while (0)
#endif
struct LutSched {
  unsigned int next[2];  // list index of the grab in each slot (>= n_chunks: no more work)
  unsigned int pos, end; // CTA's segment of the pre-assigned part: next entry, one past the last
  volatile unsigned int stage;  // of the prefetch of the next grab: 2 = entry under way, 3 = descriptor under way / nothing to do
  unsigned int pad_[3];
  uint4 item[2];         // {descriptor, tiles, first tile in descriptor (lo, hi)}
  TdxInitDesc tab;       // the descriptor the table in shared memory was built for
};
static_assert(sizeof(TdxInitDesc) == 128, "a descriptor is one 128-byte line");
static_assert(sizeof(LutSched) % 16 == 0 && offsetof(LutSched, item) % 16 == 0, "LutSched follows the table in dynamic shared memory");

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(smem))), "l"(gmem)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Thread 0's side job during a grab (every function is called by thread 0 only).
__device__ __forceinline__ void lut_prefetch_start(const GroupArgs& g, LutSched* s, unsigned int slot) {
  unsigned int c = s->pos;
  if (c < s->end) {
    s->pos = c + 1u;
  } else {
    // dynamic tail: this thread waits for the counter here (~1 us, once per tail grab)
    c = g.n_static + atomicAdd(g.counter, 1u);
#ifdef TDX_LUT_TIMELINE
    if (g_lut_timeline && g_lut_timeline[blockIdx.x * 16 + 12] == 0) g_lut_timeline[blockIdx.x * 16 + 12] = tl_now();
#endif
  }
  s->next[slot] = c;
  if (c < g.n_chunks) {
    cp_async16(&s->item[slot], g.chunks + c);
    cp_async_commit();
    s->stage = 2u;
  } else {
    s->stage = 3u;
  }
}
__device__ __forceinline__ void lut_prefetch_desc(const GroupArgs& g, LutSched* s, unsigned int slot) {
  cp_async_wait_all();
  const uint32_t di = *reinterpret_cast<volatile uint32_t*>(&s->item[slot]);
  // (into this SM's L1, where the loads after the barrier find it; the descriptor table is 128-byte
  // aligned, so a descriptor is one line.  A copy into shared memory was tried: with the descriptor
  // behind a shared-memory reference ptxas no longer keeps the generator's constants in uniform
  // registers across the hot loop and reloads them from local memory for every group of vectors.)
  asm volatile("prefetch.global.L1 [%0];" ::"l"(g.descs + di));
  s->stage = 3u;
}
// hook of the tile loop (after the grab's first tile)
__device__ __forceinline__ void lut_prefetch_hook(const GroupArgs& g, LutSched* s, unsigned int slot) {
  if (s->stage == 2u) lut_prefetch_desc(g, s, slot);
}
__device__ __forceinline__ void lut_prefetch_finish(const GroupArgs& g, LutSched* s, unsigned int slot) {
  if (s->stage == 2u) lut_prefetch_desc(g, s, slot);  // (the grab was too short, or too odd, to overlap it)
  cp_async_wait_all();
}

// f(desc_index, descriptor, first tile, tiles, slot of the next grab)
//
// The SHAPE of this loop is load-bearing.  ptxas keeps the generator's constants (warp-reduction
// results) in uniform registers across the hot loop inside f only if it can prove the warps converged
// there, and it proves that from control flow whose conditions are shared-memory loads at
// thread-independent addresses, kernel parameters and counters derived from them.  A barrier-free
// walk of the pre-assigned share (list entries read by every thread from global memory, or per-warp
// positions, or `continue` past the barrier -- all tried) made it guard every REDUX with BRA.DIV and
// keep the constants in local memory, reloaded for every group of vectors: 0.65 instead of 0.75 of
// the roof on Llama-3-8B, every test green.  A convergence-friendly barrier-free form (the share
// copied to a shared-memory window, inner counted loop) compiles well but measured 0.4-0.8 % SLOWER
// than this one barrier per grab (profiles/r2_codegen_regression.md): with the next grab prefetched,
// the barrier costs less than warps that drift apart.  tests/test_lut_codegen.py guards the codegen.
template <class F>
__device__ __forceinline__ void for_each_listed_chunk(const GroupArgs& g, LutSched* s, F&& f) {
#ifdef TDX_LUT_TIMELINE
  if (threadIdx.x == 0 && g_lut_timeline)
    for (int i = 0; i < 16; ++i) g_lut_timeline[blockIdx.x * 16 + i] = 0;
  TL_SET(0, tl_now());
  unsigned long long tl_bar = 0;
#endif
  if (threadIdx.x == 0) {
    unsigned int lo = 0u, hi = 0u;
    if (g.n_static) {
      lo = __ldg(g.seg_off + blockIdx.x);
      hi = __ldg(g.seg_off + blockIdx.x + 1u);
    }
    s->pos = lo;
    s->end = hi;
    unsigned int c;
    if (lo < hi) {
      c = lo;
      s->pos = lo + 1u;
    } else {
      c = g.n_static + atomicAdd(g.counter, 1u);
    }
    s->next[0] = c;
    if (c < g.n_chunks) s->item[0] = __ldg(g.chunks + c);
  }
  __syncthreads();
  TL_SET(1, tl_now());
  for (unsigned int it = 0;; ++it) {
    const unsigned int slot = it & 1u;
    if (s->next[slot] >= g.n_chunks) {
#ifdef TDX_LUT_TIMELINE
      TL_SET(3, tl_bar);
      TL_SET(4, it);
#endif
      leave_grid(g);
      TL_SET(7, tl_now());
      return;
    }
    const uint4 e = s->item[slot];
#ifdef TDX_LUT_TIMELINE
    const unsigned long long tl_a = tl_now();
#endif
    if (threadIdx.x == 0) lut_prefetch_start(g, s, slot ^ 1u);
#ifdef TDX_LUT_TIMELINE
    const unsigned long long tl_b = tl_now();
    TL_ADD(11, tl_b - tl_a);
    TL_ADD(8, e.y);
#endif
    f(e.x, g.descs[e.x], static_cast<unsigned long long>(e.z) | (static_cast<unsigned long long>(e.w) << 32),
      static_cast<unsigned long long>(e.y), slot ^ 1u);
#ifdef TDX_LUT_TIMELINE
    const unsigned long long tl_c = tl_now();
#endif
    if (threadIdx.x == 0) lut_prefetch_finish(g, s, slot ^ 1u);
#ifdef TDX_LUT_TIMELINE
    const unsigned long long tl_d = tl_now();
    TL_ADD(11, tl_d - tl_c);
#endif
    __syncthreads();
#ifdef TDX_LUT_TIMELINE
    tl_bar = tl_now();
    TL_ADD(5, tl_bar - tl_d);
#endif
  }
}

// VECS = 16-byte vectors per thread and tile.  Measured (r1 sweep, 4 GiB): the fp32 generators gain
// 3-5 % from 16 independent vectors in flight, the 16-bit ones (more registers per vector) lose.
template <class Gen, int VECS = kVecsPerThread>
__global__ void __launch_bounds__(kThreads) tdx_rng_kernel(const __grid_constant__ GroupArgs g) {
  using Out = typename Gen::OutT;
  using T = OutTraits<Out>;
  constexpr int EPV = Gen::kEpv;
  constexpr int kVecsPerThread = VECS;
  constexpr int kTileVecs = kThreads * VECS;
  for_each_tile_run(g, [&](uint32_t di, unsigned long long tile0, unsigned long long ntiles) {
    const TdxInitDesc& d = *desc_of(g, di);
    const typename Gen::Params P = Gen::setup(d);
    const uint64_t begin = d.elem_begin, count = d.elem_count;
    const uint64_t gv0 = begin / EPV;
    const uint64_t nvec = (begin + count - 1) / EPV - gv0 + 1;
    char* const dst = static_cast<char*>(d.dst);
    // vector stores need: shard starts on a vector boundary and lands on a 16-byte address
    const bool aligned = (begin % EPV == 0) && ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0);
    const uint64_t nfull = aligned ? count / EPV : 0;  // vectors [0, nfull) are full + aligned
    for (unsigned long long tile = tile0; tile < tile0 + ntiles; ++tile) {
      const uint64_t base = tile * kTileVecs + threadIdx.x;
      // Inside a tile the upper 32 bits of the vector index (= Philox counter.y for the one-block
      // generators) are the same for every thread: handing them to the generator as a uniform value
      // makes round 1's second product and round 2's first product loop-invariant (18 IMAD.WIDE + 19
      // LOP3 per block instead of 20 + 20) and the per-vector index a 32-bit add.  A tile that
      // straddles a 2^32 boundary takes the edge path.
      const uint64_t gfirst = gv0 + (base - threadIdx.x);
      const uint32_t hi = static_cast<uint32_t>(gfirst >> 32);
      const bool one_hi = static_cast<uint32_t>((gfirst + (kTileVecs - 1)) >> 32) == hi;
      if (base - threadIdx.x + kTileVecs <= nfull && one_hi) {  // hot path: whole tile is full vectors
        const uint32_t lo0 = static_cast<uint32_t>(gfirst) + threadIdx.x;  // no carry: one_hi
        char* const p0 = dst + base * 16;
        if constexpr (HasGenVec<Gen>::value) {
          if (Gen::packed_ok(P)) {
#pragma unroll
            for (int i = 0; i < kVecsPerThread; ++i) {
              const uint64_t gv = (static_cast<uint64_t>(hi) << 32) | (lo0 + static_cast<uint32_t>(i) * kThreads);
              store_vec(p0 + static_cast<size_t>(i) * (kThreads * 16), Gen::gen_vec(P, gv));
            }
            continue;
          }
        }
#pragma unroll
        for (int i = 0; i < kVecsPerThread; ++i) {
          const uint64_t gv = (static_cast<uint64_t>(hi) << 32) | (lo0 + static_cast<uint32_t>(i) * kThreads);
          float v[EPV];
          Gen::gen(P, gv, v);
          store_vec(p0 + static_cast<size_t>(i) * (kThreads * 16), T::pack(v));
        }
      } else {  // ragged edge: partial vectors, unaligned shards, last tile
        for (int i = 0; i < kVecsPerThread; ++i) {
          const uint64_t j = base + static_cast<uint64_t>(i) * kThreads;
          if (j >= nvec) break;
          float v[EPV];
          Gen::gen(P, gv0 + j, v);
          if (j < nfull) {
            store_vec(dst + j * 16, T::pack(v));
          } else {
#pragma unroll
            for (int e = 0; e < EPV; ++e) {
              const uint64_t gidx = (gv0 + j) * EPV + e;
              if (gidx >= begin && gidx < begin + count) T::store_one(dst, gidx - begin, v[e]);
            }
          }
        }
      }
    }
  });
}

// constant fill: 16-byte pattern, frame = absolute 16-byte lines of the destination
__global__ void __launch_bounds__(kThreads) tdx_fill_kernel(const __grid_constant__ GroupArgs g) {
  for_each_tile_run(g, [&](uint32_t di, unsigned long long tile0, unsigned long long ntiles) {
    const TdxInitDesc& d = *desc_of(g, di);
    const uint4 pat = make_uint4(static_cast<uint32_t>(d.fill_bits[0]),
                                 static_cast<uint32_t>(d.fill_bits[0] >> 32),
                                 static_cast<uint32_t>(d.fill_bits[1]),
                                 static_cast<uint32_t>(d.fill_bits[1] >> 32));
    const int isz = d.dtype == TDX_F32 || d.dtype == TDX_RAW32 ? 4
                    : d.dtype == TDX_RAW64                    ? 8
                    : d.dtype == TDX_RAW8                     ? 1
                                                              : 2;
    const uintptr_t a = reinterpret_cast<uintptr_t>(d.dst);
    const uintptr_t end = a + d.elem_count * isz;
    const uintptr_t a0 = a & ~static_cast<uintptr_t>(15);
    const uint64_t nvec = (end - a0 + 15) / 16;
    for (unsigned long long tile = tile0; tile < tile0 + ntiles; ++tile) {
#pragma unroll
      for (int i = 0; i < kVecsPerThread; ++i) {
        const uint64_t j = tile * kTileVecs + static_cast<uint64_t>(i) * kThreads + threadIdx.x;
        if (j >= nvec) break;
        const uintptr_t p = a0 + j * 16;
        if (p >= a && p + 16 <= end) {
          store_vec(reinterpret_cast<void*>(p), pat);
        } else {  // head / tail line: byte-granular, the pattern phase follows the address
          const unsigned char* pb = reinterpret_cast<const unsigned char*>(&pat);
          for (int b = 0; b < 16; ++b) {
            const uintptr_t q = p + b;
            if (q >= a && q < end) *reinterpret_cast<unsigned char*>(q) = pb[(q - a) % isz];
          }
        }
      }
    }
  });
}


// One element of an index program: the integer, then (floating outputs) the epilogue in fp32, then
// -- TDX_BF16 / TDX_F16: `inv_freq.to(torch.bfloat16)`, what Module.to(dtype) does to a rotary
// buffer -- one rounding to nearest even at the store.
__device__ __forceinline__ uint64_t iota_itemsize(int dtype) { return dtype == TDX_I64 ? 8 : dtype == TDX_F32 ? 4 : 2; }
__device__ __forceinline__ void iota_store(const TdxInitDesc& d, const EpiParams& epi, uint64_t j, long long val) {
  if (d.dtype == TDX_I64) {
    static_cast<long long*>(d.dst)[j] = val;
    return;
  }
  float v = static_cast<float>(val);
  if (epi.n) v = apply_epi<float>(epi, v);
  if (d.dtype == TDX_F32) static_cast<float*>(d.dst)[j] = v;
  else if (d.dtype == TDX_BF16) static_cast<__nv_bfloat16*>(d.dst)[j] = __float2bfloat16_rn(v);
  else static_cast<__half*>(d.dst)[j] = __float2half_rn(v);
}

// index programs: element g = start + g * step (integers), then the epilogue (in fp32).  Tiny
// buffers (rotary inv_freq: 64 elements, position ids: a few K): nothing to optimise but the launch
// they no longer need -- they ride in the module's plan like any other descriptor.
__global__ void __launch_bounds__(kThreads) tdx_iota_kernel(const __grid_constant__ GroupArgs g) {
  for_each_tile_run(g, [&](uint32_t di, unsigned long long tile0, unsigned long long ntiles) {
    const TdxInitDesc& d = *desc_of(g, di);
    const long long start = static_cast<long long>(d.p0), step = static_cast<long long>(d.p1);
    const EpiParams epi = load_epi(d);
    const uint64_t per_vec = 16 / iota_itemsize(d.dtype);
    // (frame: element i of the descriptor is global element elem_begin + i; one "vector" = 16 bytes)
    for (unsigned long long tile = tile0; tile < tile0 + ntiles; ++tile) {
      for (int i = 0; i < kVecsPerThread; ++i) {
        const uint64_t vec = tile * kTileVecs + static_cast<uint64_t>(i) * kThreads + threadIdx.x;
        for (uint64_t e = 0; e < per_vec; ++e) {
          const uint64_t j = vec * per_vec + e;
          if (j >= d.elem_count) break;
          iota_store(d, epi, j, start + static_cast<long long>(d.elem_begin + j) * step);
        }
      }
    }
  });
}

// ---------------------------------------------------------------------------------------------
// 16-bit outputs through a shared-memory table
// ---------------------------------------------------------------------------------------------
// A 16-bit normal or uniform is a pure function of the Philox half-word k in [0, 65535] once the
// descriptor's parameters are fixed, so a CTA evaluates it once per k into a 128 KiB shared-memory
// table and then turns most half-words into outputs with one LDS.U16 each.  The table is built by
// the SAME device function as the direct kernel, so the two are bit-identical
// (tests/test_kernels_gpu.py checks it); the host picks this kernel for large descriptors only,
// because the table costs 65536 evaluations per CTA.
#ifndef TDX_LUT_THREADS
#define TDX_LUT_THREADS 1024
#endif
#ifndef TDX_LUT_VECS
#define TDX_LUT_VECS 16
#endif
constexpr int kLutThreads = TDX_LUT_THREADS;
constexpr int kLutVecsPerThread = TDX_LUT_VECS;
constexpr int kLutTileVecs = kLutThreads * kLutVecsPerThread;  // bytes per tile = 16 x this
constexpr int kLutTilesPerChunk = (1 << 20) / (kLutTileVecs * 16) > 0 ? (1 << 20) / (kLutTileVecs * 16) : 1;
#ifndef TDX_LUT_MAX_CHUNK_LOG2
#define TDX_LUT_MAX_CHUNK_LOG2 22
#endif
constexpr unsigned long long kLutMaxTilesPerChunk =
    (1ull << TDX_LUT_MAX_CHUNK_LOG2) / (kLutTileVecs * 16) > 0 ? (1ull << TDX_LUT_MAX_CHUNK_LOG2) / (kLutTileVecs * 16) : 1;
constexpr uint32_t kLutBytes = 65536u * 2u;
// A launch uses the table kernel when it holds at least this many table-eligible elements
// (TDX_LUT_MIN_LAUNCH_ELEMS overrides): the table costs every CTA ~4 us.
constexpr uint64_t kLutMinLaunchElemsDefault = 1ull << 25;  // 64 MB of 16-bit output (normal; x2 for the uniform)
constexpr unsigned short kLutSentinel = 0xffffu;  // a NaN pattern in bf16 and fp16: never a value

// cold path of the table kernel (a vector that contains k == 0): out of line, to keep the hot loop
// small enough for the instruction cache
template <class Gen>
__device__ __noinline__ uint4 lut_slow_vector(const TdxInitDesc* d, uint64_t gv) {
  using Out = typename Gen::OutT;
  const typename Gen::Params P = Gen::setup(*d);
  float v[8];
  Gen::gen(P, gv, v);
  return OutTraits<Out>::pack(v);
}

// ---------------------------------------------------------------------------------------------
// the table kernel
// ---------------------------------------------------------------------------------------------
// What shaped it (ncu of its first version, profiles/r1_ncu_bench_llama3_8b_lut_kernel.json: 93
// instructions per 8-element vector, LSU wavefronts 88 % of peak; and benchmarks/instr_rate.cu):
//  * byte address of a table entry = 2*k in ONE instruction: IDP.2A (dp2a.lo with the byte pair
//    (2,0) or (0,2) selects a half-word and doubles it) instead of PRMT + IADD3; the table's base
//    is the LDS immediate;
//  * counter.y (blk >> 32) is uniform inside a tile, so Philox round 1's second product and round
//    2's first product are loop-invariant: 18 IMAD.WIDE + 19 LOP3 per block instead of 20 + 20;
//    the round keys are kernel parameters (constant bank operands of the LOP3s);
//  * the normal's k == 0 sentinel (NaN) is accumulated with 2 HFMA2 per vector (NaN survives
//    a*b+c) and tested once per GROUP vectors; a hit re-derives the group's Philox blocks and
//    repairs only the vectors that really contain a zero half-word;
//  * LUT_ELEMS of the 8 elements of a vector go through the table, the others are computed (same
//    device function as the table builder => same bits): random indices cost 3.56 shared-memory
//    wavefronts per LDS and an SM delivers one wavefront per cycle, so a pure table kernel is
//    LSU-bound at 0.63 of the HBM roof;
//  * work comes from a host-built list of descriptor-pure grabs (for_each_listed_chunk).
#ifndef TDX_LUT_ELEMS_A
#define TDX_LUT_ELEMS_A 7
#endif
#ifndef TDX_LUT_ELEMS_B
#define TDX_LUT_ELEMS_B 7
#endif
#ifndef TDX_LUT_GROUP
#define TDX_LUT_GROUP 4
#endif
#ifndef TDX_LUT_PACK
#define TDX_LUT_PACK 0
#endif

// byte offset 2*k of the table entry of the low / high half-word of a Philox word, in one IDP.2A
__device__ __forceinline__ uint32_t lut_off_lo(uint32_t w) {
  uint32_t a;
  asm("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(a) : "r"(w), "r"(0x00000002u), "r"(0u));
  return a;
}
__device__ __forceinline__ uint32_t lut_off_hi(uint32_t w) {
  uint32_t a;
  asm("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(a) : "r"(w), "r"(0x00000200u), "r"(0u));
  return a;
}
// Dynamic shared memory of a kernel without static shared memory starts at this offset of the
// shared window on sm_100 (the first KiB is reserved by the system).  The table kernel folds it
// into the immediate field of its LDS instructions; it checks the assumption at run time and
// takes the generic path if it ever does not hold.
constexpr uint32_t kDynSmemBase = 1024;
__device__ __forceinline__ uint32_t lds_u16_tab(uint32_t off) {
  unsigned short v;
  asm volatile("ld.shared.u16 %0, [%1+1024];" : "=h"(v) : "r"(off));
  return v;
}
__device__ __forceinline__ uint32_t pack_u16(uint32_t lo, uint32_t hi) {
#if TDX_LUT_PACK == 1
  return __byte_perm(lo, hi, 0x5410);
#elif TDX_LUT_PACK == 2
  return lo | (hi << 16);
#else
  uint32_t r;
  asm("mad.lo.u32 %0, %1, 65536, %2;" : "=r"(r) : "r"(hi), "r"(lo));
  return r;
#endif
}
// exact test for "some 16-bit half of w is zero"
__device__ __forceinline__ bool has_zero_half(uint32_t w) {
  return ((w - 0x00010001u) & ~w & 0x80008000u) != 0u;
}

template <int ROUNDS>
__device__ __forceinline__ uint4 philox4x32_rk(uint4 c, const uint32_t (&rk)[20]) {
#pragma unroll
  for (int r = 0; r < ROUNDS; ++r) {
    unsigned long long p0, p1;
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(p0) : "r"(c.x), "r"(kPhiloxM0));
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(p1) : "r"(c.z), "r"(kPhiloxM1));
    const uint32_t hi0 = static_cast<uint32_t>(p0 >> 32), lo0 = static_cast<uint32_t>(p0);
    const uint32_t hi1 = static_cast<uint32_t>(p1 >> 32), lo1 = static_cast<uint32_t>(p1);
    c = make_uint4(hi1 ^ c.y ^ rk[2 * r], lo1, hi0 ^ c.w ^ rk[2 * r + 1], lo0);
  }
  return c;
}

template <class Out>
struct NanAcc;
template <>
struct NanAcc<__nv_bfloat16> {
  uint32_t acc = 0;
  __device__ __forceinline__ void add(uint32_t a, uint32_t b) {
    asm("fma.rn.bf16x2 %0, %1, %2, %0;" : "+r"(acc) : "r"(a), "r"(b));
  }
};
template <>
struct NanAcc<__half> {
  uint32_t acc = 0;
  __device__ __forceinline__ void add(uint32_t a, uint32_t b) {
    asm("fma.rn.f16x2 %0, %1, %2, %0;" : "+r"(acc) : "r"(a), "r"(b));
  }
};

// What the table kernel needs to know about a 16-bit generator: the value of half-word k (the table
// entry and, for the elements that are not looked up, the directly computed value -- one device
// function, hence the same bits), the parameters that identify a table, and whether some k needs the
// exact out-of-line path (the normal's k == 0 tail).
// Two descriptors of one table-kernel family share a table iff their generator parameters and
// epilogues are equal (dtype, source and "has an epilogue" are the family's).
__device__ __forceinline__ bool same_table(const TdxInitDesc& a, const TdxInitDesc& b) {
  bool same = a.p0 == b.p0 && a.p1 == b.p1 && a.n_epi == b.n_epi && a.reserved == b.reserved;
  for (int i = 0; i < TDX_MAX_EPI; ++i)
    if (i < a.n_epi)
      same = same && a.epi[i].op == b.epi[i].op && a.epi[i].a == b.epi[i].a && a.epi[i].b == b.epi[i].b;
  return same;
}

// EPI: the descriptor carries epilogue steps (trunc_normal_, randn * s + m, ...).  The table then
// holds the FINAL value for every half-word -- the epilogue is as elementwise as the source
// transform -- so an erfinv costs the same as nothing once the table is built.
template <class Out, int R, bool EPI = false>
struct TabNormal {
  using Gen = GenNormalICDF16<Out, R, EPI>;
  using Params = typename Gen::Params;
  static constexpr bool kHasTail = true;
  // The grid x = k/32768 - 1 is symmetric about k = 32768 and, with mean == +0 and no epilogue, so is
  // the value (fma(q, -x, +0) == -fma(q, x, +0), rounding included): the upper half of the table is
  // the lower half with the sign bit flipped -- half the evaluations per table build.
  __device__ static __forceinline__ bool mirrored(const Params& p) { return !EPI && __float_as_uint(p.mean) == 0u; }
  __device__ static __forceinline__ float value(const Params& p, float magic) {
    float t;
    const float v = Gen::element(p, magic, t);  // k == 0: +-inf or NaN (lg2(0) = -inf)
    return EPI ? apply_epi<Out>(p.epi, v) : v;
  }
  // the same value from k as a float (hot loop: halfword_as_float)
  __device__ static __forceinline__ float value_kf(const Params& p, float kf) {
    const float v = Gen::element_kf(p, kf);
    return EPI ? apply_epi<Out>(p.epi, v) : v;
  }
  __device__ static __forceinline__ void uniformize(Params& p) {
    p.mean = __uint_as_float(__reduce_or_sync(0xffffffffu, __float_as_uint(p.mean)));
    p.c0 = __uint_as_float(__reduce_or_sync(0xffffffffu, __float_as_uint(p.c0)));
    p.c1 = __uint_as_float(__reduce_or_sync(0xffffffffu, __float_as_uint(p.c1)));
    p.c2 = __uint_as_float(__reduce_or_sync(0xffffffffu, __float_as_uint(p.c2)));
    p.c3 = __uint_as_float(__reduce_or_sync(0xffffffffu, __float_as_uint(p.c3)));
    p.c4 = __uint_as_float(__reduce_or_sync(0xffffffffu, __float_as_uint(p.c4)));
    p.c5 = __uint_as_float(__reduce_or_sync(0xffffffffu, __float_as_uint(p.c5)));
  }
};
template <class Out, int R, bool EPI = false>
struct TabUniform {
  using Gen = GenUniform16<Out, R, EPI>;
  using Params = typename Gen::Params;
  static constexpr bool kHasTail = false;
  __device__ static __forceinline__ bool mirrored(const Params&) { return false; }
  __device__ static __forceinline__ float value(const Params& p, float magic) {
    const float x = fminf(fmaf(magic - 8388608.0f, p.scale, p.from), p.to_prev);  // == Gen::gen, element by element
    return EPI ? apply_epi<Out>(p.epi, x) : x;
  }
  __device__ static __forceinline__ float value_kf(const Params& p, float kf) {
    const float x = fminf(fmaf(kf, p.scale, p.from), p.to_prev);  // magic - 2^23 == k exactly
    return EPI ? apply_epi<Out>(p.epi, x) : x;
  }
  __device__ static __forceinline__ void uniformize(Params& p) {
    p.from = __uint_as_float(__reduce_or_sync(0xffffffffu, __float_as_uint(p.from)));
    p.scale = __uint_as_float(__reduce_or_sync(0xffffffffu, __float_as_uint(p.scale)));
    p.to_prev = __uint_as_float(__reduce_or_sync(0xffffffffu, __float_as_uint(p.to_prev)));
  }
};

// Tiles that are not made of full, aligned vectors only (unaligned shards, ragged ends) or that
// straddle a 2^32-block boundary: out of line and self-contained (everything is re-derived from the
// descriptor), so that none of its state is live across the hot loop of the table kernel.
template <class Gen>
__device__ __noinline__ void lut_ragged_tile(const TdxInitDesc* dp, unsigned long long tile) {
  using Out = typename Gen::OutT;
  using T = OutTraits<Out>;
  const TdxInitDesc& d = *dp;
  const typename Gen::Params P = Gen::setup(d);
  const uint64_t begin = d.elem_begin, count = d.elem_count;
  const uint64_t gv0 = begin / 8;
  const uint64_t nvec = (begin + count - 1) / 8 - gv0 + 1;
  char* const dst = static_cast<char*>(d.dst);
  const bool aligned = (begin % 8 == 0) && ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0);
  const uint64_t nfull = aligned ? count / 8 : 0;
  const uint64_t base = tile * kLutTileVecs + threadIdx.x;
  for (int i = 0; i < kLutVecsPerThread; ++i) {
    const uint64_t j = base + static_cast<uint64_t>(i) * kLutThreads;
    if (j >= nvec) break;
    float v[8];
    Gen::gen(P, gv0 + j, v);
    if (j < nfull) {
      store_vec(dst + j * 16, T::pack(v));
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const uint64_t gidx = (gv0 + j) * 8 + e;
        if (gidx >= begin && gidx < begin + count) T::store_one(dst, gidx - begin, v[e]);
      }
    }
  }
}

// Constant-fill descriptors ride along in the table kernel's work list (the host folds a module's
// few small fills -- norm weights, biases -- into it instead of paying a kernel launch for them).
__device__ __noinline__ void lut_fill_tiles(const TdxInitDesc* dp, unsigned long long tile0,
                                            unsigned long long ntiles) {
  const TdxInitDesc& d = *dp;
  const uint4 pat = make_uint4(static_cast<uint32_t>(d.fill_bits[0]), static_cast<uint32_t>(d.fill_bits[0] >> 32),
                               static_cast<uint32_t>(d.fill_bits[1]), static_cast<uint32_t>(d.fill_bits[1] >> 32));
  const int isz = d.dtype == TDX_F32 || d.dtype == TDX_RAW32 ? 4 : d.dtype == TDX_RAW64 ? 8 : d.dtype == TDX_RAW8 ? 1 : 2;
  const uintptr_t a = reinterpret_cast<uintptr_t>(d.dst);
  const uintptr_t end = a + d.elem_count * isz;
  const uintptr_t a0 = a & ~static_cast<uintptr_t>(15);
  const uint64_t nvec = (end - a0 + 15) / 16;
  for (unsigned long long tile = tile0; tile < tile0 + ntiles; ++tile) {
    for (int i = 0; i < kLutVecsPerThread; ++i) {
      const uint64_t j = tile * kLutTileVecs + static_cast<uint64_t>(i) * kLutThreads + threadIdx.x;
      if (j >= nvec) break;
      const uintptr_t p = a0 + j * 16;
      if (p >= a && p + 16 <= end) {
        store_vec(reinterpret_cast<void*>(p), pat);
      } else {  // head / tail line: byte-granular, the pattern phase follows the address
        const unsigned char* pb = reinterpret_cast<const unsigned char*>(&pat);
        for (int b = 0; b < 16; ++b) {
          const uintptr_t q = p + b;
          if (q >= a && q < end) *reinterpret_cast<unsigned char*>(q) = pb[(q - a) % isz];
        }
      }
    }
  }
}

// ... and so do index programs (rotary inv_freq, position ids: a few hundred bytes).
__device__ __noinline__ void lut_iota_tiles(const TdxInitDesc* dp, unsigned long long tile0, unsigned long long ntiles) {
  const TdxInitDesc& d = *dp;
  const long long start = static_cast<long long>(d.p0), step = static_cast<long long>(d.p1);
  const EpiParams epi = load_epi(d);
  const uint64_t per_vec = 16 / iota_itemsize(d.dtype);
  for (unsigned long long tile = tile0; tile < tile0 + ntiles; ++tile) {
    for (int i = 0; i < kLutVecsPerThread; ++i) {
      const uint64_t vec = tile * kLutTileVecs + static_cast<uint64_t>(i) * kLutThreads + threadIdx.x;
      if (vec * per_vec >= d.elem_count) break;
      for (uint64_t e = 0; e < per_vec; ++e) {
        const uint64_t j = vec * per_vec + e;
        if (j >= d.elem_count) break;
        iota_store(d, epi, j, start + static_cast<long long>(d.elem_begin + j) * step);
      }
    }
  }
}

// Exponent-all-ones test of either half of a packed pair (inf or NaN).
template <class Out>
__device__ __forceinline__ bool any_nonfinite2(uint32_t v) {
  constexpr uint32_t m = sizeof(Out) == 2 && OutTraits<Out>::kDtype == TDX_BF16 ? 0x7f80u : 0x7c00u;
  return ((v & m) == m) || ((v & (m << 16)) == (m << 16));
}

// LUT_A / LUT_B: number of the 8 elements of an even / odd vector that go through the table (the
// others are computed).  PKEYS: Philox round keys are kernel parameters (the group's descriptors
// share one seed).  Issue-cost model (benchmarks/instr_rate.cu: IMAD.WIDE, LOP3, PRMT, IDP, IMAD,
// F2FP, HFMA2 hold the dispatch port 2 cycles, FFMA/FMNMX/LDS 1, and IMAD.WIDE its pipe 4):
// Philox 74 of the ~126 cycles a vector takes, a table element 4 (IDP + LDS + half a pack), a
// computed normal 14, a computed uniform 6 -- against 3.56 shared-memory wavefronts per looked-up
// element, of which an SM delivers one per cycle.
template <class Tab, class Out, int R, bool PKEYS, int LUT_A, int LUT_B, int GROUP = TDX_LUT_GROUP>
__global__ void __launch_bounds__(kLutThreads, 1) tdx_lut16_kernel(const GroupArgs g) {
  using Gen = typename Tab::Gen;
  using T = OutTraits<Out>;
  static_assert(kLutVecsPerThread % GROUP == 0, "GROUP must divide the vectors per thread");
  // The table is the first thing in shared memory (this kernel has no static shared memory; the
  // scheduler's two slots follow the table), so the byte address of entry k is 2*k plus a constant
  // that fits the LDS immediate, and the IDP that forms 2*k reads one vector register.
  extern __shared__ __align__(16) unsigned short lut[];
  LutSched* const sched = reinterpret_cast<LutSched*>(lut + 65536);
  const bool base_ok = static_cast<uint32_t>(__cvta_generic_to_shared(lut)) == kDynSmemBase;
  uint32_t have_di = 0xffffffffu;  // descriptor the table in shared memory was built for (its copy: sched->tab)
  const bool t0 = threadIdx.x == 0;
  for_each_listed_chunk(g, sched, [&](uint32_t di, const TdxInitDesc& d, unsigned long long tile0,
                                      unsigned long long ntiles, unsigned int next_slot) {
    if (d.src == TDX_SRC_CONST) {  // a fill folded into this launch (build_plan): the table stays as it is
      lut_fill_tiles(&d, tile0, ntiles);
      return;
    }
    if (d.src == TDX_SRC_IOTA) {
      lut_iota_tiles(&d, tile0, ntiles);
      return;
    }
    typename Gen::Params P = Gen::setup(d);
    // Loop-invariant scalars that come out of a global load: a warp reduction's result lives in a
    // uniform register by construction, so the FFMAs / LOP3s that use them read two vector
    // registers instead of three (measured: +2.5 %, fewer dispatch stalls).
    Tab::uniformize(P);
    P.ph.cz = __reduce_or_sync(0xffffffffu, P.ph.cz);
    P.ph.cw = __reduce_or_sync(0xffffffffu, P.ph.cw);
#ifdef TDX_LUT_TIMELINE
    const unsigned long long tl_t0 = tl_now();
#endif
    if (have_di == 0xffffffffu || (di != have_di && !same_table(d, sched->tab))) {
      __syncthreads();  // everyone is done reading the old table (and comparing with its descriptor)
      if (threadIdx.x < 8)
        reinterpret_cast<uint4*>(&sched->tab)[threadIdx.x] = reinterpret_cast<const uint4*>(&d)[threadIdx.x];
      if (Tab::mirrored(P)) {
        for (uint32_t k = threadIdx.x; k <= 32768u; k += kLutThreads) {
          const Out o = static_cast<Out>(Tab::value(P, __uint_as_float(0x4b000000u | k)));
          const unsigned short bits = *reinterpret_cast<const unsigned short*>(&o);
          lut[k] = (Tab::kHasTail && k == 0) ? kLutSentinel : bits;
          if (k != 0u && k != 32768u) lut[65536u - k] = bits ^ 0x8000u;
        }
      } else {
        for (uint32_t k = threadIdx.x; k < 65536u; k += kLutThreads) {
          const Out o = static_cast<Out>(Tab::value(P, __uint_as_float(0x4b000000u | k)));
          lut[k] = (Tab::kHasTail && k == 0) ? kLutSentinel : *reinterpret_cast<const unsigned short*>(&o);
        }
      }
      __syncthreads();
#ifdef TDX_LUT_TIMELINE
      TL_ADD(9, 1);
      TL_ADD(10, tl_now() - tl_t0);
      if (have_di == 0xffffffffu) TL_SET(2, tl_now());
#endif
    }
    have_di = di;
    const uint64_t begin = d.elem_begin;
    const uint64_t gv0 = begin / 8;
    char* const dst = static_cast<char*>(d.dst);
    const bool aligned = (begin % 8 == 0) && ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0);
    const uint64_t nfull = aligned ? d.elem_count / 8 : 0;
    TL_ADD(6, tl_now() - tl_t0);
    for (unsigned long long tile = tile0; tile < tile0 + ntiles; ++tile) {
      const uint64_t tbase = tile * kLutTileVecs;  // first vector of the tile (descriptor-relative)
      const uint64_t gfirst = gv0 + tbase;         // its global block index
      const uint32_t blk_hi = static_cast<uint32_t>(gfirst >> 32);
      const bool one_hi = static_cast<uint32_t>((gfirst + (kLutTileVecs - 1)) >> 32) == blk_hi;
      if (tbase + kLutTileVecs <= nfull && one_hi && base_ok) {
        const uint32_t lo0 = static_cast<uint32_t>(gfirst) + threadIdx.x;  // no carry: one_hi
        char* const p0 = dst + (tbase + threadIdx.x) * 16;
#pragma unroll
        for (int i0 = 0; i0 < kLutVecsPerThread; i0 += GROUP) {
          NanAcc<Out> nan;
#pragma unroll
          for (int i = i0; i < i0 + GROUP; ++i) {
            const int LUT_ELEMS = (i & 1) ? LUT_B : LUT_A;
            const uint4 c = make_uint4(P.ph.cz, blk_hi, P.ph.cw, lo0 + static_cast<uint32_t>(i) * kLutThreads);
            const uint4 w = PKEYS ? philox4x32_rk<R>(c, g.rk) : philox4x32<R>(c, P.ph.k0, P.ph.k1);
            const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
            uint32_t r[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const bool lo_lut = 2 * q < LUT_ELEMS, hi_lut = 2 * q + 1 < LUT_ELEMS;
              if (lo_lut && hi_lut) {
                r[q] = pack_u16(lds_u16_tab(lut_off_lo(ws[q])), lds_u16_tab(lut_off_hi(ws[q])));
              } else if (!lo_lut && !hi_lut) {
                r[q] = T::pack2(Tab::value_kf(P, halfword_as_float(w, 2 * q)),
                                Tab::value_kf(P, halfword_as_float(w, 2 * q + 1)));
              } else {  // low half from the table, high half computed
                const uint32_t hi16 = T::pack2(0.0f, Tab::value_kf(P, halfword_as_float(w, 2 * q + 1)));
                r[q] = (hi16 & 0xffff0000u) | lds_u16_tab(lut_off_lo(ws[q]));
              }
            }
            if (Tab::kHasTail) {
              nan.add(r[0], r[1]);
              nan.add(r[2], r[3]);
            }
            store_vec(p0 + static_cast<size_t>(i) * (kLutThreads * 16), make_uint4(r[0], r[1], r[2], r[3]));
          }
          // k == 0 reads the NaN sentinel from the table and makes a computed element +-inf or NaN;
          // either survives the a*b+c accumulation.  Rare (2^-16 per element; also after an
          // overflow of the accumulator, which only costs time): find the vector(s) that contain
          // a zero half-word and redo them exactly.
          if (Tab::kHasTail && any_nonfinite2<Out>(nan.acc)) {
#pragma unroll 1
            for (int i = i0; i < i0 + GROUP; ++i) {
              const uint64_t gv = gfirst + threadIdx.x + static_cast<uint64_t>(i) * kLutThreads;
              const uint4 w = philox_block<R>(P.ph, gv);
              if (has_zero_half(w.x) || has_zero_half(w.y) || has_zero_half(w.z) || has_zero_half(w.w))
                store_vec(p0 + static_cast<size_t>(i) * (kLutThreads * 16), lut_slow_vector<Gen>(&d, gv));
            }
          }
        }
      } else {
        lut_ragged_tile<Gen>(&d, tile);
      }
      // (thread 0, after the grab's first tile: the next list entry has arrived -- start on its descriptor)
      if (t0 && tile == tile0) lut_prefetch_hook(g, sched, next_slot);
    }
  });
}

// ---------------------------------------------------------------------------------------------
// host side: validation, grouping into kernel families, plan upload, launch
// ---------------------------------------------------------------------------------------------
using KernelFn = void (*)(const GroupArgs);

struct Family {
  int src, dtype, algo /*resolved, without flags*/, rounds, epi;
  KernelFn fn;
  const char* name;
  int threads = kThreads;
  int tile_vecs = kTileVecs;
  int dyn_smem = 0;
  bool lut = false;
  int tiles_per_chunk = kTilesPerChunk;
  KernelFn fn_any_seed = nullptr;  // twin of `fn` for groups whose descriptors do not share one seed
};

#define TDX_FAM(src, dt, algo, rounds, epi, ...) \
  { src, dt, algo, rounds, epi, static_cast<KernelFn>(tdx_rng_kernel<__VA_ARGS__>), #__VA_ARGS__ }
// fp32 generators: 16 vectors per thread (tile = 64 KiB, 4 tiles per 256 KiB grab)
#define TDX_FAM_V16(src, dt, algo, rounds, epi, ...)                                               \
  { src, dt, algo, rounds, epi, static_cast<KernelFn>(tdx_rng_kernel<__VA_ARGS__, 16>), #__VA_ARGS__, \
    kThreads, kThreads * 16, 0, false, (kVecsPerThread * kTilesPerChunk) / 16 }
#ifndef TDX_LUT_PKEYS
#define TDX_LUT_PKEYS 1
#endif
#ifndef TDX_LUT_UNIFORM_ELEMS
#define TDX_LUT_UNIFORM_ELEMS 6  // measured at 4 GiB bf16: 8/7/6/5/4 -> 0.674/0.748/0.815/0.813/0.803 of the HBM roof
#endif
// table-driven twins of the 16-bit generators: `fn` takes the Philox round keys from the kernel
// parameters (groups that share one seed: the normal case), `fn_any_seed` from each descriptor
#define TDX_FAM_LUT(src, dt, algo, epi, name, out, la, lb, ...)                                          \
  { src, dt, algo, 10, epi,                                                                           \
    static_cast<KernelFn>(tdx_lut16_kernel<__VA_ARGS__, out, 10, TDX_LUT_PKEYS != 0, la, lb>), name,   \
    kLutThreads, kLutTileVecs, static_cast<int>(kLutBytes + sizeof(LutSched)), true, kLutTilesPerChunk, \
    static_cast<KernelFn>(tdx_lut16_kernel<__VA_ARGS__, out, 10, false, la, lb>) }

using bf16 = __nv_bfloat16;
using f16 = __half;

static const Family kFamilies[] = {
    {TDX_SRC_CONST, -1, 0, 0, 0, tdx_fill_kernel, "fill"},
    {TDX_SRC_IOTA, -1, 0, 0, 0, tdx_iota_kernel, "iota"},
    // shipped defaults
    TDX_FAM_V16(TDX_SRC_UNIFORM, TDX_F32, 0, 10, 0, GenUniform32<float, 10, false>),
    TDX_FAM(TDX_SRC_UNIFORM, TDX_F32, 0, 10, 1, GenUniform32<float, 10, true>),
    TDX_FAM(TDX_SRC_UNIFORM, TDX_BF16, 0, 10, 0, GenUniform16<bf16, 10, false>),
    TDX_FAM(TDX_SRC_UNIFORM, TDX_BF16, 0, 10, 1, GenUniform16<bf16, 10, true>),
    TDX_FAM(TDX_SRC_UNIFORM, TDX_F16, 0, 10, 0, GenUniform16<f16, 10, false>),
    TDX_FAM(TDX_SRC_UNIFORM, TDX_F16, 0, 10, 1, GenUniform16<f16, 10, true>),
    TDX_FAM_V16(TDX_SRC_NORMAL, TDX_F32, TDX_ALGO_BM32, 10, 0, GenNormalBM32<float, 10, false>),
    TDX_FAM(TDX_SRC_NORMAL, TDX_F32, TDX_ALGO_BM32, 10, 1, GenNormalBM32<float, 10, true>),
    TDX_FAM(TDX_SRC_NORMAL, TDX_BF16, TDX_ALGO_ICDF16, 10, 0, GenNormalICDF16<bf16, 10, false>),
    TDX_FAM(TDX_SRC_NORMAL, TDX_BF16, TDX_ALGO_ICDF16, 10, 1, GenNormalICDF16<bf16, 10, true>),
    TDX_FAM(TDX_SRC_NORMAL, TDX_F16, TDX_ALGO_ICDF16, 10, 0, GenNormalICDF16<f16, 10, false>),
    TDX_FAM(TDX_SRC_NORMAL, TDX_F16, TDX_ALGO_ICDF16, 10, 1, GenNormalICDF16<f16, 10, true>),
    // "wide" 16-bit outputs (TDX_ALGO_WIDE32): the fp32 stream and arithmetic, rounded once at the
    // store -- what `t_fp32.uniform_()/normal_()` followed by `.to(bf16/fp16)` means
    TDX_FAM(TDX_SRC_UNIFORM, TDX_BF16, TDX_ALGO_WIDE32, 10, 0, GenUniform32<bf16, 10, false>),
    TDX_FAM(TDX_SRC_UNIFORM, TDX_BF16, TDX_ALGO_WIDE32, 10, 1, GenUniform32<bf16, 10, true>),
    TDX_FAM(TDX_SRC_UNIFORM, TDX_F16, TDX_ALGO_WIDE32, 10, 0, GenUniform32<f16, 10, false>),
    TDX_FAM(TDX_SRC_UNIFORM, TDX_F16, TDX_ALGO_WIDE32, 10, 1, GenUniform32<f16, 10, true>),
    TDX_FAM(TDX_SRC_NORMAL, TDX_BF16, TDX_ALGO_BM32, 10, 0, GenNormalBM32<bf16, 10, false>),
    TDX_FAM(TDX_SRC_NORMAL, TDX_BF16, TDX_ALGO_BM32, 10, 1, GenNormalBM32<bf16, 10, true>),
    TDX_FAM(TDX_SRC_NORMAL, TDX_F16, TDX_ALGO_BM32, 10, 0, GenNormalBM32<f16, 10, false>),
    TDX_FAM(TDX_SRC_NORMAL, TDX_F16, TDX_ALGO_BM32, 10, 1, GenNormalBM32<f16, 10, true>),
    // table-driven twins of the 16-bit normal for large descriptors (bit-identical output)
    TDX_FAM_LUT(TDX_SRC_NORMAL, TDX_BF16, TDX_ALGO_ICDF16, 0, "lut<normal, bf16>", bf16, TDX_LUT_ELEMS_A, TDX_LUT_ELEMS_B, TabNormal<bf16, 10>),
    TDX_FAM_LUT(TDX_SRC_NORMAL, TDX_F16, TDX_ALGO_ICDF16, 0, "lut<normal, f16>", f16, TDX_LUT_ELEMS_A, TDX_LUT_ELEMS_B, TabNormal<f16, 10>),
    TDX_FAM_LUT(TDX_SRC_UNIFORM, TDX_BF16, 0, 0, "lut<uniform, bf16>", bf16, TDX_LUT_UNIFORM_ELEMS, TDX_LUT_UNIFORM_ELEMS, TabUniform<bf16, 10>),
    TDX_FAM_LUT(TDX_SRC_UNIFORM, TDX_F16, 0, 0, "lut<uniform, f16>", f16, TDX_LUT_UNIFORM_ELEMS, TDX_LUT_UNIFORM_ELEMS, TabUniform<f16, 10>),
    // with an epilogue every element is looked up: the computed alternative may contain an erfinv
    TDX_FAM_LUT(TDX_SRC_NORMAL, TDX_BF16, TDX_ALGO_ICDF16, 1, "lut<normal+epilogue, bf16>", bf16, 8, 8, TabNormal<bf16, 10, true>),
    TDX_FAM_LUT(TDX_SRC_NORMAL, TDX_F16, TDX_ALGO_ICDF16, 1, "lut<normal+epilogue, f16>", f16, 8, 8, TabNormal<f16, 10, true>),
    TDX_FAM_LUT(TDX_SRC_UNIFORM, TDX_BF16, 0, 1, "lut<uniform+epilogue, bf16>", bf16, 8, 8, TabUniform<bf16, 10, true>),
    TDX_FAM_LUT(TDX_SRC_UNIFORM, TDX_F16, 0, 1, "lut<uniform+epilogue, f16>", f16, 8, 8, TabUniform<f16, 10, true>),
#ifdef TDX_EXPERIMENTAL_ALGOS
    // experimental variants, reachable only through an explicit TdxInitDesc.algo (bench sweeps)
    TDX_FAM(TDX_SRC_UNIFORM, TDX_BF16, 0, 7, 0, GenUniform16<bf16, 7, false>),
    TDX_FAM(TDX_SRC_NORMAL, TDX_BF16, TDX_ALGO_ICDF16, 7, 0, GenNormalICDF16<bf16, 7, false>),
    TDX_FAM(TDX_SRC_NORMAL, TDX_BF16, TDX_ALGO_BM16, 10, 0, GenNormalBM16<bf16, 10, false>),
    TDX_FAM(TDX_SRC_NORMAL, TDX_BF16, TDX_ALGO_BM16, 7, 0, GenNormalBM16<bf16, 7, false>),
    TDX_FAM(TDX_SRC_NORMAL, TDX_F32, TDX_ALGO_BM32, 7, 0, GenNormalBM32<float, 7, false>),
#endif
};
constexpr int kNumFamilies = sizeof(kFamilies) / sizeof(kFamilies[0]);

thread_local char g_err[256] = "";
thread_local int g_last_launches = 0;
thread_local size_t g_last_upload_bytes = 0;

int fail(int code, const char* msg) {
  snprintf(g_err, sizeof(g_err), "%s", msg);
  return code;
}
int cuda_fail(cudaError_t e, const char* where) {
  snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
  return static_cast<int>(e);
}

int itemsize_of(int dtype) {
  switch (dtype) {
    case TDX_F32: case TDX_RAW32: return 4;
    case TDX_BF16: case TDX_F16: case TDX_RAW16: return 2;
    case TDX_RAW8: return 1;
    case TDX_RAW64: case TDX_I64: return 8;
    default: return 0;
  }
}

int resolve_algo(const TdxInitDesc& d) {
  const int a = d.algo & 0x0f;
  if (d.src == TDX_SRC_NORMAL) {
    if (a != TDX_ALGO_DEFAULT) return a;
    return d.dtype == TDX_F32 ? TDX_ALGO_BM32 : TDX_ALGO_ICDF16;
  }
  if (d.src == TDX_SRC_UNIFORM && d.dtype != TDX_F32 && a == TDX_ALGO_WIDE32) return TDX_ALGO_WIDE32;
  return 0;
}

// Descriptors at least this large (two tiles) use the table kernel when their launch does
// (TDX_LUT_MIN_ELEMS overrides; 0 disables).  The table is paid per CTA and per distinct parameter
// set, not per descriptor -- the host sorts a family's descriptors by parameters -- so the bound only
// keeps descriptors out whose tiles would mostly take the ragged path (a sharded k_proj of
// Llama-3-8B on 8 GPUs is 2^19 elements).
uint64_t lut_min_elems() {
  static const uint64_t v = [] {
    const char* e = getenv("TDX_LUT_MIN_ELEMS");
    return e ? strtoull(e, nullptr, 10) : (1ull << 18);
  }();
  return v;
}

bool fold_fills() {
  static const bool v = [] {
    const char* e = getenv("TDX_FOLD_FILLS");
    return !(e && e[0] == '0');
  }();
  return v;
}

uint64_t lut_min_launch_elems() {
  static const uint64_t v = [] {
    const char* e = getenv("TDX_LUT_MIN_LAUNCH_ELEMS");
    return e ? strtoull(e, nullptr, 10) : kLutMinLaunchElemsDefault;
  }();
  return v;
}

int family_of(const TdxInitDesc& d) {
  if (d.src == TDX_SRC_CONST) return 0;
  if (d.src == TDX_SRC_IOTA) return 1;
  const int algo = resolve_algo(d);
  const int rounds = (d.algo & TDX_ALGO_R7) ? 7 : 10;
  const int epi = d.n_epi ? 1 : 0;
  const bool lut_kind = (d.src == TDX_SRC_NORMAL && algo == TDX_ALGO_ICDF16) ||
                        (d.src == TDX_SRC_UNIFORM && algo == 0 && d.dtype != TDX_F32);
  const bool want_lut = lut_kind && rounds == 10 && !(d.algo & TDX_ALGO_NOLUT) &&
                        lut_min_elems() != 0 && d.elem_count >= lut_min_elems();
  for (int f = 2; f < kNumFamilies; ++f) {
    const Family& F = kFamilies[f];
    if (F.src == d.src && F.dtype == d.dtype && F.algo == algo && F.rounds == rounds &&
        F.epi == epi && F.lut == want_lut)
      return f;
  }
  return -1;
}

uint64_t tiles_of(const TdxInitDesc& d, int tile_vecs) {
  if (d.elem_count == 0) return 0;
  uint64_t nvec;
  if (d.src == TDX_SRC_CONST) {
    const uint64_t a = reinterpret_cast<uintptr_t>(d.dst);
    const uint64_t end = a + d.elem_count * itemsize_of(d.dtype);
    nvec = (end - (a & ~15ull) + 15) / 16;
  } else if (d.src == TDX_SRC_IOTA) {  // vectors are counted from the descriptor's first element
    const uint64_t epv = 16 / itemsize_of(d.dtype);
    nvec = (d.elem_count + epv - 1) / epv;
  } else {
    const uint64_t epv = 16 / itemsize_of(d.dtype);
    nvec = (d.elem_begin + d.elem_count - 1) / epv - d.elem_begin / epv + 1;
  }
  return (nvec + tile_vecs - 1) / tile_vecs;
}

// Device-resident plan header.  Lives at the start of the caller's workspace.
struct PlanGroup {
  unsigned long long total_tiles;
  unsigned long long prefix_off;  // byte offsets from the workspace base
  unsigned long long desc_off;
  uint32_t n_desc;
  uint32_t family;
  unsigned long long seed;  // seed of the group's descriptors if they all share one
  uint32_t seed_shared;
  uint32_t n_chunks;             // table kernel only: entries of the work list
  unsigned long long chunk_off;  // and its byte offset
  uint32_t n_static;             // leading entries that are pre-assigned to CTAs (GroupArgs::n_static)
  uint32_t list_ctas;            // ... to this many
  unsigned long long seg_off_off;  // byte offset of the [list_ctas + 1] segment table
};
struct PlanHeader {
  uint32_t magic;
  uint32_t n_groups;
  unsigned int counters[32];
  unsigned int done[32];
  PlanGroup groups[kNumFamilies];
  TdxInitDesc one;  // the plan's descriptor if it has exactly one (GroupArgs::one)
};
constexpr uint32_t kPlanMagic = 0x58445431u;  // "TDX1"
static_assert(kNumFamilies <= 32, "counter slots");

// Work list of the table kernel: at most kMaxListChunks "large" grabs (the grab cap grows with the
// launch so that this holds), a guided tail, and one partial grab per descriptor.
constexpr size_t kMaxListChunks = 8192;
constexpr size_t kMaxListTail = 4096;
constexpr size_t kMaxListCtas = 1024;  // segment table of the pre-assigned part (one entry per CTA + 1)
unsigned long long env_ull(const char* name, unsigned long long dflt) {
  const char* e = getenv(name);
  return e && *e ? strtoull(e, nullptr, 10) : dflt;
}
// TDX_LUT_STATIC=0: every grab of the table kernel comes from the work counter (round 1's scheduler)
bool lut_static_enabled() {
  static const bool v = [] {
    const char* e = getenv("TDX_LUT_STATIC");
    return !(e && e[0] == '0');
  }();
  return v;
}
size_t lut_family_count() {
  static const size_t v = [] {
    size_t c = 0;
    for (int f = 0; f < kNumFamilies; ++f) c += kFamilies[f].lut ? 1 : 0;
    return c;
  }();
  return v;
}
size_t plan_bytes(int n) {
  // header + per-family prefix arrays (n + #families entries worst case) + descriptors + one work
  // list per table family (each at most kMaxListChunks + kMaxListTail grabs plus one per descriptor)
  return sizeof(PlanHeader) + (static_cast<size_t>(n) + kNumFamilies) * sizeof(unsigned long long) +
         static_cast<size_t>(n) * sizeof(TdxInitDesc) +
         (lut_family_count() * (kMaxListChunks + kMaxListTail) + static_cast<size_t>(n)) * sizeof(uint4) +
         lut_family_count() * kMaxListCtas * sizeof(uint32_t) +
         (16 + 128) * static_cast<size_t>(kNumFamilies) + 64;
}

struct DeviceInfo {
  int sm_count = 0;
  int blocks_per_sm[kNumFamilies] = {};
};
DeviceInfo* device_info() {
  static DeviceInfo infos[64];
  static bool ready[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  if (!ready[dev]) {
    DeviceInfo& I = infos[dev];
    if (cudaDeviceGetAttribute(&I.sm_count, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
      return nullptr;
    for (int f = 0; f < kNumFamilies; ++f) {
      int nb = 0;
      if (kFamilies[f].dyn_smem &&
          cudaFuncSetAttribute(reinterpret_cast<const void*>(kFamilies[f].fn),
                               cudaFuncAttributeMaxDynamicSharedMemorySize,
                               kFamilies[f].dyn_smem) != cudaSuccess)
        return nullptr;
      if (kFamilies[f].dyn_smem && kFamilies[f].fn_any_seed &&
          cudaFuncSetAttribute(reinterpret_cast<const void*>(kFamilies[f].fn_any_seed),
                               cudaFuncAttributeMaxDynamicSharedMemorySize,
                               kFamilies[f].dyn_smem) != cudaSuccess)
        return nullptr;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(
              &nb, reinterpret_cast<const void*>(kFamilies[f].fn), kFamilies[f].threads,
              kFamilies[f].dyn_smem) != cudaSuccess)
        return nullptr;
      I.blocks_per_sm[f] = std::max(nb, 1);
    }
    ready[dev] = true;
  }
  return &infos[dev];
}

// Builds the host image of the plan.  Returns 0 or an error code.
// `img` is scratch that only grows (zero-filling 1.6 MB per call would cost more host time than the
// planning of a small module); `used` = bytes of the image to copy to the device.
int build_plan(const TdxInitDesc* descs, int n, std::vector<unsigned char>& img, PlanHeader& hdr, size_t& used) {
  if (n < 0 || (n > 0 && descs == nullptr)) return fail(TDX_E_BADARG, "descs == NULL or n < 0");
  std::vector<int> fam(static_cast<size_t>(n));
  int per_family[kNumFamilies] = {};
  for (int i = 0; i < n; ++i) {
    const TdxInitDesc& d = descs[i];
    const int isz = itemsize_of(d.dtype);
    if (isz == 0) return fail(TDX_E_BADARG, "unknown dtype");
    if (d.src != TDX_SRC_CONST && d.dtype >= TDX_RAW8)
      return fail(TDX_E_BADARG, "raw dtypes are only valid with TDX_SRC_CONST");
    if ((d.dtype == TDX_I64 && d.src != TDX_SRC_IOTA) ||
        (d.src == TDX_SRC_IOTA && !(d.dtype == TDX_F32 || d.dtype == TDX_BF16 || d.dtype == TDX_F16 || (d.dtype == TDX_I64 && d.n_epi == 0))))
      return fail(TDX_E_BADARG, "TDX_SRC_IOTA writes TDX_F32 / TDX_BF16 / TDX_F16 (epilogue allowed) or TDX_I64 (none); TDX_I64 is IOTA-only");
    if (d.n_epi > TDX_MAX_EPI) return fail(TDX_E_BADARG, "n_epi > TDX_MAX_EPI");
    if (d.elem_count && d.dst == nullptr) return fail(TDX_E_BADARG, "dst == NULL");
    if (reinterpret_cast<uintptr_t>(d.dst) % isz) return fail(TDX_E_BADARG, "dst not element-aligned");
    fam[i] = family_of(d);
    if (fam[i] < 0) return fail(TDX_E_BADARG, "no kernel for (src, dtype, algo, epilogue)");
    per_family[fam[i]]++;
  }
  // The table kernel pays 65536 evaluations per CTA up front: worth it only if the launch as a
  // whole has enough table-eligible work.  Otherwise its descriptors go to the direct twin
  // (same bits either way).
  {
    uint64_t lut_elems[kNumFamilies] = {};
    for (int i = 0; i < n; ++i)
      if (kFamilies[fam[i]].lut) lut_elems[fam[i]] += descs[i].elem_count;
    for (int f = 0; f < kNumFamilies; ++f) {
      // measured break-even (profiles/r1_lut_launch_threshold.txt): ~64 MB for the normal, whose
      // direct kernel is the slowest, ~128 MB for the uniform
      const uint64_t need = lut_min_launch_elems() * (kFamilies[f].src == TDX_SRC_UNIFORM ? 2 : 1);
      if (!kFamilies[f].lut || lut_elems[f] == 0 || lut_elems[f] >= need) continue;
      int twin = -1;
      for (int t = 1; t < kNumFamilies; ++t) {
        const Family &A = kFamilies[f], &B = kFamilies[t];
        if (!B.lut && B.src == A.src && B.dtype == A.dtype && B.algo == A.algo && B.rounds == A.rounds &&
            B.epi == A.epi)
          twin = t;
      }
      if (twin < 0) continue;
      for (int i = 0; i < n; ++i)
        if (fam[i] == f) {
          fam[i] = twin;
          per_family[f]--;
          per_family[twin]++;
        }
    }
  }
  // A module's constant fills (norm weights, biases: KBs) ride in the largest table launch's work
  // list instead of costing a launch of their own -- unless they are a real share of the bytes.
  static_assert(TDX_SRC_CONST == 0, "family 0 = fills");
  const int iota_family = [] {
    for (int f = 0; f < kNumFamilies; ++f)
      if (kFamilies[f].src == TDX_SRC_IOTA) return f;
    return -1;
  }();
  if ((per_family[0] > 0 || (iota_family >= 0 && per_family[iota_family] > 0)) && fold_fills()) {
    int host = -1;
    uint64_t host_bytes = 0, fill_bytes = 0;
    uint64_t bytes_of[kNumFamilies] = {};
    for (int i = 0; i < n; ++i) bytes_of[fam[i]] += descs[i].elem_count * itemsize_of(descs[i].dtype);
    fill_bytes = bytes_of[0] + (iota_family >= 0 ? bytes_of[iota_family] : 0);
    for (int f = 1; f < kNumFamilies; ++f)
      if (kFamilies[f].lut && per_family[f] > 0 && bytes_of[f] > host_bytes) {
        host = f;
        host_bytes = bytes_of[f];
      }
    if (host >= 0 && fill_bytes * 16 <= host_bytes) {
      for (int i = 0; i < n; ++i)
        if (fam[i] == 0 || fam[i] == iota_family) {
          per_family[fam[i]]--;
          fam[i] = host;
          per_family[host]++;
        }
    }
  }
  memset(&hdr, 0, sizeof(hdr));
  hdr.magic = kPlanMagic;
  if (img.size() < plan_bytes(n)) img.resize(plan_bytes(n));
  int sm_count = 148;
  if (DeviceInfo* info = device_info()) sm_count = info->sm_count;
  std::vector<int> order;
  thread_local std::vector<uint32_t> seg_fill;
  order.reserve(static_cast<size_t>(n));
  size_t off = (sizeof(PlanHeader) + 15) & ~static_cast<size_t>(15);
  for (int f = 0; f < kNumFamilies; ++f) {
    if (!per_family[f]) continue;
    PlanGroup& G = hdr.groups[hdr.n_groups++];
    G.family = f;
    G.n_desc = per_family[f];
    G.prefix_off = off;
    auto* prefix = reinterpret_cast<unsigned long long*>(img.data() + off);
    off += (static_cast<size_t>(G.n_desc) + 1) * sizeof(unsigned long long);
    off = (off + 127) & ~static_cast<size_t>(127);  // (a descriptor = one cache line: lut_prefetch_desc)
    G.desc_off = off;
    auto* out = reinterpret_cast<TdxInitDesc*>(img.data() + off);
    off += static_cast<size_t>(G.n_desc) * sizeof(TdxInitDesc);
    unsigned long long acc = 0;
    uint32_t k = 0;
    G.seed_shared = 1;
    // Table kernel: a CTA rebuilds its 65536-entry table (~10 us) whenever the parameters change
    // from one grab to the next, so descriptors with equal parameters are made neighbours (GPT-2's
    // alternating std 0.02 / 0.02/sqrt(2L) would otherwise cost a rebuild per tensor).  The order
    // inside a family is free: every descriptor is self-contained.
    order.clear();
    for (int i = 0; i < n; ++i)
      if (fam[i] == f) order.push_back(i);
    if (kFamilies[f].lut)
      std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
        const TdxInitDesc &x = descs[a], &y = descs[b];
        const bool xr = x.src != TDX_SRC_CONST && x.src != TDX_SRC_IOTA, yr = y.src != TDX_SRC_CONST && y.src != TDX_SRC_IOTA;
        if (xr != yr) return yr;  // (folded fills and index programs first)
        if (!xr) return false;
        if (x.p0 != y.p0) return x.p0 < y.p0;
        if (x.p1 != y.p1) return x.p1 < y.p1;
        if (x.n_epi != y.n_epi) return x.n_epi < y.n_epi;
        return x.n_epi != 0 && memcmp(x.epi, y.epi, sizeof(TdxEpiStep) * x.n_epi) < 0;
      });
    bool have_seed = false;
    for (int i : order) {
      if (descs[i].src != TDX_SRC_CONST && descs[i].src != TDX_SRC_IOTA) {
        if (!have_seed) G.seed = descs[i].philox_seed;
        else if (descs[i].philox_seed != G.seed) G.seed_shared = 0;
        have_seed = true;
      }
      prefix[k] = acc;
      out[k] = descs[i];
      if (n == 1) hdr.one = descs[i];
      acc += tiles_of(descs[i], kFamilies[f].tile_vecs);
      ++k;
    }
    prefix[k] = acc;
    G.total_tiles = acc;
    if (kFamilies[f].lut) {
      // Work list of the table kernel (for_each_listed_chunk).  Head: everything but four tiles per
      // CTA of tail (measured, profiles/r2_lut_tail_sweep.txt: two leave the CTAs 27 us apart at the
      // end of a 16 GB launch, eight cost more grabs than they buy balance) is cut into one
      // contiguous, equal share per CTA -- a share is a few grabs, broken only where a descriptor
      // ends -- and CTA b owns entries [seg[b], seg[b + 1]).  Tail: guided sizes handed out by the
      // work counter (grab = remaining / (2 * CTAs), between one tile and the cap), which is what
      // lets the CTAs finish together whatever their speeds were.  Fewer grabs matter: a grab ends
      // in a barrier that re-aligns the 32 warps of a CTA, and the lock-step phase that follows
      // costs ~2-3 us of throughput (benchmarks/lut_timeline.py, profiles/r2_lut_timeline.jsonl).
      off = (off + 15) & ~static_cast<size_t>(15);
      G.chunk_off = off;
      auto* list = reinterpret_cast<uint4*>(img.data() + off);
      const unsigned long long ctas = static_cast<unsigned long long>(std::max(sm_count, 1));
      const unsigned long long cap =
          std::max<unsigned long long>(kLutMaxTilesPerChunk, (acc + kMaxListChunks - 1) / kMaxListChunks);
      unsigned long long static_tiles = 0;
      static const unsigned long long tail_min = env_ull("TDX_LUT_TAIL_MIN", 4), tail_max = env_ull("TDX_LUT_TAIL_MAX", 4),
                                      tail_div = std::max(env_ull("TDX_LUT_TAIL_DIV", 2), 1ull);
      if (lut_static_enabled() && ctas < kMaxListCtas && acc >= ctas * 16ull)
        static_tiles = acc - std::min(std::max(acc / 8ull, ctas * tail_min), ctas * tail_max);
      seg_fill.assign(static_cast<size_t>(ctas) + 1, 0u);  // seg_fill[b] = first entry of CTA b
      unsigned long long remaining = acc;
      uint32_t nc = 0, cta = 0;
      unsigned long long room = static_tiles / ctas + (0 < static_tiles % ctas ? 1 : 0);  // CTA 0's share
      for (uint32_t di = 0; di < G.n_desc; ++di) {
        unsigned long long t = 0;
        const unsigned long long nt = prefix[di + 1] - prefix[di];
        while (t < nt) {
          unsigned long long sz;
          if (cta < ctas && static_tiles) {  // pre-assigned part
            sz = std::min<unsigned long long>(std::min(nt - t, room), 0xffffffffull);
          } else {
            sz = std::min(std::min(std::max(remaining / ((static_tiles ? tail_div : 2ull) * ctas), 1ull), cap), nt - t);
          }
          if (off + (static_cast<size_t>(nc) + 1) * sizeof(uint4) > img.size())
            return fail(TDX_E_WORKSPACE, "internal: work list exceeds its bound");
          list[nc++] = make_uint4(di, static_cast<uint32_t>(sz), static_cast<uint32_t>(t),
                                  static_cast<uint32_t>(t >> 32));
          t += sz;
          remaining -= sz;
          if (cta < ctas && static_tiles) {
            room -= sz;
            if (room == 0) {
              ++cta;
              seg_fill[cta] = nc;
              room = static_tiles / ctas + (cta < static_tiles % ctas ? 1 : 0);
            }
          }
        }
      }
      G.n_chunks = nc;
      off += static_cast<size_t>(nc) * sizeof(uint4);
      G.n_static = 0;
      G.list_ctas = static_cast<uint32_t>(ctas);
      G.seg_off_off = off;
      if (static_tiles && cta == ctas) {
        if (off + (static_cast<size_t>(ctas) + 1) * sizeof(uint32_t) > img.size())
          return fail(TDX_E_WORKSPACE, "internal: segment table exceeds its bound");
        memcpy(img.data() + off, seg_fill.data(), (static_cast<size_t>(ctas) + 1) * sizeof(uint32_t));
        G.n_static = seg_fill[ctas];
        off += (static_cast<size_t>(ctas) + 1) * sizeof(uint32_t);
      }
    }
  }
  memcpy(img.data(), &hdr, sizeof(hdr));
  used = off;
  return 0;
}

// ---- pinned staging ring for plan uploads ------------------------------------------------------
struct StageSlot {
  void* host = nullptr;
  size_t cap = 0;
  cudaEvent_t done = nullptr;
  int device = -1;
  bool pending = false;
};
constexpr int kStageSlots = 8;
thread_local StageSlot g_stage[kStageSlots];
thread_local int g_stage_pos = 0;

int stage(const void* src, size_t bytes, cudaStream_t stream, void* dst_device, void** pinned_out) {
  StageSlot& s = g_stage[g_stage_pos];
  g_stage_pos = (g_stage_pos + 1) % kStageSlots;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
  if (s.pending) {  // the copy that last used this slot must have left it
    e = cudaEventSynchronize(s.done);
    if (e != cudaSuccess) return cuda_fail(e, "cudaEventSynchronize(staging)");
    s.pending = false;
  }
  if (s.done == nullptr || s.device != dev) {
    if (s.done) cudaEventDestroy(s.done);
    e = cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming);
    if (e != cudaSuccess) return cuda_fail(e, "cudaEventCreate(staging)");
    s.device = dev;
  }
  if (s.cap < bytes) {
    if (s.host) cudaFreeHost(s.host);
    s.cap = std::max<size_t>(bytes * 2, 64 << 10);
    e = cudaHostAlloc(&s.host, s.cap, cudaHostAllocPortable);
    if (e != cudaSuccess) { s.host = nullptr; s.cap = 0; return cuda_fail(e, "cudaHostAlloc(staging)"); }
  }
  memcpy(s.host, src, bytes);
  e = cudaMemcpyAsync(dst_device, s.host, bytes, cudaMemcpyHostToDevice, stream);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpyAsync(plan)");
  e = cudaEventRecord(s.done, stream);
  if (e != cudaSuccess) return cuda_fail(e, "cudaEventRecord(staging)");
  s.pending = true;
  *pinned_out = s.host;
  return 0;
}

int launch_groups(const PlanHeader& hdr, void* workspace, cudaStream_t stream) {
  DeviceInfo* info = device_info();
  if (!info) return fail(TDX_E_NODEVICE, "no CUDA device / occupancy query failed");
  auto* base = static_cast<unsigned char*>(workspace);
  auto* dev_hdr = reinterpret_cast<PlanHeader*>(base);
  int launches = 0;
  for (uint32_t gi = 0; gi < hdr.n_groups; ++gi) {
    const PlanGroup& G = hdr.groups[gi];
    if (G.total_tiles == 0) continue;
    GroupArgs a;
    a.tile_prefix = reinterpret_cast<const unsigned long long*>(base + G.prefix_off);
    a.descs = reinterpret_cast<const TdxInitDesc*>(base + G.desc_off);
    a.total_tiles = G.total_tiles;
    a.counter = &dev_hdr->counters[gi];  // zero: uploaded so, and put back by the kernel (leave_grid)
    a.done = &dev_hdr->done[gi];
    a.n_desc = G.n_desc;
    a.inline_desc = (hdr.n_groups == 1 && G.n_desc == 1) ? 1u : 0u;  // (hdr.one is set for one-descriptor plans)
    a.one = hdr.one;
    a.seed_shared = G.seed_shared;
    {
      uint32_t k0 = static_cast<uint32_t>(G.seed), k1 = static_cast<uint32_t>(G.seed >> 32);
      for (int r = 0; r < 10; ++r) {
        a.rk[2 * r] = k0;
        a.rk[2 * r + 1] = k1;
        k0 += kPhiloxW0;
        k1 += kPhiloxW1;
      }
    }
    const unsigned long long resident =
        static_cast<unsigned long long>(info->sm_count) * info->blocks_per_sm[G.family];
    // 256-thread kernels: full-size grabs only when every resident CTA still gets a few of them
    int tpc = kFamilies[G.family].tiles_per_chunk;
    a.static_grabs = 0;
    if (!kFamilies[G.family].lut) {
      if (G.total_tiles <= resident * kStaticTilesPerCta) {
        // small launch: one grab of consecutive tiles per CTA, assigned by block index
        tpc = static_cast<int>((G.total_tiles + resident - 1) / resident);
        a.static_grabs = 1;
      } else {
        const unsigned long long want = G.total_tiles / (resident * 2ull);
        tpc = static_cast<int>(std::min<unsigned long long>(std::max<unsigned long long>(want, 1ull), tpc));
      }
    }
    a.tiles_per_chunk = static_cast<uint32_t>(tpc);
    a.n_chunks = G.n_chunks;
    a.chunks = reinterpret_cast<const uint4*>(base + G.chunk_off);
    a.n_static = G.n_static;
    a.seg_off = reinterpret_cast<const uint32_t*>(base + G.seg_off_off);
    const unsigned long long chunks =
        kFamilies[G.family].lut ? G.n_chunks : (G.total_tiles + tpc - 1) / tpc;
    const unsigned int grid = static_cast<unsigned int>(std::min(chunks, resident));
    if (G.n_static && grid != G.list_ctas) return fail(TDX_E_BADARG, "internal: work list was cut for another grid");
    const Family& F = kFamilies[G.family];
    const KernelFn fn = (F.fn_any_seed && !G.seed_shared) ? F.fn_any_seed : F.fn;
    fn<<<grid, F.threads, F.dyn_smem, stream>>>(a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, kFamilies[G.family].name);
    ++launches;
  }
  g_last_launches = launches;
  return 0;
}

}  // namespace tdx

extern "C" {

TDX_C_API size_t tdx_init_workspace_bytes(int n) { return tdx::plan_bytes(n < 0 ? 0 : n); }

TDX_C_API int tdx_plan_upload(const TdxInitDesc* descs, int n, void* workspace,
                              size_t workspace_bytes, void* stream, TdxPlan* plan) {
  static_assert(sizeof(tdx::PlanHeader) <= sizeof(TdxPlan), "TdxPlan too small");
  thread_local std::vector<unsigned char> img;
  tdx::PlanHeader hdr;
  if (plan == nullptr) return tdx::fail(TDX_E_BADARG, "plan == NULL");
  size_t used = 0;
  if (int rc = tdx::build_plan(descs, n, img, hdr, used)) return rc;
  memcpy(plan, &hdr, sizeof(hdr));
  if (workspace == nullptr || workspace_bytes < used)
    return tdx::fail(TDX_E_WORKSPACE, "workspace too small (see tdx_init_workspace_bytes)");
  tdx::g_last_upload_bytes = used;
  cudaError_t e = cudaMemcpyAsync(workspace, img.data(), used, cudaMemcpyHostToDevice,
                                  static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return tdx::cuda_fail(e, "cudaMemcpyAsync(plan)");
  // the staging image is reused by the next call on this thread: make sure the copy has left it
  e = cudaStreamSynchronize(static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return tdx::cuda_fail(e, "cudaStreamSynchronize(plan upload)");
  return 0;
}

TDX_C_API int tdx_plan_launch(const TdxPlan* plan, void* workspace, void* stream) {
  if (workspace == nullptr) return tdx::fail(TDX_E_WORKSPACE, "workspace == NULL");
  if (plan == nullptr) return tdx::fail(TDX_E_BADARG, "plan == NULL");
  tdx::PlanHeader hdr;
  memcpy(&hdr, plan, sizeof(hdr));
  if (hdr.magic != tdx::kPlanMagic) return tdx::fail(TDX_E_BADARG, "plan was not produced by tdx_plan_upload");
  return tdx::launch_groups(hdr, workspace, static_cast<cudaStream_t>(stream));
}

// The host image of the plan tdx_init_prepare() built on this thread, waiting for tdx_init_submit().
namespace tdx {
thread_local std::vector<unsigned char> g_prepared_img;
thread_local PlanHeader g_prepared_hdr;
thread_local size_t g_prepared_used = 0;
thread_local bool g_prepared = false;
}  // namespace tdx

TDX_C_API int tdx_init_prepare(const TdxInitDesc* descs, int n, size_t* workspace_bytes) {
  tdx::g_prepared = false;
  if (workspace_bytes == nullptr) return tdx::fail(TDX_E_BADARG, "workspace_bytes == NULL");
  if (int rc = tdx::build_plan(descs, n, tdx::g_prepared_img, tdx::g_prepared_hdr, tdx::g_prepared_used)) return rc;
  *workspace_bytes = tdx::g_prepared_hdr.n_groups ? tdx::g_prepared_used : 0;
  tdx::g_prepared = true;
  return 0;
}

TDX_C_API int tdx_init_submit(void* workspace, size_t workspace_bytes, void* stream) {
  if (!tdx::g_prepared) return tdx::fail(TDX_E_BADARG, "tdx_init_submit without tdx_init_prepare on this thread");
  tdx::g_prepared = false;
  const tdx::PlanHeader& hdr = tdx::g_prepared_hdr;
  if (hdr.n_groups == 0) {
    tdx::g_last_launches = 0;
    return 0;
  }
  const size_t used = tdx::g_prepared_used;
  if (workspace == nullptr || workspace_bytes < used)
    return tdx::fail(TDX_E_WORKSPACE, "workspace too small (see tdx_init_prepare)");
  // The plan image goes through a small ring of pinned staging buffers so that the copy is truly
  // asynchronous: the host can go on planning the next batch while the GPU works on this one.
  tdx::g_last_upload_bytes = used;
  void* pinned = nullptr;
  if (int rc = tdx::stage(tdx::g_prepared_img.data(), used, static_cast<cudaStream_t>(stream), workspace, &pinned))
    return rc;
  return tdx::launch_groups(hdr, workspace, static_cast<cudaStream_t>(stream));
}

TDX_C_API int tdx_init_launch(const TdxInitDesc* descs, int n, void* workspace,
                              size_t workspace_bytes, void* stream) {
  size_t need = 0;
  if (int rc = tdx_init_prepare(descs, n, &need)) return rc;
  if (need == 0) {
    tdx::g_prepared = false;
    tdx::g_last_launches = 0;
    return 0;
  }
  if (workspace == nullptr || workspace_bytes < need) {
    tdx::g_prepared = false;
    return tdx::fail(TDX_E_WORKSPACE, "workspace too small (see tdx_init_workspace_bytes)");
  }
  return tdx_init_submit(workspace, workspace_bytes, stream);
}

TDX_C_API int tdx_last_launch_count(void) { return tdx::g_last_launches; }
TDX_C_API size_t tdx_last_upload_bytes(void) { return tdx::g_last_upload_bytes; }

TDX_C_API int tdx_elems_per_block(int dtype, int src, int algo) {
  if (src == TDX_SRC_CONST) return 0;
  if (dtype == TDX_F32) return 4;
  if (dtype != TDX_BF16 && dtype != TDX_F16) return 0;
  if ((algo & 0x0f) == TDX_ALGO_WIDE32) return 4;  // BM32 == WIDE32
  return 8;
}

#ifdef TDX_LUT_TIMELINE
TDX_C_API int tdx_debug_lut_timeline(unsigned long long* dev_buf) {
  return cudaMemcpyToSymbol(tdx::g_lut_timeline, &dev_buf, sizeof(dev_buf)) == cudaSuccess ? 0 : -1;
}
#endif
TDX_C_API int tdx_abi_version(void) { return TDX_ABI_VERSION; }
TDX_C_API const char* tdx_last_error(void) { return tdx::g_err; }

}  // extern "C"
