/*
 * tdx_oracle.c -- CPU restatement of the fused initialisation kernels (TEST INFRASTRUCTURE).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this file; the
 * product (torchdistx_b200) never does and has no CPU path for CUDA tensors.
 *
 * What is restated, and against what it is pinned:
 *   philox4x32_R     the Philox4x32 generator ATen/curand use on CUDA.  Published algorithm:
 *                    Salmon, Moraes, Dror, Shaw, SC'11 (Random123 1.x).  Dependency of the
 *                    reference path: PyTorch (un-vendored; pinned here: torch 2.11.0+cu128,
 *                    $TORCH/include/ATen/core/PhiloxRNGEngine.h:60-220).  Pinned by
 *                    tests/golden/philox_kat.json = Random123 known-answer vectors + outputs of
 *                    at::Philox4_32 itself (tests/golden/make_philox_kat.py).
 *   transforms       element g of a tensor := f(seed, offset, g) exactly as
 *                    torchdistx_b200/csrc/kernels/tdx_init_kernels.cu computes it, with libm in
 *                    place of the MUFU approximations (lg2/sqrt/sin/cos), so RNG floats agree to
 *                    the tolerance stated in tests/test_kernels_gpu.py and everything else
 *                    (constants, uniform, rounding, epilogue arithmetic, sharding) bit for bit.
 *   semantics        what uniform_/normal_/fill_ MEAN follows ATen, the arithmetic owner of the
 *                    reference's replay (reference src/cc/torchdistx/deferred_init.cc:218-220):
 *                    bounds cast to the tensor dtype, value = u*(to-from)+from, half-open range
 *                    ($TORCH/include/ATen/native/cuda/DistributionTemplates.h:428-470,
 *                     $TORCH/include/ATen/core/TransformationHelper.h:84-99).
 *                    The element-wise random STREAM is this engine's own (shard-invariant Philox
 *                    indexing); agreement with the reference's CPU mt19937 stream is
 *                    distributional and checked against oracle/_ref in tests/.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "tdx_init.h"

#define PHILOX_M0 0xD2511F53u
#define PHILOX_M1 0xCD9E8D57u
#define PHILOX_W0 0x9E3779B9u
#define PHILOX_W1 0xBB67AE85u

/* out[4] = Philox4x32-`rounds`(counter ctr[4], key[2]) */
void tdx_oracle_philox4x32(const uint32_t ctr[4], const uint32_t key[2], int rounds, uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
  for (int r = 0; r < rounds; ++r) {
    const uint64_t p0 = (uint64_t)c0 * PHILOX_M0, p1 = (uint64_t)c2 * PHILOX_M1;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += PHILOX_W0; k1 += PHILOX_W1;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* the engine's counter layout (include/tdx_init.h "Random stream specification") */
static void block_of(const TdxInitDesc* d, uint64_t blk, uint32_t wflag, int rounds, uint32_t out[4]) {
  const uint32_t ctr[4] = {(uint32_t)d->philox_offset, (uint32_t)(blk >> 32),
                           (uint32_t)(d->philox_offset >> 32) | 0x80000000u | wflag, (uint32_t)blk};
  const uint32_t key[2] = {(uint32_t)d->philox_seed, (uint32_t)(d->philox_seed >> 32)};
  tdx_oracle_philox4x32(ctr, key, rounds, out);
}

/* ---- dtype helpers ------------------------------------------------------------------------- */
static float bits_f(uint32_t b) { float f; memcpy(&f, &b, 4); return f; }
static uint32_t f_bits(float f) { uint32_t b; memcpy(&b, &f, 4); return b; }

static uint16_t f32_to_bf16(float f) { /* round to nearest even */
  uint32_t b = f_bits(f);
  if ((b & 0x7fffffffu) > 0x7f800000u) return 0x7fff;
  b += 0x7fffu + ((b >> 16) & 1u);
  return (uint16_t)(b >> 16);
}
static float bf16_to_f32(uint16_t h) { return bits_f((uint32_t)h << 16); }
static uint16_t f32_to_f16(float f) { _Float16 h = (_Float16)f; uint16_t b; memcpy(&b, &h, 2); return b; }
static float f16_to_f32(uint16_t b) { _Float16 h; memcpy(&h, &b, 2); return (float)h; }

static float round_through(float v, int dtype) {
  if (dtype == TDX_BF16) return bf16_to_f32(f32_to_bf16(v));
  if (dtype == TDX_F16) return f16_to_f32(f32_to_f16(v));
  return v;
}
/* largest representable value of `dtype` strictly below `to` (kernel: OutTraits<>::prev) */
static float prev_of(float to, int dtype) {
  if (dtype == TDX_F32) {
    if (to > 0.f) return bits_f(f_bits(to) - 1);
    if (to < 0.f) return bits_f(f_bits(to) + 1);
    return bits_f(0x80000001u);
  }
  if (dtype == TDX_BF16) {
    if (to > 0.f) return bits_f(f_bits(to) - 0x10000u);
    if (to < 0.f) return bits_f(f_bits(to) + 0x10000u);
    return bits_f(0x80010000u);
  }
  {
    const uint16_t h = f32_to_f16(to);
    if (to > 0.f) return f16_to_f32((uint16_t)(h - 1));
    if (to < 0.f) return f16_to_f32((uint16_t)(h + 1));
    return f16_to_f32(0x8001);
  }
}

/* ---- inverse normal CDF in double (Wichura, AS 241, PPND16): used for the rare tail bin and for
 * erfinv in the epilogue; relative accuracy ~1e-16, far below the fp32 the kernel works in. ------- */
static double ndtri(double p) {
  static const double a[8] = {3.3871328727963666080e0, 1.3314166789178437745e+2, 1.9715909503065514427e+3,
                              1.3731693765509461125e+4, 4.5921953931549871457e+4, 6.7265770927008700853e+4,
                              3.3430575583588128105e+4, 2.5090809287301226727e+3};
  static const double b[8] = {1.0, 4.2313330701600911252e+1, 6.8718700749205790830e+2,
                              5.3941960214247511077e+3, 2.1213794301586595867e+4, 3.9307895800092710610e+4,
                              2.8729085735721942674e+4, 5.2264952788528545610e+3};
  static const double c[8] = {1.42343711074968357734e0, 4.63033784615654529590e0, 5.76949722146069140550e0,
                              3.64784832476320460504e0, 1.27045825245236838258e0, 2.41780725177450611770e-1,
                              2.27238449892691845833e-2, 7.74545014278341407640e-4};
  static const double d[8] = {1.0, 2.05319162663775882187e0, 1.67638483018380384940e0,
                              6.89767334985100004550e-1, 1.48103976427480074590e-1, 1.51986665636164571966e-2,
                              5.47593808499534494600e-4, 1.05075007164441684324e-9};
  static const double e[8] = {6.65790464350110377720e0, 5.46378491116411436990e0, 1.78482653991729133580e0,
                              2.96560571828504891230e-1, 2.65321895265761230930e-2, 1.24266094738807843860e-3,
                              2.71155556874348757815e-5, 2.01033439929228813265e-7};
  static const double f[8] = {1.0, 5.99832206555887937690e-1, 1.36929880922735805310e-1,
                              1.48753612908506148525e-2, 7.86869131145613259100e-4, 1.84631831751005468180e-5,
                              1.42151175831644588870e-7, 2.04426310338993978564e-15};
  const double q = p - 0.5;
  double r, num, den;
  int i;
  if (fabs(q) <= 0.425) {
    r = 0.180625 - q * q;
    num = a[7]; den = b[7];
    for (i = 6; i >= 0; --i) { num = num * r + a[i]; den = den * r + b[i]; }
    return q * num / den;
  }
  r = q < 0 ? p : 1.0 - p;
  r = sqrt(-log(r));
  if (r <= 5.0) {
    r -= 1.6;
    num = c[7]; den = d[7];
    for (i = 6; i >= 0; --i) { num = num * r + c[i]; den = den * r + d[i]; }
  } else {
    r -= 5.0;
    num = e[7]; den = f[7];
    for (i = 6; i >= 0; --i) { num = num * r + e[i]; den = den * r + f[i]; }
  }
  return q < 0 ? -num / den : num / den;
}

static float apply_epi(const TdxInitDesc* d, float v) {
  if (!(d->reserved & TDX_FLAG_SRC_NOROUND)) v = round_through(v, d->dtype);
  for (int i = 0; i < d->n_epi && i < TDX_MAX_EPI; ++i) {
    const uint32_t op = d->epi[i].op;
    switch (op & 0xffu) {
      case TDX_EPI_MUL: v = v * d->epi[i].a; break;
      case TDX_EPI_ADD: v = v + d->epi[i].a; break;
      case TDX_EPI_ERFINV: v = (float)(ndtri(((double)v + 1.0) * 0.5) * 0.70710678118654752440); break;
      case TDX_EPI_CLAMP: v = fminf(fmaxf(v, d->epi[i].a), d->epi[i].b); break;
      case TDX_EPI_RPOW: v = powf(d->epi[i].a, v); break; /* libm vs CUDA powf: both ~1 ulp, not each other's bits */
      case TDX_EPI_RECIP: v = 1.0f / v; break;
      default: break;
    }
    if (!(op & TDX_EPI_NOROUND)) v = round_through(v, d->dtype);
  }
  return v;
}

static uint32_t halfword(const uint32_t w[4], int e) { return (w[e >> 1] >> (16 * (e & 1))) & 0xffffu; }

/* kernel: icdf16_tail */
static float icdf16_tail(uint32_t word) {
  const float sgn = (word & 0x80000000u) ? 1.0f : -1.0f;
  const float p = ((float)(word & 0x7fffffffu) + 0.5f) * 4.656612873077393e-10f * 7.62939453125e-06f;
  return sgn * -(float)ndtri((double)p);
}

/* kernel: box_muller32 */
static void box_muller32(uint32_t x, uint32_t y, float* r, float* c, float* s) {
  const float u1 = fmaf((float)x, 2.3283064365386963e-10f, 1.1641532182693481e-10f);
  *r = sqrtf(-1.3862943611198906f * log2f(u1));
  const float ang = (float)(int32_t)y * 1.4629180792671596e-09f; /* y signed, times pi / 2^31 */
  *c = cosf(ang);
  *s = sinf(ang);
}

static int resolve_algo(const TdxInitDesc* d) {
  const int a = d->algo & 0x0f;
  if (d->src == TDX_SRC_NORMAL) return a ? a : (d->dtype == TDX_F32 ? TDX_ALGO_BM32 : TDX_ALGO_ICDF16);
  if (d->src == TDX_SRC_UNIFORM && d->dtype != TDX_F32 && a == TDX_ALGO_WIDE32) return TDX_ALGO_WIDE32;
  return 0;
}

/* Value (fp32, before the rounding of the store) of GLOBAL element g of the descriptor's tensor. */
static float element(const TdxInitDesc* d, uint64_t g) {
  const int rounds = (d->algo & TDX_ALGO_R7) ? 7 : 10;
  const int algo = resolve_algo(d);
  uint32_t w[4];
  float v = 0.f;
  if (d->src == TDX_SRC_UNIFORM) {
    const float from = (float)d->p0, to = (float)d->p1;
    /* the wide form of a 16-bit output keeps fp32 bounds (kernel: GenUniform32<Out>) */
    const int wide = d->dtype == TDX_F32 || algo == TDX_ALGO_WIDE32;
    const float to_prev = to > from ? prev_of(to, wide ? TDX_F32 : d->dtype) : to;
    if (wide) {
      block_of(d, g / 4, 0, rounds, w);
      const float k = (float)w[g % 4]; /* all 32 bits, round to nearest (kernel: I2FP) */
      v = fminf(fmaf(k, (to - from) * 2.3283064365386963e-10f, from), to_prev);
    } else {
      block_of(d, g / 8, 0, rounds, w);
      const float k = (float)halfword(w, (int)(g % 8));
      v = fminf(fmaf(k, (to - from) * 1.52587890625e-05f, from), to_prev);
    }
  } else { /* TDX_SRC_NORMAL */
    const float mean = (float)d->p0, std = (float)d->p1;
    if (algo == TDX_ALGO_BM32) {
      float r[2], dir[4];
      block_of(d, g / 4, 0, rounds, w);
      box_muller32(w[0], w[1], &r[0], &dir[0], &dir[1]);
      box_muller32(w[2], w[3], &r[1], &dir[2], &dir[3]);
      v = fmaf(r[(g % 4) >> 1] * std, dir[g % 4], mean);
    } else { /* TDX_ALGO_ICDF16 */
      const int e = (int)(g % 8);
      block_of(d, g / 8, 0, rounds, w);
      const uint32_t k = halfword(w, e);
      if (k == 0) {
        uint32_t t[4];
        block_of(d, g / 8, e < 4 ? 0x40000000u : 0x20000000u, rounds, t);
        v = fmaf(icdf16_tail(t[e & 3]), std, mean);
      } else {
        const float x = fmaf(8388608.0f + (float)k, 3.0517578125e-05f, -257.0f);
        const float t1 = fmaf(-x, x, 1.0f);
        const float l = log2f(t1);
        float q = fmaf(std * 0x1.152c90p-19f, l, std * 0x1.9c35c0p-14f);
        q = fmaf(q, l, std * 0x1.bfaecap-10f);
        q = fmaf(q, l, std * 0x1.26ef76p-7f);
        q = fmaf(q, l, std * -0x1.d03266p-3f);
        q = fmaf(q, l, std * 0x1.40de66p+0f);
        v = fmaf(q, x, mean);
      }
    }
  }
  return d->n_epi ? apply_epi(d, v) : v;
}

/*
 * Fills `out` (host memory, elem_count elements of the descriptor's dtype) with what the GPU
 * kernels write for this descriptor.  Returns 0, or -1 for an unsupported descriptor.
 */
int tdx_oracle_generate(const TdxInitDesc* d, void* out) {
  if (d->src == TDX_SRC_CONST) {
    const int isz = (d->dtype == TDX_F32 || d->dtype == TDX_RAW32) ? 4 : d->dtype == TDX_RAW64 ? 8
                    : d->dtype == TDX_RAW8 ? 1 : 2;
    unsigned char pat[16];
    memcpy(pat, d->fill_bits, 16);
    for (uint64_t i = 0; i < d->elem_count; ++i) memcpy((unsigned char*)out + i * isz, pat, isz);
    return 0;
  }
  if (d->src == TDX_SRC_IOTA) { /* element g = p0 + g * p1 (integers), then the fp32 epilogue */
    const int64_t start = (int64_t)d->p0, step = (int64_t)d->p1;
    if (!(d->dtype == TDX_F32 || d->dtype == TDX_BF16 || d->dtype == TDX_F16 || (d->dtype == TDX_I64 && d->n_epi == 0)))
      return -1;
    for (uint64_t i = 0; i < d->elem_count; ++i) {
      const int64_t val = start + (int64_t)(d->elem_begin + i) * step;
      if (d->dtype == TDX_I64) { ((int64_t*)out)[i] = val; continue; }
      float v = (float)val;
      if (d->n_epi) { /* index programs run in fp32 whatever the output type (kernel: apply_epi<float>) */
        TdxInitDesc f = *d;
        f.dtype = TDX_F32;
        v = apply_epi(&f, v);
      }
      if (d->dtype == TDX_F32) ((float*)out)[i] = v;
      else if (d->dtype == TDX_BF16) ((uint16_t*)out)[i] = f32_to_bf16(v); /* one rounding, at the store */
      else ((uint16_t*)out)[i] = f32_to_f16(v);
    }
    return 0;
  }
  if (d->dtype != TDX_F32 && d->dtype != TDX_BF16 && d->dtype != TDX_F16) return -1;
  {
    const int algo = resolve_algo(d);
    if (d->src == TDX_SRC_NORMAL && algo != TDX_ALGO_BM32 && !(algo == TDX_ALGO_ICDF16 && d->dtype != TDX_F32))
      return -1; /* experimental sweep variants are not part of the specification */
    if (d->algo & TDX_ALGO_R7) return -1;
  }
  for (uint64_t i = 0; i < d->elem_count; ++i) {
    const float v = element(d, d->elem_begin + i);
    if (d->dtype == TDX_F32) ((float*)out)[i] = v;
    else if (d->dtype == TDX_BF16) ((uint16_t*)out)[i] = f32_to_bf16(v);
    else ((uint16_t*)out)[i] = f32_to_f16(v);
  }
  return 0;
}

/* Philox blocks one RNG op consumes from the generator (planner.cc assign_rng). */
uint64_t tdx_oracle_offset_increment(uint64_t numel) {
  const uint64_t blocks = (numel + 3) / 4;
  return ((blocks + 3) / 4) * 4 + 4;
}
