/*
 * tdx_init.h -- C ABI of the B200 (sm_100a) fused-initialisation kernel library
 * (libtdx_init.so).
 *
 * This is the drop-in boundary BELOW the Python/pybind surface of
 * torchdistx.deferred_init.  The reference replays every recorded aten op
 * through the PyTorch dispatcher, one ATen kernel per op
 * (reference: src/cc/torchdistx/deferred_init.cc:218-220 `handle.callBoxed`,
 * :256-272 `Op::materialize`, :506-528 `OpNode::materialize`).  Here the whole
 * recorded expression of one tensor (empty -> uniform_/normal_/fill_/zero_ ->
 * mul/add/erfinv/clamp ...) is folded by the host-side planner into ONE
 * `TdxInitDesc`, and a table of descriptors is executed by one persistent
 * kernel launch per kernel family.
 *
 * Rules of the ABI: plain C, no torch types, no exceptions, no allocation.
 * Every buffer (destination tensors, device workspace) is owned by the caller.
 * All functions return 0 on success or a (positive) cudaError_t value /
 * negative TDX_E* code on failure; tdx_last_error() gives a message.
 *
 * Random stream specification (normative; restated on the CPU in
 * oracle/tdx_oracle.c and pinned by tests/golden):
 *
 *   Philox4x32-R (R = 10 unless TDX_ALGO_*_R7), key = (seed_lo, seed_hi),
 *   counter = (off_lo, blk_hi, off_hi | 0x80000000, blk_lo)
 *   where off = TdxInitDesc.philox_offset (the torch generator offset at the
 *   time the op was issued == unique id of this RNG op under that seed) and
 *   blk = floor(g / EPB) for GLOBAL linear element index g of the unsharded
 *   tensor.  EPB = 8 for 16-bit outputs (16 random bits per element), 4 for
 *   32-bit outputs and for 16-bit outputs generated "wide" (TDX_ALGO_WIDE32).  Bit 31 of counter.z
 *   keeps the stream disjoint from ATen's own (offset, thread-id) use of the same generator
 *   (ATen/native/cuda/DistributionTemplates.h:72-88: ATen's third counter word is the high half of a
 *   thread index, far below 2^31); bits 30/29 of counter.z select the two tail-refinement blocks of
 *   the 16-bit normal.  The per-element word blk_lo is the LAST counter word: a Philox round only
 *   XORs and moves it, so the first multiplications of a block are per-tile constants (see
 *   tdx_init_kernels.cu philox_block); every bit of it still passes nine multiplying rounds.
 *   Because the value of element g depends only on (seed, off, g), any
 *   partition of [0, numel) over ranks reproduces the unsharded tensor
 *   bit-for-bit.
 */
#ifndef TDX_INIT_H_
#define TDX_INIT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define TDX_C_API __attribute__((visibility("default")))
#else
#define TDX_C_API
#endif

#define TDX_ABI_VERSION 2 /* 2: TdxPlan is 4 KiB; TDX_SRC_IOTA, TDX_I64, TDX_EPI_RPOW/RECIP; counter layout of the stream */

/* element type of the destination */
enum {
  TDX_F32 = 0,
  TDX_BF16 = 1,
  TDX_F16 = 2,
  TDX_I64 = 3, /* only valid with TDX_SRC_IOTA (arange buffers: position ids) */
  /* raw widths: only valid with TDX_SRC_CONST (bit-pattern fill) */
  TDX_RAW8 = 8,
  TDX_RAW16 = 9,
  TDX_RAW32 = 10,
  TDX_RAW64 = 11,
};

/* source of the expression (what the LAST full-tensor writer produces) */
enum {
  TDX_SRC_CONST = 0,   /* fill_/zero_/zeros/ones/full, constant-folded chains */
  TDX_SRC_UNIFORM = 1, /* uniform_(from,to), rand, kaiming_uniform_, xavier_uniform_ */
  TDX_SRC_NORMAL = 2,  /* normal_(mean,std), randn, kaiming_normal_, xavier_normal_ */
  /* element g = p0 + g * p1 (both integers, exactly representable in the destination), then the
   * epilogue: `arange` buffers and the index programs built on them -- rotary inv_freq is
   * arange -> float -> / dim -> base ** x -> reciprocal -> * scale.  dtype TDX_F32, or TDX_BF16 /
   * TDX_F16 (the fp32 value rounded once, to nearest even, at the store: `inv_freq.to(bf16)`)
   * -- epilogue allowed -- or TDX_I64 (none).  The reference replays these op by op through ATen
   * (deferred_init.cc:256-272); here they are one descriptor in the module's launch. */
  TDX_SRC_IOTA = 3,
};

/* sampling algorithm.  0 = the shipped default for (src, dtype). */
enum {
  TDX_ALGO_DEFAULT = 0,
  TDX_ALGO_ICDF16 = 1, /* 16-bit dtypes: inverse-CDF on 16 random bits/element + tail refinement */
  TDX_ALGO_BM32 = 2,   /* Box-Muller on 32 random bits/element (f32 default) */
  TDX_ALGO_WIDE32 = 2, /* same value: for 16-bit outputs of EITHER source, use the fp32 stream and
                        * arithmetic (EPB = 4) and round once at the store -- the meaning of
                        * `fp32_tensor.uniform_()/normal_()` followed by `.to(bf16/fp16)` */
  TDX_ALGO_BM16 = 3,   /* experimental: Box-Muller on 16-bit pairs, no tail refinement */
  TDX_ALGO_R7 = 0x10,  /* OR-able flag, experimental: Philox4x32-7 instead of -10 */
  TDX_ALGO_NOLUT = 0x20, /* OR-able flag: never use the shared-memory-table twin of ICDF16
                          * (same bits either way; for A/B measurements) */
};

/* epilogue steps, applied in order after the source transform; every step
 * rounds its result to the destination dtype, exactly like a chain of in-place
 * ATen ops on a tensor of that dtype does. */
enum {
  TDX_EPI_MUL = 1,    /* x = x * a            (mul_, mul.Tensor/Scalar)  */
  TDX_EPI_ADD = 2,    /* x = x + a            (add_, add.Tensor/Scalar with alpha folded) */
  TDX_EPI_ERFINV = 3, /* x = erfinv(x)        (trunc_normal_)            */
  TDX_EPI_CLAMP = 4,  /* x = min(max(x,a),b)  (clamp_)                   */
  TDX_EPI_RPOW = 5,   /* x = powf(a, x)       (pow.Scalar: scalar ** tensor) */
  TDX_EPI_RECIP = 6,  /* x = 1 / x            (reciprocal; `1.0 / t`)    */
  /* (`t / c` with a scalar c is TDX_EPI_MUL by the fp32 reciprocal of c: what ATen's CUDA kernel
   * computes, $TORCH/.../cuda/BinaryDivTrueKernel.cu) */
};
/* OR-able into TdxEpiStep.op: do not round to the destination dtype after this step (the step
 * was recorded on an fp32 intermediate that is cast to the destination dtype later). */
#define TDX_EPI_NOROUND 0x100u
/* TdxInitDesc.reserved bit 0: the source op itself produced an fp32 intermediate (do not round
 * the generated value to the destination dtype before the first epilogue step). */
#define TDX_FLAG_SRC_NOROUND 0x1u
#define TDX_MAX_EPI 4

typedef struct TdxEpiStep {
  uint32_t op;
  float a;
  float b;
} TdxEpiStep;

/* One fused per-tensor (or per-shard) initialisation program.  128 bytes. */
typedef struct TdxInitDesc {
  void* dst;              /* device address where element `elem_begin` is written */
  uint64_t elem_begin;    /* global linear index (unsharded tensor) of the first element written */
  uint64_t elem_count;    /* number of consecutive elements written */
  uint64_t philox_seed;   /* generator seed           (RNG sources) */
  uint64_t philox_offset; /* generator offset at issue (RNG sources) */
  double p0;              /* UNIFORM: from   NORMAL: mean */
  double p1;              /* UNIFORM: to     NORMAL: std  */
  uint64_t fill_bits[2];  /* CONST: 16-byte store pattern (element bits replicated) */
  uint8_t dtype;          /* TDX_F32 ... */
  uint8_t src;            /* TDX_SRC_* */
  uint8_t algo;           /* TDX_ALGO_* */
  uint8_t n_epi;          /* number of valid epilogue steps */
  uint32_t reserved;
  TdxEpiStep epi[TDX_MAX_EPI];
} TdxInitDesc;

/* Bytes of device workspace tdx_init_launch() needs for `n` descriptors. */
TDX_C_API size_t tdx_init_workspace_bytes(int n);

/*
 * Validates and executes `n` descriptors on `stream` (a cudaStream_t passed as
 * void*; NULL = legacy default stream) of the CURRENT device.
 * `descs` is host memory; `workspace` is device memory of at least
 * tdx_init_workspace_bytes(n) bytes that must stay untouched until the
 * launches complete.  Asynchronous w.r.t. the host.
 * Replaces: the per-op replay loop of OpNode::materialize
 * (reference deferred_init.cc:506-528).
 */
TDX_C_API int tdx_init_launch(const TdxInitDesc* descs, int n, void* workspace,
                              size_t workspace_bytes, void* stream);

/*
 * The same in two steps, for callers that want to allocate exactly what a table
 * needs instead of the upper bound of tdx_init_workspace_bytes(n) (which has to
 * allow for the table kernels' work lists: ~1.6 MB): tdx_init_prepare validates
 * and lays the plan out on the host and returns its size (0: nothing to launch);
 * tdx_init_submit copies it into `workspace` and launches.  The prepared plan
 * belongs to the calling thread and is consumed by the next tdx_init_submit.
 */
TDX_C_API int tdx_init_prepare(const TdxInitDesc* descs, int n, size_t* workspace_bytes);
TDX_C_API int tdx_init_submit(void* workspace, size_t workspace_bytes, void* stream);

/*
 * Two-phase variant for callers that re-launch one plan (benchmarks, CUDA
 * graphs): upload once, launch many times.  `tdx_plan_upload` copies the
 * grouped descriptor table into `workspace`; `tdx_plan_launch` only issues the
 * kernels (the work counters in the workspace are uploaded as zero and every
 * kernel puts its own back to zero when its last CTA leaves, so there is no
 * memset between launches; one plan must not run on two streams at once).
 */
typedef struct TdxPlan {
  uint64_t opaque[512]; /* host-side copy of the plan header; owned by the caller */
} TdxPlan;
TDX_C_API int tdx_plan_upload(const TdxInitDesc* descs, int n, void* workspace,
                              size_t workspace_bytes, void* stream, TdxPlan* plan);
TDX_C_API int tdx_plan_launch(const TdxPlan* plan, void* workspace, void* stream);

/* Number of kernel launches the last tdx_plan_launch/tdx_init_launch on this
 * thread issued (one per non-empty kernel family). */
TDX_C_API int tdx_last_launch_count(void);

/* Bytes of the plan image (descriptor table, prefix sums, work lists) the last
 * tdx_plan_upload/tdx_init_launch on this thread copied host -> device. */
TDX_C_API size_t tdx_last_upload_bytes(void);

/* Elements of the global tensor covered by one Philox block for this dtype/algo
 * (the planner rounds RNG consumption with it). */
TDX_C_API int tdx_elems_per_block(int dtype, int src, int algo);

TDX_C_API int tdx_abi_version(void);
TDX_C_API const char* tdx_last_error(void);

#define TDX_E_BADARG (-1)
#define TDX_E_WORKSPACE (-2)
#define TDX_E_NODEVICE (-3)

#ifdef __cplusplus
}
#endif
#endif /* TDX_INIT_H_ */
