// Deferred-init recording: a flat, append-only tape of operator records.
//
// The reference records a pointer graph of OpNode objects holding std::function closures over
// the dispatcher (reference src/cc/torchdistx/deferred_init.cc:157-301 Op, :313-402 OpNode) and
// materialises a tensor by walking dependency / dependent edges (:530-622).  Here recording is
// an arena ("Tape"): every intercepted operator appends one TapeOp that names its tensor inputs
// and outputs by integer value ids and its storages by integer storage ids.  A tape is a plain
// data structure the planner (planner.h) can analyse -- last-writer analysis, dead-op
// elimination, pattern matching to fused sm_100a kernels -- without touching the dispatcher.
// Ops that the planner does not understand are still replayable through the dispatcher
// (generic replay), which keeps the reference's semantics for the long tail.
//
// Observable behaviour kept from the reference:
//   * ops are recorded only if they consume or produce a fake tensor   (deferred_init.cc:790-796)
//   * `aten::item` is terminal: fake arguments are materialised first  (:773-781, :813-825)
//   * external (real) tensor arguments are version-checked at replay   (:482-489, :652-654)
//   * Tensor.data get/set are recorded through the autograd hooks      (:889-948, :1050-1073)
//   * thread-local state at record time is restored at replay          (:205-215, :256-272)
//   * modes are per-thread and nestable; op order is per-thread        (:672, :1135-1161)
#pragma once

#include <ATen/Tensor.h>
#include <ATen/ThreadLocalState.h>
#include <ATen/core/dispatch/Dispatcher.h>
#include <ATen/core/ivalue.h>
#include <c10/core/Storage.h>
#include <c10/util/SmallVector.h>

#include <cstdint>
#include <memory>
#include <optional>
#include <string>
#include <unordered_map>
#include <vector>

namespace tdx {

constexpr uint32_t kNoValue = 0xffffffffu;

// What the planner knows about an operator (classified once per schema).
enum class OpKind : uint8_t {
  Generic = 0,
  // factories (new storage)
  Empty, Zeros, Ones, Full, Randn, Rand, Arange,
  // full aliases of the input (same storage, same elements)
  Alias,  // detach, alias, view-like ops are classified Alias only when they cover the storage
  // a view of PART of the input's storage (select / narrow / slice / a[i]): same memory, nothing
  // written; later in-place ops through it are partial writes of the storage
  View,
  // in-place writers
  UniformInplace, NormalInplace, FillInplace, ZeroInplace,
  MulInplace, AddInplace, ErfinvInplace, ClampInplace,
  // out-of-place unary elementwise (new storage, same geometry)
  MulOut, AddOut, CastOut, CloneOut, DivOut, PowScalarOut, ReciprocalOut,
  // dst.copy_(src): dst's elements become src's
  CopyInplace,
  // autograd hook pseudo-ops
  HookVariableData, HookSetData,
  // more affine maps by a number: x - c, -x, x / c (folded as + (-c), * (-1), * (1 / c))
  SubInplace, NegInplace, DivInplace, SubOut, NegOut,
};

struct Tape;

// Geometry of one recorded tensor value, captured from the meta twin at record time.
struct ValueInfo {
  uint32_t op = 0;        // producing op (index into Tape::ops)
  uint32_t storage = 0;   // index into Tape::storages
  c10::ScalarType dtype = c10::ScalarType::Undefined;
  c10::Device device = c10::Device(c10::kCPU);
  c10::SmallVector<int64_t, 4> sizes, strides;  // inline: materialising reads them on a cold cache
  int64_t storage_offset = 0;
  int64_t numel = 0;
  bool covers_storage = false;  // contiguous, offset 0, numel*itemsize == storage bytes
  bool requires_grad = false;
  at::Tensor real;        // set once materialised
  at::Tensor py_wrapped;  // what the Python binding returned for it (keeps object identity stable)
};

struct StorageTemplate;  // planner.cc: what the analysis at the end of the recording found

struct StorageInfo {
  c10::Storage meta;      // keeps the meta StorageImpl (our identity key) alive
  size_t nbytes = 0;
  // every op with an input or output on this storage, in order (inline: a parameter's storage sees
  // ~5 ops, and a recording is ~10^4 small allocations that somebody has to free again)
  c10::SmallVector<uint32_t, 6> touching_ops;
  at::Tensor base;        // real backing tensor once the fused path materialised the storage
  bool fused_done = false;
  bool base_taken = false;  // `base` itself has been handed out as some value's tensor
  bool base_is_shard = false;  // `base` holds one rank's dim-0 chunk only (shard=(r, W), W > 1)
  bool replayed = false;    // an op writing this storage went through generic replay: never fuse it afterwards
  int32_t live = 0;         // fake tensors (TensorRecords) that currently name a value on this storage
  uint64_t fused_epoch = 0;  // submission epoch of the batch that holds (held) its descriptor
  at::Tensor full_base;   // unsharded copy built for a reader of a storage whose `base` is a shard
  // Symbolic state of the storage, computed once when the outermost deferred_init scope ends
  // (analyze_tape): materialising is then allocation + descriptor fill, not program analysis.
  std::shared_ptr<const StorageTemplate> tmpl;
};

// A tensor argument of a recorded op.
struct InputRef {
  uint32_t value = kNoValue;          // value id in this tape, or kNoValue
  std::shared_ptr<Tape> foreign;      // set if the fake argument was recorded on another tape
  uint32_t foreign_value = kNoValue;
  at::Tensor real;                    // a real (external) tensor argument
  int64_t real_version = 0;           // its version counter at record time
};

struct TapeOp {
  std::optional<c10::OperatorHandle> handle;  // empty for hook pseudo-ops
  OpKind kind = OpKind::Generic;
  uint64_t seq = 0;                     // thread-wide chronological number
  c10::SmallVector<c10::IValue, 6> args;  // deep-copied call frame; fake tensors replaced by undefined
  c10::SmallVector<InputRef, 2> inputs;   // one per tensor slot of `args`, in stack_walk order
  c10::SmallVector<uint32_t, 2> outputs;  // value id per tensor output (kNoValue for non-fake outputs)
  uint32_t num_returns = 0;
  // thread-local state at record time; consecutive ops recorded under the same grad-mode /
  // autocast / dispatch-key state share one snapshot
  std::shared_ptr<const at::ThreadLocalState> tls;
  std::vector<c10::IValue> results;     // real outputs after generic replay
  bool done = false;
  uint32_t rng_slot = kNoValue;  // RNG ops the planner has met: index into Tape::rng
  const char* name() const;
};

// RNG ops on the fused path: the Philox stream id they were given (once, in materialise order).
// Kept in a dense array of its own: a materialise call touches two of these per tensor and nothing
// else of the ops.
struct RngSlot {
  uint32_t op = kNoValue;
  bool assigned = false;
  bool explicit_generator = false;  // the op was recorded with a `generator=` argument
  uint64_t seed = 0, offset = 0;
};

struct Tape : std::enable_shared_from_this<Tape> {
  uint64_t uid = 0;  // process-wide number of the recording (makes op ids unique across recordings)
  std::vector<RngSlot> rng;
  std::vector<TapeOp> ops;
  std::vector<ValueInfo> values;
  std::vector<StorageInfo> storages;
  std::unordered_map<const c10::StorageImpl*, uint32_t> storage_ids;
  uint32_t storage_id(const c10::Storage& s);
  // From the analysis at the end of the recording: bytes / number of the storages that the fused
  // path can build and that a fake tensor still names -- an estimate of what materialising the whole
  // recording will write (the planner sizes its submissions with it).
  uint64_t fused_bytes = 0;
  uint32_t fused_storages = 0;
};

// Recording state attached to a fake tensor (FakeTensorImpl::record()).
struct TensorRecord {
  std::shared_ptr<Tape> tape;
  uint32_t value = kNoValue;  // the tensor's CURRENT value (updated by in-place ops)
  // Re-points the record; keeps StorageInfo::live (how many fake tensors can still ask for a
  // storage's content) in step.
  void point_at(std::shared_ptr<Tape> t, uint32_t v) {
    if (tape && value != kNoValue) tape->storages[tape->values[value].storage].live--;
    tape = std::move(t);
    value = v;
    if (tape && value != kNoValue) tape->storages[tape->values[value].storage].live++;
  }
  TensorRecord() = default;
  TensorRecord(const TensorRecord&) = delete;
  TensorRecord& operator=(const TensorRecord&) = delete;
  ~TensorRecord() {
    if (tape && value != kNoValue) tape->storages[tape->values[value].storage].live--;
  }
};

// ---- runtime API (mirrors reference src/cc/torchdistx/deferred_init.h:25-37) -----------------
void enter_deferred_init();
void leave_deferred_init() noexcept;
bool can_materialize(const at::Tensor& t) noexcept;

struct NoDeferredInit {
  c10::impl::ExcludeDispatchKeyGuard guard{c10::DispatchKey::DeferredInit};
};

OpKind classify(const c10::OperatorHandle& op);

// planner.cc: evaluates every storage of a finished tape symbolically and caches the result on it
// (called by leave_deferred_init; never throws -- a storage whose analysis fails is analysed again,
// and reports its error, when it is materialised).
void analyze_tape(Tape& tape) noexcept;

// Lets go of every tensor a tape that is about to die still refers to (materialised outputs, replay
// results, real arguments).  For the thread that holds the GIL: since torch 2.10 dropping the last
// C++ reference to a tensor that has a Python wrapper (refcount 2 -> 1) hands ownership back to the
// wrapper with a Py_DECREF -- which takes the GIL.  A background thread tearing down a recording of
// 300 materialised parameters would fight the caller for it 300 times.
void drop_tensor_refs(Tape& tape) noexcept;

}  // namespace tdx
