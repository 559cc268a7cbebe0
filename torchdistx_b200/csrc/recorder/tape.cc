#include "tape.h"

#include <ATen/autocast_mode.h>
#include <ATen/core/VariableHooksInterface.h>
#include <c10/core/AutogradState.h>
#include <c10/core/impl/LocalDispatchKeySet.h>
#include <torch/library.h>

#include <atomic>
#include <mutex>

#include "fake_tensor.h"
#include "planner.h"
#include "stack_walk.h"

namespace tdx {

using c10::DispatchKey;
using c10::DispatchKeySet;
using c10::IValue;
using c10::OperatorHandle;
using torch::jit::Stack;

// ---------------------------------------------------------------------------------------------
// operator classification (once per schema)
// ---------------------------------------------------------------------------------------------
namespace {

struct OpInfo {
  OpKind kind = OpKind::Generic;
  bool terminal = false;  // needs real inputs to produce its (non-tensor) result
};

OpKind kind_from_name(const std::string& name, const std::string& overload) {
  // name is "aten::xyz"
  if (name.rfind("aten::", 0) != 0) return OpKind::Generic;
  const std::string n = name.substr(6);
  if (n == "empty" || n == "empty_strided" || n == "empty_like" || n == "new_empty" ||
      n == "new_empty_strided")
    return OpKind::Empty;
  if (n == "zeros" || n == "zeros_like" || n == "new_zeros") return OpKind::Zeros;
  if (n == "ones" || n == "ones_like" || n == "new_ones") return OpKind::Ones;
  if (n == "full" || n == "full_like" || n == "new_full") return OpKind::Full;
  if (n == "randn" || n == "randn_like") return OpKind::Randn;
  if (n == "rand" || n == "rand_like") return OpKind::Rand;
  if (n == "detach" || n == "alias") return OpKind::Alias;
  if (n == "uniform_") return OpKind::UniformInplace;
  if (n == "normal_") return OpKind::NormalInplace;
  if (n == "fill_") return OpKind::FillInplace;
  if (n == "zero_") return OpKind::ZeroInplace;
  if (n == "mul_") return OpKind::MulInplace;
  if (n == "add_") return OpKind::AddInplace;
  if (n == "sub_" && (overload == "Tensor" || overload == "Scalar")) return OpKind::SubInplace;
  if (n == "neg_") return OpKind::NegInplace;
  if (n == "div_" && (overload == "Tensor" || overload == "Scalar")) return OpKind::DivInplace;  // (no rounding_mode)
  if (n == "sub" && (overload == "Tensor" || overload == "Scalar")) return OpKind::SubOut;
  if (n == "neg") return OpKind::NegOut;
  if (n == "erfinv_") return OpKind::ErfinvInplace;
  if (n == "clamp_" && overload.empty()) return OpKind::ClampInplace;
  if (n == "mul" && (overload == "Tensor" || overload == "Scalar")) return OpKind::MulOut;
  if (n == "add" && (overload == "Tensor" || overload == "Scalar")) return OpKind::AddOut;
  if (n == "clone") return OpKind::CloneOut;
  if (n == "arange") return OpKind::Arange;
  if (n == "div" && (overload == "Tensor" || overload == "Scalar")) return OpKind::DivOut;
  if (n == "pow" && overload == "Scalar") return OpKind::PowScalarOut;
  if (n == "reciprocal") return OpKind::ReciprocalOut;
  if (n == "copy_") return OpKind::CopyInplace;
  if (n == "_to_copy" || (n == "to" && (overload == "dtype" || overload == "dtype_layout")))
    return OpKind::CastOut;
  return OpKind::Generic;
}

const OpInfo& info_of(const OperatorHandle& op) {
  thread_local std::unordered_map<const c10::FunctionSchema*, OpInfo> cache;
  const c10::FunctionSchema* key = &op.schema();
  auto it = cache.find(key);
  if (it != cache.end()) return it->second;
  OpInfo i;
  i.kind = kind_from_name(key->name(), key->overload_name());
  i.terminal = key->name() == "aten::item" || key->name() == "aten::_local_scalar_dense";
  return cache.emplace(key, i).first->second;
}

}  // namespace

OpKind classify(const OperatorHandle& op) { return info_of(op).kind; }

const char* TapeOp::name() const {
  if (handle) return handle->schema().name().c_str();
  return kind == OpKind::HookSetData ? "VariableHooks::set_data" : "VariableHooks::variable_data";
}

uint32_t Tape::storage_id(const c10::Storage& s) {
  const c10::StorageImpl* key = s.unsafeGetStorageImpl();
  auto it = storage_ids.find(key);
  if (it != storage_ids.end()) return it->second;
  const uint32_t id = static_cast<uint32_t>(storages.size());
  StorageInfo info;
  info.meta = s;
  info.nbytes = s.nbytes();
  storages.push_back(std::move(info));
  storage_ids.emplace(key, id);
  return id;
}

// ---------------------------------------------------------------------------------------------
// recording
// ---------------------------------------------------------------------------------------------
namespace {

thread_local uint64_t tls_seq = 0;             // chronological op number of this thread
thread_local size_t tls_level = 0;             // deferred_init nesting
thread_local std::shared_ptr<Tape> tls_tape;   // tape of the outermost active deferred_init
thread_local uint64_t tls_session = 0;         // bumped by every outermost enter_deferred_init

// Thread-local state (grad mode, autocast, dispatch keys, ...) is restored when an op is replayed
// (reference deferred_init.cc:205-215, 256-272).  Snapshots are expensive to take and to destroy
// and almost never change between consecutive ops: reuse the previous one while a cheap
// fingerprint of the state is unchanged.
std::shared_ptr<const at::ThreadLocalState> current_tls_snapshot() {
  struct Fingerprint {
    uint64_t included = 0, excluded = 0;
    bool grad = false, inference = false, fw_grad = false, multithreading = false, view_replay = false;
    uint64_t autocast = 0;
    bool operator==(const Fingerprint& o) const {
      return included == o.included && excluded == o.excluded && grad == o.grad &&
             inference == o.inference && fw_grad == o.fw_grad && multithreading == o.multithreading &&
             view_replay == o.view_replay && autocast == o.autocast;
    }
  };
  // a handful of states alternate during module construction (grad mode on for constructors, off
  // inside nn.init.*): keep one snapshot per state
  struct Slot {
    Fingerprint fp;
    std::shared_ptr<const at::ThreadLocalState> tls;
  };
  thread_local std::vector<Slot> slots;
  thread_local uint64_t last_session = ~0ull;
  if (last_session != tls_session) {  // never carry a snapshot over from an earlier deferred_init
    slots.clear();
    last_session = tls_session;
  }
  Fingerprint fp;
  const auto ks = c10::impl::tls_local_dispatch_key_set();
  fp.included = ks.included_.raw_repr();
  fp.excluded = ks.excluded_.raw_repr();
  const auto ag = c10::AutogradState::get_tls_state();
  fp.grad = ag.get_grad_mode();
  fp.inference = ag.get_inference_mode();
  fp.fw_grad = ag.get_fw_grad_mode();
  fp.multithreading = ag.get_multithreading_enabled();
  fp.view_replay = ag.get_view_replay_enabled();
  for (int dt = 0; dt < static_cast<int>(c10::DeviceType::COMPILE_TIME_MAX_DEVICE_TYPES); ++dt) {
    const auto d = static_cast<c10::DeviceType>(dt);
    if (d == c10::kCPU || d == c10::kCUDA)
      fp.autocast = fp.autocast * 131 + (at::autocast::is_autocast_enabled(d) ? 1 + static_cast<int>(at::autocast::get_autocast_dtype(d)) : 0);
  }
  for (const Slot& sl : slots)
    if (sl.fp == fp) return sl.tls;
  if (slots.size() >= 8) slots.erase(slots.begin());
  slots.push_back(Slot{fp, std::make_shared<const at::ThreadLocalState>()});
  return slots.back().tls;
}

// Call frames are replayed much later: deep-copy containers so that a caller mutating its list
// cannot change the recording, and refuse values whose state we cannot freeze.
IValue freeze(const IValue& v, const char* op_name) {
  if (v.isTensor() || v.isNone() || v.isBool() || v.isInt() || v.isDouble() || v.isString() ||
      v.isDevice() || v.isGenerator() || v.isComplexDouble() || v.isSymInt() || v.isEnum()) {
    return v;
  }
  if (v.isList()) {
    const auto& src = v.toList();
    c10::impl::GenericList out(src.elementType());
    out.reserve(src.size());
    for (const IValue& e : v.toListRef()) out.push_back(freeze(e, op_name));
    return out;
  }
  if (v.isTuple()) {
    std::vector<IValue> elems;
    for (const IValue& e : v.toTupleRef().elements()) elems.push_back(freeze(e, op_name));
    return c10::ivalue::Tuple::create(std::move(elems));
  }
  if (v.isGenericDict()) {
    const auto& src = v.toGenericDict();
    c10::impl::GenericDict out(src.keyType(), src.valueType());
    for (const auto& kv : src) out.insert(freeze(kv.key(), op_name), freeze(kv.value(), op_name));
    return out;
  }
  TORCH_CHECK(false, "`", op_name, "` has an argument of type `", v.type()->str(),
              "` which is not supported in a deferred-init context.");
}

ValueInfo describe(const at::Tensor& fake, uint32_t op, uint32_t storage, size_t storage_bytes) {
  ValueInfo v;
  v.op = op;
  v.storage = storage;
  v.dtype = fake.scalar_type();
  v.device = fake.device();
  v.sizes.assign(fake.sizes().begin(), fake.sizes().end());
  v.strides.assign(fake.strides().begin(), fake.strides().end());
  v.storage_offset = fake.storage_offset();
  v.numel = fake.numel();
  v.covers_storage = fake.is_contiguous() && v.storage_offset == 0 &&
                     static_cast<size_t>(v.numel) * fake.element_size() == storage_bytes;
  return v;
}

void touch(Tape& tape, uint32_t storage, uint32_t op) {
  auto& t = tape.storages[storage].touching_ops;
  if (t.empty() || t.back() != op) t.push_back(op);
}

// Appends one record.  `frame` is the frozen argument frame, `outputs` the live result stack.
void append(std::optional<OperatorHandle> handle, OpKind kind, Stack frame, size_t nargs,
            Stack& outputs, size_t nret) {
  if (!tls_tape) {
    static std::atomic<uint64_t> next_uid{1};
    tls_tape = std::make_shared<Tape>();
    tls_tape->uid = next_uid.fetch_add(1);
  }
  Tape& tape = *tls_tape;
  const uint32_t op_idx = static_cast<uint32_t>(tape.ops.size());
  tape.ops.emplace_back();
  TapeOp* op = &tape.ops.back();
  op->handle = std::move(handle);
  op->kind = kind;
  op->seq = tls_seq++;
  op->num_returns = static_cast<uint32_t>(nret);
  op->tls = current_tls_snapshot();

  c10::SmallVector<const c10::TensorImpl*, 4> fake_inputs;
  for_each_tensor_mut(frame, nargs, [&](at::Tensor& t) {
    InputRef in;
    if (is_fake(t)) {
      const auto& rec = fake_impl(t)->record();
      if (rec->tape.get() == &tape) {
        in.value = rec->value;
        touch(tape, tape.values[in.value].storage, op_idx);
      } else {
        in.foreign = rec->tape;
        in.foreign_value = rec->value;
      }
      fake_inputs.push_back(t.unsafeGetTensorImpl());
    } else if (t.defined()) {
      in.real = t;
      in.real_version = t.is_inference() ? 0 : static_cast<int64_t>(t._version());
    }
    op->inputs.push_back(std::move(in));
    t = at::Tensor();  // the frame never keeps tensors alive; InputRef does
  });
  op->args.reserve(frame.size());
  for (IValue& v : frame) op->args.push_back(std::move(v));

  for_each_tensor_mut(outputs, nret, [&](at::Tensor& t) {
    if (!is_fake(t)) {
      op->outputs.push_back(kNoValue);
      return;
    }
    const uint32_t sid = tape.storage_id(meta_storage(t));
    const uint32_t vid = static_cast<uint32_t>(tape.values.size());
    tape.values.push_back(describe(t, op_idx, sid, tape.storages[sid].nbytes));
    touch(tape, sid, op_idx);
    op->outputs.push_back(vid);
    auto* impl = fake_impl(t);
    if (!impl->record()) impl->set_record(std::make_shared<TensorRecord>());
    // in-place results keep their tensor (and record); the record now names the new value
    impl->record()->point_at(tls_tape, vid);
  });

  // A generic, non-mutating op whose single fake result lives on the storage of its first fake
  // argument is a view.  If it re-describes exactly the argument's elements (view / reshape /
  // flatten / unsqueeze of a contiguous tensor) it is a plain alias; otherwise (select / narrow /
  // slice / a[i] / t() ...) it names part of the storage -- nothing is read or written either way.
  // The same goes for a conversion that had nothing to convert: `x.to(dtype)` / `x.float()` /
  // `Module.to(device)` on a tensor that already is what was asked for returns `x` itself (the
  // CompositeImplicit `aten::to.*` is recorded whole, as the reference does) -- without this, every
  // tensor of `Model().float()` or `Model().to(device)` looked like a cast onto its own storage and
  // fell off the fused path.
  if ((kind == OpKind::Generic || kind == OpKind::CastOut) && op->outputs.size() == 1 && op->outputs[0] != kNoValue &&
      !op->inputs.empty() && op->inputs[0].value != kNoValue) {
    const ValueInfo& out = tape.values[op->outputs[0]];
    const ValueInfo& in = tape.values[op->inputs[0].value];
    bool only_one_tensor_input = true;
    for (size_t i = 1; i < op->inputs.size(); ++i)
      only_one_tensor_input &= !(op->inputs[i].value != kNoValue || op->inputs[i].foreign ||
                                 op->inputs[i].real.defined());
    if (only_one_tensor_input && out.storage == in.storage && out.dtype == in.dtype && op->handle &&
        op->handle->schema().is_mutable() == false) {
      op->kind = (out.covers_storage && in.covers_storage) ? OpKind::Alias : OpKind::View;
    }
  }
}

bool any_fake(const Stack& s, size_t n) {
  bool hit = false;
  for_each_tensor(s, n, [&](const at::Tensor& t) { return hit = is_fake(t); });
  return hit;
}

const DispatchKeySet kBelowDeferredInit{DispatchKeySet::FULL_AFTER, DispatchKey::DeferredInit};

void deferred_init_fallback(const OperatorHandle& op, DispatchKeySet ks, Stack* stack) {
  NoDeferredInit no_reentry;
  const auto& schema = op.schema();
  const size_t nargs = schema.arguments().size();
  const size_t nret = schema.returns().size();

  for_each_tensor(*stack, nargs, [&](const at::Tensor& t) {
    TORCH_CHECK_VALUE(!is_fake(t) || fake_impl(t)->record() != nullptr, "`", schema.name(),
                      "` has a fake `Tensor` argument which was not constructed in a deferred-init "
                      "context.");
  });

  // below DeferredInit, with Fake forced on so that new tensors come out fake
  const DispatchKeySet next = ks.add(DispatchKey::Fake) & kBelowDeferredInit;
  const OpInfo& info = info_of(op);

  if (info.terminal) {
    // e.g. Tensor.item(): the value is needed now -- materialise the arguments and run for real
    for_each_tensor_mut(*stack, nargs, [&](at::Tensor& t) {
      if (is_fake(t)) t = materialize_one(t, MaterializeOptions{});
    });
    op.redispatchBoxed(next, stack);
    return;
  }

  const bool fake_in = any_fake(*stack, nargs);
  Stack frame;
  frame.reserve(nargs);
  for (size_t i = 0; i < nargs; ++i)
    frame.push_back(freeze(torch::jit::peek(*stack, i, nargs), schema.name().c_str()));

  op.redispatchBoxed(next, stack);

  if (fake_in || any_fake(*stack, nret))
    append(op, info.kind, std::move(frame), nargs, *stack, nret);
}

}  // namespace
}  // namespace tdx

TORCH_LIBRARY_IMPL(_, DeferredInit, m) {
  m.fallback(torch::CppFunction::makeFromBoxedFunction<&tdx::deferred_init_fallback>());
}

// ---------------------------------------------------------------------------------------------
// Tensor.data get / set: not dispatcher ops, so they are observed through the autograd hooks
// ---------------------------------------------------------------------------------------------
namespace tdx {
namespace {

using at::TensorBase;
using at::impl::VariableHooksInterface;

bool recording_on_this_thread() noexcept {
  return c10::impl::tls_is_dispatch_key_included(DispatchKey::DeferredInit) &&
         !c10::impl::tls_is_dispatch_key_excluded(DispatchKey::DeferredInit);
}

void check_hook_arg(const char* hook, const TensorBase& t) {
  TORCH_CHECK_VALUE(!is_fake(t) || fake_impl(t)->record() != nullptr, "`VariableHooks::", hook,
                    "` has a fake `Tensor` argument which was not constructed in a deferred-init "
                    "context.");
}

// Forwards everything to autograd's own hooks; taps variable_data() and set_data().
class RecordingHooks final : public VariableHooksInterface {
 public:
  explicit RecordingHooks(VariableHooksInterface* inner) : inner_(inner) {}
  VariableHooksInterface* inner() const { return inner_; }

  TensorBase variable_data(const TensorBase& self) const override {
    const bool on = recording_on_this_thread();
    if (on) check_hook_arg("variable_data", self);
    TensorBase data = inner_->variable_data(self);
    if (on && is_fake(self) && is_fake(data)) {
      Stack frame{IValue(at::Tensor(self))}, out{IValue(at::Tensor(data))};
      append(std::nullopt, OpKind::HookVariableData, std::move(frame), 1, out, 1);
    }
    return data;
  }
  void set_data(const TensorBase& self, const TensorBase& data) const override {
    const bool on = recording_on_this_thread();
    if (on) {
      check_hook_arg("set_data", self);
      check_hook_arg("set_data", data);
    }
    const bool rec = on && is_fake(self) && is_fake(data);
    Stack frame;
    if (rec) frame = Stack{IValue(at::Tensor(self)), IValue(at::Tensor(data))};
    inner_->set_data(self, data);  // self now shares data's (meta) storage
    if (rec) {
      Stack out{IValue(at::Tensor(self))};
      append(std::nullopt, OpKind::HookSetData, std::move(frame), 2, out, 1);
    }
  }

  // pure forwarding
  TensorBase tensor_data(const TensorBase& s) const override { return inner_->tensor_data(s); }
  const std::shared_ptr<torch::autograd::Node>& grad_fn(const TensorBase& s) const override {
    return inner_->grad_fn(s);
  }
  unsigned _register_hook(const TensorBase& s,
                          std::function<TensorBase(const TensorBase&)> h) const override {
    return inner_->_register_hook(s, std::move(h));
  }
  void remove_hook(const TensorBase& s, unsigned pos) const override { inner_->remove_hook(s, pos); }
  bool is_view(const TensorBase& s) const override { return inner_->is_view(s); }
  const TensorBase& base(const TensorBase& s) const override { return inner_->base(s); }
  const std::string& name(const TensorBase& s) const override { return inner_->name(s); }
  bool is_leaf(const TensorBase& s) const override { return inner_->is_leaf(s); }
  int64_t output_nr(const TensorBase& s) const override { return inner_->output_nr(s); }
  TensorBase data(const TensorBase& s) const override { return inner_->data(s); }
  int64_t _version(const TensorBase& s) const override { return inner_->_version(s); }
  void retain_grad(const TensorBase& s) const override { inner_->retain_grad(s); }
  bool retains_grad(const TensorBase& s) const override { return inner_->retains_grad(s); }
  void _backward(const at::Tensor& s, at::TensorList inputs, const std::optional<at::Tensor>& g,
                 std::optional<bool> keep, bool create) const override {
    inner_->_backward(s, inputs, g, keep, create);
  }
  void requires_grad_(const TensorBase& s, bool v) const override { inner_->requires_grad_(s, v); }
  void basic_autograd_not_implemented_fallback(const c10::OperatorHandle& op, DispatchKeySet ks,
                                               Stack* stack) const override {
    inner_->basic_autograd_not_implemented_fallback(op, ks, stack);
  }
  std::optional<c10::ScalarType> grad_dtype(const TensorBase& s) const override {
    return inner_->grad_dtype(s);
  }
  void set_grad_dtype(const TensorBase& s, const std::optional<c10::ScalarType>& d) const override {
    inner_->set_grad_dtype(s, d);
  }

 private:
  VariableHooksInterface* inner_;
};

// The autograd hook table is process-global: install once for any number of recording threads.
std::mutex g_hooks_mutex;
size_t g_hooks_users = 0;
std::unique_ptr<RecordingHooks> g_hooks;

void install_hooks() {
  std::lock_guard<std::mutex> lock(g_hooks_mutex);
  if (g_hooks_users++ == 0) {
    g_hooks = std::make_unique<RecordingHooks>(at::impl::GetVariableHooks());
    at::impl::SetVariableHooks(g_hooks.get());
  }
}
void uninstall_hooks() noexcept {
  std::lock_guard<std::mutex> lock(g_hooks_mutex);
  if (g_hooks_users == 0) return;
  if (--g_hooks_users == 0) {
    at::impl::SetVariableHooks(g_hooks->inner());
    g_hooks.reset();
  }
}

}  // namespace

void enter_deferred_init() {
  if (++tls_level == 1) {
    tls_tape.reset();  // a fresh tape per outermost scope; old tapes live on in their tensors
    ++tls_session;
    c10::impl::tls_set_dispatch_key_included(DispatchKey::DeferredInit, true);
    install_hooks();
  }
}

void leave_deferred_init() noexcept {
  if (tls_level == 0) return;
  if (--tls_level == 0) {
    c10::impl::tls_set_dispatch_key_included(DispatchKey::DeferredInit, false);
    uninstall_hooks();
    if (tls_tape) analyze_tape(*tls_tape);
    tls_tape.reset();
  }
}

void drop_tensor_refs(Tape& tape) noexcept {
  for (ValueInfo& v : tape.values) {
    v.real.reset();
    v.py_wrapped.reset();
  }
  for (StorageInfo& s : tape.storages) {
    s.base.reset();
    s.full_base.reset();
  }
  for (TapeOp& op : tape.ops) {
    for (InputRef& in : op.inputs) in.real.reset();
    if (!op.results.empty()) op.results.clear();
  }
}

bool can_materialize(const at::Tensor& t) noexcept {
  return is_fake(t) && fake_impl(t)->record() != nullptr;
}

}  // namespace tdx
