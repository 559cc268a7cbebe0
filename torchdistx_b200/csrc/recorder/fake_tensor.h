// Zero-storage ("fake") tensors behind DispatchKey::Fake.
//
// Behavioural contract (what a user of torchdistx.fake observes) follows the reference:
//   enter/leave fake mode, nesting, fake CUDA          reference src/cc/torchdistx/fake.cc:554-623
//   a fake tensor reports a real device, has no storage reference fake.cc:73-127, 152-165
//   ops run on the Meta backend for shape inference    reference fake.cc:318-336, 476-489
//   output-device heuristic (device arg > tensors > cpu) reference fake.cc:346-432
// The implementation is new: per-operator traits are computed once and cached, the
// deferred-init recorder hangs a typed record off the tensor instead of a generic
// per-dispatch-key map, and the meta twin is reused for in-place results by pointer lookup
// in a small inline array.
#pragma once

#include <ATen/Tensor.h>
#include <ATen/core/dispatch/Dispatcher.h>
#include <c10/core/Device.h>
#include <c10/core/Storage.h>
#include <c10/core/TensorImpl.h>

#include <cstdio>
#include <memory>
#include <string>

namespace tdx {

// A double as an ostream would print it ("%g").  Error messages are built from this: the
// extension's own instantiation of std::ostream::operator<<(double) is not usable (it is emitted
// into this hidden-visibility library without its locale facets and crashes).
inline std::string fmt_double(double v) {
  char buf[32];
  std::snprintf(buf, sizeof(buf), "%g", v);
  return buf;
}

struct TensorRecord;  // deferred-init recording state of one fake tensor (tape.h)

class FakeTensorImpl final : public c10::TensorImpl {
 public:
  // Builds a fake twin of `meta` that claims to live on `device`.
  static c10::intrusive_ptr<FakeTensorImpl> make(c10::intrusive_ptr<c10::TensorImpl> meta,
                                                 c10::Device device);

  // Re-reads sizes/strides/dtype from the meta twin (after an in-place op changed it).
  void sync_from_meta();

  const c10::intrusive_ptr<c10::TensorImpl>& meta() const noexcept { return meta_; }

  // deferred-init recording slot
  const std::shared_ptr<TensorRecord>& record() const noexcept { return record_; }
  void set_record(std::shared_ptr<TensorRecord> r) noexcept { record_ = std::move(r); }

  void shallow_copy_from(const c10::intrusive_ptr<c10::TensorImpl>& impl) override;
  c10::intrusive_ptr<c10::TensorImpl> shallow_copy_and_detach(
      const c10::VariableVersion& version_counter, bool allow_tensor_metadata_change) const override;
  c10::intrusive_ptr<c10::TensorImpl> shallow_copy_and_detach(
      c10::VariableVersion&& version_counter, bool allow_tensor_metadata_change) const override;
  void release_resources() override;

  FakeTensorImpl();  // use make()

 protected:
  const char* tensorimpl_type_name() const override { return "tdx::FakeTensorImpl"; }

 private:
  template <class VC>
  c10::intrusive_ptr<c10::TensorImpl> detach_impl(VC&& vc, bool allow_change) const;
  void adopt_metadata(const c10::TensorImpl& meta, c10::Device device, c10::DispatchKeySet ks);

  c10::intrusive_ptr<c10::TensorImpl> meta_;
  std::shared_ptr<TensorRecord> record_;
};

// ---- public runtime API (mirrors reference src/cc/torchdistx/fake.h:34-83) ------------------
void enter_fake_mode(bool fake_cuda);
void leave_fake_mode() noexcept;
bool fake_mode_active() noexcept;

inline bool is_fake(const at::TensorBase& t) noexcept {
  return t.defined() && t.key_set().has(c10::DispatchKey::Fake);
}
// Unchecked downcast; caller guarantees is_fake(t).
inline FakeTensorImpl* fake_impl(const at::TensorBase& t) noexcept {
  return static_cast<FakeTensorImpl*>(t.unsafeGetTensorImpl());
}
// A detached meta tensor with the fake's geometry; TORCH_CHECK_VALUE if `t` is not fake.
at::Tensor meta_like(const at::Tensor& t);
// Index of the schema argument that selects the output device of `op` (factory-like operators
// only), or -1.  Used by the replay engine to retarget a recording to another device.
int device_argument_index(const c10::OperatorHandle& op);
// The Storage of the meta twin: identity of the (virtual) memory a fake tensor occupies.
const c10::Storage& meta_storage(const at::TensorBase& fake);

}  // namespace tdx
