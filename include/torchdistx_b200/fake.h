// C++ surface of the fake-tensor runtime, for hosts that link torchdistx_b200/_C.so directly.
//
// Same names, signatures and error behaviour as the reference's installed header
// (reference src/cc/torchdistx/fake.h:34-83), so a C++ caller of the reference switches by changing
// the include path and the library it links.  The two libraries cannot be loaded into one process:
// both register the boxed fallbacks of DispatchKey::Fake / DispatchKey::DeferredInit.
//
// Not provided: FakeTensor::setData / hasData / getData / unsafeGetData (reference fake.h:57-72).
// They are the extension point through which the reference's own deferred_init.cc hangs its
// recording off a fake tensor (a per-dispatch-key map of shared_ptr<void>); here the recording is a
// typed slot of the tensor (csrc/recorder/fake_tensor.h) and there is no generic map to expose.
#pragma once

#include <ATen/Tensor.h>
#include <c10/core/Storage.h>

#ifndef TDX_API
#define TDX_API __attribute__((visibility("default")))
#endif

namespace torchdistx {

// Forces all newly-constructed tensors on the calling thread to be fake; nests.  With `fake_cuda`,
// fake CUDA tensors can be constructed on a machine without CUDA (reference fake.h:38-42).
TDX_API void enterFakeMode(bool fake_cuda = false);
// Leaves the fake mode of the calling thread (one level).
TDX_API void leaveFakeMode() noexcept;
TDX_API bool isFakeModeActive() noexcept;
TDX_API bool isFake(const at::TensorBase& tensor) noexcept;

// Access to the properties of a fake tensor (reference fake.h:54-81).
class TDX_API FakeTensor {
 public:
  // Raises c10::ValueError ("`tensor` was expected to be a fake tensor.") unless `unsafe`.
  explicit FakeTensor(const at::TensorBase& tensor, bool unsafe = false);
  // A detached meta tensor with the same geometry and dtype.
  at::Tensor toMeta() const;
  // Identity of the (virtual) memory the tensor occupies: fake views of one tensor share it.
  const at::Storage& meta_storage() const noexcept;

 private:
  void* impl_;
};

TDX_API FakeTensor asFake(const at::TensorBase& tensor);
TDX_API FakeTensor unsafeAsFake(const at::TensorBase& tensor) noexcept;

}  // namespace torchdistx
