// C++ surface of deferred initialisation, for hosts that link torchdistx_b200/_C.so directly.
//
// Same names, signatures and behaviour as the reference's installed header
// (reference src/cc/torchdistx/deferred_init.h:25-37).  What differs is underneath: a recorded
// CUDA tensor is not replayed op by op through the dispatcher (reference deferred_init.cc:256-272)
// but folded into descriptors of include/tdx_init.h and written by one kernel launch.
#pragma once

#include <ATen/Tensor.h>
#include <c10/core/Device.h>
#include <c10/core/DispatchKey.h>
#include <c10/core/impl/LocalDispatchKeySet.h>

#include <cstdint>
#include <optional>

#ifndef TDX_API
#define TDX_API __attribute__((visibility("default")))
#endif

namespace torchdistx {

// Forces all newly-constructed tensors on the calling thread to be fake while recording every
// operation performed on them; such tensors are materialised later by materializeTensor().  Nests.
TDX_API void enterDeferredInit();
TDX_API void leaveDeferredInit() noexcept;

// Whether `tensor` was constructed in a deferred-init context (and can be materialised).
TDX_API bool canMaterialize(const at::Tensor& tensor) noexcept;

// Materialises `tensor`; a tensor that is not deferred is returned as it is.  Materialising the same
// tensor again returns the same TensorImpl.  Errors as the reference: c10::ValueError for a fake
// argument that does not come from a deferred-init context, c10::Error (RuntimeError) for an
// external tensor mutated since the recording or an inference tensor (deferred_init.cc:231-246,
// 647-654).
TDX_API at::Tensor materializeTensor(const at::Tensor& tensor);

// Temporarily disables deferred-init on the calling thread (reference deferred_init.h:33-35).
class NoDeferredInit {
  c10::impl::ExcludeDispatchKeyGuard guard_{c10::DispatchKey::DeferredInit};
};

// ---- beyond the reference --------------------------------------------------------------------
// Build on another device than the recorded one (record on cpu, materialise on cuda:3), and/or only
// this rank's dim-0 chunk (the torch.chunk / FSDP2 Shard(0) layout).  `world` <= 1: whole tensor.
TDX_API at::Tensor materializeTensor(const at::Tensor& tensor, std::optional<c10::Device> device,
                                     int64_t rank = 0, int64_t world = 1);

}  // namespace torchdistx
