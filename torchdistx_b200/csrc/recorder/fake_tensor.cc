#include "fake_tensor.h"

#include <ATen/Functions.h>
#include <ATen/core/dispatch/Dispatcher.h>
#include <ATen/core/ivalue.h>
#include <ATen/core/stack.h>
#include <c10/core/TensorOptions.h>
#include <c10/core/impl/DeviceGuardImplInterface.h>
#include <c10/core/impl/LocalDispatchKeySet.h>
#include <c10/core/impl/VirtualGuardImpl.h>
#include <torch/library.h>

#include <array>
#include <unordered_map>

#include "stack_walk.h"

namespace tdx {

using c10::Device;
using c10::DispatchKey;
using c10::DispatchKeySet;
using c10::IValue;
using c10::OperatorHandle;
using c10::TensorImpl;
using torch::jit::Stack;

// ---------------------------------------------------------------------------------------------
// FakeTensorImpl
// ---------------------------------------------------------------------------------------------
FakeTensorImpl::FakeTensorImpl() : TensorImpl(DispatchKeySet{}, caffe2::TypeMeta{}, std::nullopt) {}

namespace {

// Dispatch keys of a tensor with `meta`'s dtype/layout that lives on `device`, plus Fake so that
// every op on it reaches our handler before any backend kernel.
DispatchKeySet keys_for(TensorImpl& meta, Device device) {
  const auto dtype = c10::typeMetaToScalarType(meta.dtype());
  DispatchKeySet ks{c10::computeDispatchKey(dtype, meta.layout(), device), DispatchKey::Fake};
  if (!meta.is_inference()) {
    const auto backend = ks.highestBackendKey();
    ks = ks | c10::getAutocastRelatedKeySetFromBackend(backend) |
         c10::getAutogradRelatedKeySetFromBackend(backend);
  }
  return ks;
}

}  // namespace

void FakeTensorImpl::adopt_metadata(const TensorImpl& meta, Device device, DispatchKeySet ks) {
  copy_tensor_metadata(&meta, this, version_counter_, allow_tensor_metadata_change_);
  // copy_tensor_metadata also copied the meta storage, device and key set: undo that, a fake
  // tensor owns no memory and answers for a real device.
  storage_ = {};
  storage_access_should_throw_ = true;
  device_opt_ = device;
  key_set_ = ks;
  refresh_numel();
  refresh_contiguous();
}

c10::intrusive_ptr<FakeTensorImpl> FakeTensorImpl::make(c10::intrusive_ptr<TensorImpl> meta,
                                                        Device device) {
  TORCH_INTERNAL_ASSERT(meta->is_meta(), "fake tensors are built from meta tensors");
  auto impl = c10::make_intrusive<FakeTensorImpl>();
  impl->adopt_metadata(*meta, device, keys_for(*meta, device));
  impl->meta_ = std::move(meta);
  return impl;
}

void FakeTensorImpl::sync_from_meta() { adopt_metadata(*meta_, *device_opt_, key_set_); }

void FakeTensorImpl::shallow_copy_from(const c10::intrusive_ptr<TensorImpl>& impl) {
  TORCH_CHECK(impl->key_set().has(DispatchKey::Fake),
              "The source tensor was expected to be a fake tensor.");
  const auto* src = static_cast<const FakeTensorImpl*>(impl.get());
  copy_tensor_metadata(src, this, version_counter_, allow_tensor_metadata_change_);
  refresh_numel();
  refresh_contiguous();
  meta_->shallow_copy_from(src->meta_);
}

template <class VC>
c10::intrusive_ptr<TensorImpl> FakeTensorImpl::detach_impl(VC&& vc, bool allow_change) const {
  auto impl = c10::make_intrusive<FakeTensorImpl>();
  copy_tensor_metadata(this, impl.get(), std::forward<VC>(vc), allow_change);
  impl->refresh_numel();
  impl->refresh_contiguous();
  // the detached twin shares the meta storage: that is how aliasing is tracked while recording
  impl->meta_ = meta_->shallow_copy_and_detach(/*version_counter=*/0, /*allow_tensor_metadata_change=*/false);
  return impl;
}

c10::intrusive_ptr<TensorImpl> FakeTensorImpl::shallow_copy_and_detach(
    const c10::VariableVersion& vc, bool allow_change) const {
  return detach_impl(vc, allow_change);
}
c10::intrusive_ptr<TensorImpl> FakeTensorImpl::shallow_copy_and_detach(c10::VariableVersion&& vc,
                                                                       bool allow_change) const {
  return detach_impl(std::move(vc), allow_change);
}

void FakeTensorImpl::release_resources() {
  TensorImpl::release_resources();
  meta_.reset();
  record_.reset();
}

at::Tensor meta_like(const at::Tensor& t) {
  TORCH_CHECK_VALUE(is_fake(t), "`tensor` was expected to be a fake tensor.");
  auto impl = fake_impl(t)->meta()->shallow_copy_and_detach(/*version_counter=*/0,
                                                           /*allow_tensor_metadata_change=*/false);
  impl->set_autograd_meta(nullptr);
  return at::Tensor(std::move(impl));
}

const c10::Storage& meta_storage(const at::TensorBase& fake) {
  return fake_impl(fake)->meta()->storage();
}

// ---------------------------------------------------------------------------------------------
// the handler behind DispatchKey::Fake
// ---------------------------------------------------------------------------------------------
namespace {

// What we need to know about an operator, computed once per schema.
struct OpTraits {
  int device_arg = -1;    // index of the argument that names the output device, or -1
  bool meta_ok = false;   // a Meta (or composite) kernel exists
  // `normal_(self, mean, std, *, generator)` / `uniform_(self, from, to, *, generator)`: the result
  // IS self, so there is nothing to infer -- and torch 2.11's Meta kernels for the two go through
  // Python (200 us and 14 us per call: two thirds of the time it takes to record Llama-3-8B)
  enum { None, NormalInplace, UniformInplace } inplace_rng = None;
};

bool has_tensor_options_quartet(const c10::FunctionSchema& s) {
  static constexpr std::array<const char*, 4> names{"dtype", "layout", "device", "pin_memory"};
  const auto& args = s.arguments();
  for (size_t i = 0; i + names.size() <= args.size(); ++i) {
    bool hit = true;
    for (size_t j = 0; j < names.size() && hit; ++j) hit = args[i + j].name() == names[j];
    if (hit) return true;
  }
  return false;
}

const OpTraits& traits_of(const OperatorHandle& op) {
  thread_local std::unordered_map<const c10::FunctionSchema*, OpTraits> cache;
  const c10::FunctionSchema* key = &op.schema();
  auto it = cache.find(key);
  if (it != cache.end()) return it->second;
  OpTraits t;
  // An argument called `device` only selects the output device for factory-like operators:
  // those with a BackendSelect kernel or a full TensorOptions quartet.
  if (op.hasKernelForDispatchKey(DispatchKey::BackendSelect) || has_tensor_options_quartet(*key)) {
    const auto& args = key->arguments();
    for (size_t i = 0; i < args.size(); ++i)
      if (args[i].name() == "device") { t.device_arg = static_cast<int>(i); break; }
  }
  if (key->overload_name().empty() && key->arguments().size() == 4) {
    if (key->name() == "aten::normal_") t.inplace_rng = OpTraits::NormalInplace;
    if (key->name() == "aten::uniform_") t.inplace_rng = OpTraits::UniformInplace;
  }
  t.meta_ok = op.hasKernelForDispatchKey(DispatchKey::Meta) ||
              op.hasKernelForDispatchKey(DispatchKey::CompositeExplicitAutograd) ||
              op.hasKernelForDispatchKey(DispatchKey::CompositeExplicitAutogradNonFunctional) ||
              op.hasKernelForDispatchKey(DispatchKey::CompositeImplicitAutograd);
  return cache.emplace(key, t).first->second;
}

}  // namespace

int device_argument_index(const OperatorHandle& op) { return traits_of(op).device_arg; }

namespace {

const DispatchKeySet kBelowFake{DispatchKeySet::FULL_AFTER, DispatchKey::Fake};

void fake_fallback(const OperatorHandle& op, DispatchKeySet ks, Stack* stack) {
  c10::impl::ExcludeDispatchKeyGuard no_reentry{DispatchKey::Fake};
  const auto& schema = op.schema();
  const size_t nargs = schema.arguments().size();

  {
    // In-place RNG on a fake floating-point tensor: validate the arguments like ATen does
    // ($TORCH/include/ATen/native/DistributionTemplates.h normal_impl_ / uniform_impl_) and hand
    // `self` back; everything else (integer tensors, tensor-valued arguments) takes the Meta kernel.
    const OpTraits& tr = traits_of(op);
    if (tr.inplace_rng != OpTraits::None) {
      const IValue& self_iv = torch::jit::peek(*stack, 0, nargs);
      const IValue& a_iv = torch::jit::peek(*stack, 1, nargs);
      const IValue& b_iv = torch::jit::peek(*stack, 2, nargs);
      if (self_iv.isTensor() && is_fake(self_iv.toTensor()) && at::isFloatingType(self_iv.toTensor().scalar_type()) &&
          (a_iv.isDouble() || a_iv.isInt()) && (b_iv.isDouble() || b_iv.isInt())) {
        const double a = a_iv.isDouble() ? a_iv.toDouble() : static_cast<double>(a_iv.toInt());
        const double b = b_iv.isDouble() ? b_iv.toDouble() : static_cast<double>(b_iv.toInt());
        if (tr.inplace_rng == OpTraits::NormalInplace) {
          TORCH_CHECK(b >= 0.0, "normal expects std >= 0.0, but found std ", fmt_double(b));
        } else {
          TORCH_CHECK(a <= b, "uniform_ expects to return a [from, to) range, but found from=", fmt_double(a),
                      " > to=", fmt_double(b));
        }
        at::Tensor self = self_iv.toTensor();
        torch::jit::drop(*stack, nargs);
        torch::jit::push(*stack, std::move(self));
        return;
      }
    }
  }

  bool has_fake = false, has_tensor = false;
  std::optional<Device> tensor_device;
  // meta twin -> fake, for results that are one of the (in-place modified) arguments
  c10::SmallVector<std::pair<const TensorImpl*, c10::intrusive_ptr<FakeTensorImpl>>, 4> twins;

  for_each_tensor_mut(*stack, nargs, [&](at::Tensor& t) {
    has_tensor = true;
    if (!(t.dim() == 0 && t.is_cpu())) {  // 0-dim CPU tensors act as scalars
      if (tensor_device) {
        TORCH_CHECK(*tensor_device == t.device(),
                    "Expected all tensors to be on the same device, but found at least two devices, ",
                    *tensor_device, " and ", t.device(), "!");
      } else {
        tensor_device = t.device();
      }
    }
    if (is_fake(t)) {
      has_fake = true;
      auto self = c10::intrusive_ptr<FakeTensorImpl>::reclaim_copy(fake_impl(t));
      const auto& meta = self->meta();
      twins.emplace_back(meta.get(), std::move(self));
      t = at::Tensor::wrap_tensor_impl(meta);
    }
  });

  const OpTraits& traits = traits_of(op);
  IValue* device_arg =
      traits.device_arg >= 0 ? &torch::jit::peek(*stack, traits.device_arg, nargs) : nullptr;

  if (!(has_fake || device_arg != nullptr || !has_tensor)) {
    // an ordinary op on real tensors: not our business
    op.redispatchBoxed(ks & kBelowFake, stack);
    return;
  }

  const Device out_device = (device_arg && device_arg->isDevice()) ? device_arg->toDevice()
                            : tensor_device                        ? *tensor_device
                                                                   : Device(c10::kCPU);
  if (device_arg) *device_arg = Device(c10::kMeta);
  // Real tensor operands (beyond 0-dim CPU scalars, which Meta kernels accept) are shadowed by
  // meta tensors of the same geometry for shape inference; the recorder keeps the real ones.
  for_each_tensor_mut(*stack, nargs, [&](at::Tensor& t) {
    if (t.defined() && !t.is_meta() && !(t.dim() == 0 && t.is_cpu())) {
      t = at::empty_strided(t.sizes(), t.strides(), t.options().device(c10::kMeta));
    }
  });

  TORCH_CHECK_NOT_IMPLEMENTED(
      traits.meta_ok, "`", schema.name(),
      "` cannot be run with fake tensor(s) because the meta backend has no kernel for it.");
  op.redispatchBoxed(DispatchKeySet(DispatchKey::Meta), stack);

  for_each_tensor_mut(*stack, schema.returns().size(), [&](at::Tensor& t) {
    if (!t.defined() || !t.is_meta()) return;
    const TensorImpl* meta = t.unsafeGetTensorImpl();
    for (auto& tw : twins) {
      if (tw.first == meta) {  // in-place result: same fake tensor, refreshed geometry
        tw.second->sync_from_meta();
        t = at::Tensor::wrap_tensor_impl(tw.second);
        return;
      }
    }
    t = at::Tensor::wrap_tensor_impl(FakeTensorImpl::make(t.getIntrusivePtr(), out_device));
  });
}

}  // namespace
}  // namespace tdx

TORCH_LIBRARY_IMPL(_, Fake, m) {
  m.fallback(torch::CppFunction::makeFromBoxedFunction<&tdx::fake_fallback>());
}

// ---------------------------------------------------------------------------------------------
// fake mode (thread-local, nestable) and fake CUDA
// ---------------------------------------------------------------------------------------------
namespace tdx {
namespace {

thread_local size_t tls_fake_level = 0;
thread_local std::unique_ptr<c10::impl::DeviceGuardImplInterface> tls_noop_cuda_guard;

constexpr auto kCudaSlot = static_cast<size_t>(c10::DeviceType::CUDA);

// Without a CUDA build there is no device guard for CUDA and any `device="cuda"` factory would
// assert; a no-op guard makes PyTorch accept the device for the lifetime of the fake mode.
void install_noop_cuda_guard() {
  if (c10::impl::device_guard_impl_registry[kCudaSlot].load() != nullptr) return;
  tls_noop_cuda_guard = std::make_unique<c10::impl::NoOpDeviceGuardImpl<c10::DeviceType::CUDA>>();
  c10::impl::device_guard_impl_registry[kCudaSlot].store(tls_noop_cuda_guard.get());
}
void remove_noop_cuda_guard() noexcept {
  const auto* cur = c10::impl::device_guard_impl_registry[kCudaSlot].load();
  if (cur == nullptr || cur != tls_noop_cuda_guard.get()) return;
  c10::impl::device_guard_impl_registry[kCudaSlot].store(nullptr);
  tls_noop_cuda_guard.reset();
}

}  // namespace

void enter_fake_mode(bool fake_cuda) {
  if (++tls_fake_level == 1) {
    if (fake_cuda) install_noop_cuda_guard();
    c10::impl::tls_set_dispatch_key_included(DispatchKey::Fake, true);
  }
}

void leave_fake_mode() noexcept {
  if (tls_fake_level == 0) return;
  if (--tls_fake_level == 0) {
    remove_noop_cuda_guard();
    c10::impl::tls_set_dispatch_key_included(DispatchKey::Fake, false);
  }
}

bool fake_mode_active() noexcept { return tls_fake_level > 0; }

}  // namespace tdx
