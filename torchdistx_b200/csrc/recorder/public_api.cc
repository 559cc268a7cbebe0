// The C++ surface of include/torchdistx_b200/{fake,deferred_init}.h: thin exported wrappers over
// the runtime (the rest of this library has hidden visibility).  Names and behaviour follow the
// reference's installed headers (src/cc/torchdistx/fake.h:34-83, deferred_init.h:25-37).
#include "torchdistx_b200/deferred_init.h"
#include "torchdistx_b200/fake.h"

#include <c10/util/Exception.h>

#include "fake_tensor.h"
#include "planner.h"
#include "tape.h"

namespace torchdistx {

void enterFakeMode(bool fake_cuda) { tdx::enter_fake_mode(fake_cuda); }
void leaveFakeMode() noexcept { tdx::leave_fake_mode(); }
bool isFakeModeActive() noexcept { return tdx::fake_mode_active(); }
bool isFake(const at::TensorBase& tensor) noexcept { return tdx::is_fake(tensor); }

FakeTensor::FakeTensor(const at::TensorBase& tensor, bool unsafe) : impl_(tensor.unsafeGetTensorImpl()) {
  TORCH_CHECK_VALUE(unsafe || tdx::is_fake(tensor), "`tensor` was expected to be a fake tensor.");
}

at::Tensor FakeTensor::toMeta() const {
  auto meta = static_cast<tdx::FakeTensorImpl*>(impl_)->meta()->shallow_copy_and_detach(
      /*version_counter=*/0, /*allow_tensor_metadata_change=*/false);
  meta->set_autograd_meta(nullptr);
  return at::Tensor(std::move(meta));
}

const at::Storage& FakeTensor::meta_storage() const noexcept {
  return static_cast<tdx::FakeTensorImpl*>(impl_)->meta()->storage();
}

FakeTensor asFake(const at::TensorBase& tensor) { return FakeTensor{tensor}; }
FakeTensor unsafeAsFake(const at::TensorBase& tensor) noexcept { return FakeTensor{tensor, /*unsafe=*/true}; }

void enterDeferredInit() { tdx::enter_deferred_init(); }
void leaveDeferredInit() noexcept { tdx::leave_deferred_init(); }
bool canMaterialize(const at::Tensor& tensor) noexcept { return tdx::can_materialize(tensor); }

at::Tensor materializeTensor(const at::Tensor& tensor) {
  if (!tdx::can_materialize(tensor)) return tensor;
  return tdx::materialize_one(tensor, tdx::MaterializeOptions{});
}

at::Tensor materializeTensor(const at::Tensor& tensor, std::optional<c10::Device> device, int64_t rank, int64_t world) {
  if (!tdx::can_materialize(tensor)) return tensor;
  tdx::MaterializeOptions o;
  o.device = device;
  if (world > 1) {
    TORCH_CHECK_VALUE(rank >= 0 && rank < world, "shard must be (rank, world_size) with 0 <= rank < world_size");
    o.shard = tdx::ShardSpec{rank, world};
  }
  return tdx::materialize_one(tensor, o);
}

}  // namespace torchdistx
