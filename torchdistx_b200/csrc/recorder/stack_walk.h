// Visiting the tensors of the last `n` values of a JIT stack, including tensors inside
// Tensor[] lists (same coverage as the reference's processTensors/convertTensors,
// reference src/cc/torchdistx/stack_utils.cc:23-58; header-only templates here so the
// per-op visitors inline into the dispatch handlers).
#pragma once

#include <ATen/Tensor.h>
#include <ATen/core/ivalue.h>
#include <ATen/core/stack.h>

namespace tdx {

// f(const at::Tensor&) -> bool (true stops the walk) or -> void
template <class F>
inline void for_each_tensor(const torch::jit::Stack& s, size_t n, F&& f) {
  auto call = [&](const at::Tensor& t) -> bool {
    if constexpr (std::is_void_v<decltype(f(t))>) {
      f(t);
      return false;
    } else {
      return f(t);
    }
  };
  for (size_t i = 0; i < n; ++i) {
    const c10::IValue& v = torch::jit::peek(s, i, n);
    if (v.isTensor()) {
      if (call(v.toTensor())) return;
    } else if (v.isList()) {
      for (const c10::IValue& e : v.toListRef())
        if (e.isTensor() && call(e.toTensor())) return;
    }
  }
}

// f(at::Tensor&): may replace the tensor in place
template <class F>
inline void for_each_tensor_mut(torch::jit::Stack& s, size_t n, F&& f) {
  for (size_t i = 0; i < n; ++i) {
    c10::IValue& v = torch::jit::peek(s, i, n);
    if (v.isTensor()) {
      f(v.toTensor());
    } else if (v.isList()) {
      for (const c10::IValue& e : v.toListRef())
        if (e.isTensor()) f(const_cast<c10::IValue&>(e).toTensor());
    }
  }
}

}  // namespace tdx
