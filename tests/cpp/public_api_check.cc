// Exercises include/torchdistx_b200/{fake,deferred_init}.h the way a C++ caller of the reference's
// installed headers (src/cc/torchdistx/fake.h, deferred_init.h) would.  TEST INFRASTRUCTURE: built by
// tests/test_cpp_surface.py into a small shared object, loaded into the test process with ctypes.
// The scenarios restate the reference's Python tests (tests/python/test_fake.py,
// test_deferred_init.py) at the C++ level, where the reference has no tests of its own.
#include <ATen/ATen.h>
#include <c10/util/Exception.h>

#include <cstdio>
#include <string>

#include "torchdistx_b200/deferred_init.h"
#include "torchdistx_b200/fake.h"

namespace tdx = torchdistx;

static std::string g_error;

#define EXPECT(cond)                                                                  \
  do {                                                                                \
    if (!(cond)) {                                                                    \
      g_error = std::string("line ") + std::to_string(__LINE__) + ": " #cond;         \
      return __LINE__;                                                                \
    }                                                                                 \
  } while (0)

static int run() {
  // ---- fake mode: per thread, nests --------------------------------------------------------
  EXPECT(!tdx::isFakeModeActive());
  tdx::enterFakeMode();
  tdx::enterFakeMode();
  at::Tensor f = at::ones({3, 5});
  tdx::leaveFakeMode();
  EXPECT(tdx::isFakeModeActive());
  at::Tensor f2 = f.t();
  tdx::leaveFakeMode();
  EXPECT(!tdx::isFakeModeActive());
  EXPECT(tdx::isFake(f) && tdx::isFake(f2) && !tdx::canMaterialize(f));
  EXPECT(!f.has_storage() || f.storage().nbytes() == 0);  // no memory behind a fake tensor
  at::Tensor real = at::ones({3, 5});
  EXPECT(!tdx::isFake(real));
  // meta twin: same geometry, meta device, detached
  at::Tensor meta = tdx::asFake(f2).toMeta();
  EXPECT(meta.is_meta() && meta.sizes() == f2.sizes() && meta.strides() == f2.strides() && meta.dtype() == f2.dtype());
  EXPECT(!tdx::isFake(meta));
  // views of one fake tensor share their (virtual) storage
  EXPECT(tdx::asFake(f).meta_storage().is_alias_of(tdx::asFake(f2).meta_storage()));
  bool raised = false;
  try {
    tdx::asFake(real);
  } catch (const c10::ValueError&) {
    raised = true;
  }
  EXPECT(raised);
  // a fake tensor that was not recorded cannot be materialised: it comes back as it is
  EXPECT(tdx::materializeTensor(f).unsafeGetTensorImpl() == f.unsafeGetTensorImpl());

  // ---- deferred init: record, then materialise ---------------------------------------------
  at::Tensor w, b, skipped;
  tdx::enterDeferredInit();
  w = at::empty({128, 64});
  w.uniform_(-0.1, 0.1);  // dead
  w.normal_(0.0, 0.02);
  b = at::zeros({64}).add_(1.5);
  {
    tdx::NoDeferredInit off;  // the recorder is what makes new tensors fake: without it, a real tensor
    skipped = at::ones({4});
  }
  tdx::leaveDeferredInit();
  EXPECT(!tdx::isFakeModeActive());
  EXPECT(tdx::isFake(w) && tdx::canMaterialize(w) && tdx::canMaterialize(b));
  EXPECT(!tdx::isFake(skipped) && !tdx::canMaterialize(skipped) && at::equal(skipped, at::ones({4})));

  at::manual_seed(7);
  at::Tensor rw = tdx::materializeTensor(w);
  at::Tensor rb = tdx::materializeTensor(b);
  EXPECT(!tdx::isFake(rw) && rw.sizes() == w.sizes() && rw.device().is_cpu());
  // CPU tensors: the recorded ops replayed in order, dead ones included -> the eager stream (T0')
  at::manual_seed(7);
  at::Tensor ew = at::empty({128, 64});
  ew.uniform_(-0.1, 0.1);
  ew.normal_(0.0, 0.02);
  EXPECT(at::equal(rw, ew));
  EXPECT(at::equal(rb, at::full({64}, 1.5)));
  // identity: materialising again returns the same tensor; real tensors pass through
  EXPECT(tdx::materializeTensor(w).unsafeGetTensorImpl() == rw.unsafeGetTensorImpl());
  EXPECT(tdx::materializeTensor(rw).unsafeGetTensorImpl() == rw.unsafeGetTensorImpl());

  // ---- beyond the reference: this rank's dim-0 chunk ---------------------------------------
  at::Tensor s;
  tdx::enterDeferredInit();
  s = at::empty({10, 4}).normal_();
  tdx::leaveDeferredInit();
  at::manual_seed(11);
  at::Tensor part = tdx::materializeTensor(s, std::nullopt, /*rank=*/2, /*world=*/3);  // rows 8..9 of ceil(10/3)=4 per rank
  at::manual_seed(11);
  at::Tensor whole = at::empty({10, 4}).normal_();
  EXPECT(part.size(0) == 2 && at::equal(part, whole.narrow(0, 8, 2)));
  raised = false;
  try {
    at::Tensor s2;
    tdx::enterDeferredInit();
    s2 = at::empty({4});
    tdx::leaveDeferredInit();
    tdx::materializeTensor(s2, std::nullopt, 3, 3);
  } catch (const c10::ValueError&) {
    raised = true;
  }
  EXPECT(raised);
  return 0;
}

extern "C" __attribute__((visibility("default"))) int tdx_public_api_check() {
  try {
    return run();
  } catch (const std::exception& e) {
    g_error = std::string("exception: ") + e.what();
    return -1;
  }
}

extern "C" __attribute__((visibility("default"))) const char* tdx_public_api_check_error() { return g_error.c_str(); }
