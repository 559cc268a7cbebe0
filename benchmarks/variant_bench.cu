// Kernel-variant A/B harness: dlopen()s several builds of libtdx_init.so (different -D knobs),
// runs the same descriptor through each over the C ABI (include/tdx_init.h), times the launch
// with CUDA events and prints a checksum of the output so that variants can be compared bit for
// bit.  No torch, no Python: start-up is a fraction of a second, GPU minutes go to the kernels.
//
//   nvcc -O2 -Iinclude benchmarks/variant_bench.cu -o benchmarks/variant_bench -ldl
//   benchmarks/variant_bench [--gib 4] [--dtype bf16|f16|f32] [--src normal|uniform] [--algo N]
//                            [--iters 9] lib1.so lib2.so ...
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "tdx_init.h"

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
      exit(2);                                                                     \
    }                                                                              \
  } while (0)

__global__ void checksum_kernel(const uint32_t* p, size_t n, unsigned long long* out) {
  unsigned long long acc = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    acc += (unsigned long long)p[i] * (unsigned long long)((i * 2654435761ull) | 1ull);
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}

struct Lib {
  void* h;
  size_t (*ws_bytes)(int);
  int (*upload)(const TdxInitDesc*, int, void*, size_t, void*, TdxPlan*);
  int (*launch)(const TdxPlan*, void*, void*);
  const char* (*last_error)();
};

int main(int argc, char** argv) {
  double gib = 4.0;
  int iters = 9, algo = 0;
  std::string dtype = "bf16", src = "normal";
  std::vector<std::string> libs;
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    if (a == "--gib") gib = atof(argv[++i]);
    else if (a == "--iters") iters = atoi(argv[++i]);
    else if (a == "--dtype") dtype = argv[++i];
    else if (a == "--src") src = argv[++i];
    else if (a == "--algo") algo = atoi(argv[++i]);
    else libs.push_back(a);
  }
  const int isz = dtype == "f32" ? 4 : 2;
  const size_t bytes = (size_t)(gib * (1ull << 30));
  const size_t n = bytes / isz;
  void* buf;
  CK(cudaMalloc(&buf, bytes));
  unsigned long long* dsum;
  CK(cudaMalloc(&dsum, 8));
  cudaStream_t st;
  CK(cudaStreamCreate(&st));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  for (const std::string& path : libs) {
    Lib L;
    L.h = dlopen(path.c_str(), RTLD_NOW | RTLD_LOCAL);
    if (!L.h) { fprintf(stderr, "dlopen %s: %s\n", path.c_str(), dlerror()); return 2; }
    L.ws_bytes = (size_t(*)(int))dlsym(L.h, "tdx_init_workspace_bytes");
    L.upload = (int (*)(const TdxInitDesc*, int, void*, size_t, void*, TdxPlan*))dlsym(L.h, "tdx_plan_upload");
    L.launch = (int (*)(const TdxPlan*, void*, void*))dlsym(L.h, "tdx_plan_launch");
    L.last_error = (const char* (*)())dlsym(L.h, "tdx_last_error");
    TdxInitDesc d;
    memset(&d, 0, sizeof(d));
    d.dst = buf;
    d.elem_begin = 0;
    d.elem_count = n;
    d.philox_seed = 1234;
    d.philox_offset = 8;
    d.p0 = src == "normal" ? 0.0 : -0.05;
    d.p1 = src == "normal" ? 0.02 : 0.05;
    d.dtype = dtype == "f32" ? TDX_F32 : dtype == "f16" ? TDX_F16 : TDX_BF16;
    d.src = src == "normal" ? TDX_SRC_NORMAL : TDX_SRC_UNIFORM;
    d.algo = (uint8_t)algo;
    const size_t wsb = L.ws_bytes(1);
    void* ws;
    CK(cudaMalloc(&ws, wsb));
    TdxPlan plan;
    if (L.upload(&d, 1, ws, wsb, st, &plan)) { fprintf(stderr, "upload: %s\n", L.last_error()); return 2; }
    CK(cudaMemsetAsync(buf, 0, bytes, st));
    for (int i = 0; i < 3; ++i)
      if (L.launch(&plan, ws, st)) { fprintf(stderr, "launch: %s\n", L.last_error()); return 2; }
    CK(cudaStreamSynchronize(st));
    std::vector<float> ms;
    for (int i = 0; i < iters; ++i) {
      CK(cudaEventRecord(e0, st));
      L.launch(&plan, ws, st);
      CK(cudaEventRecord(e1, st));
      CK(cudaEventSynchronize(e1));
      float t;
      CK(cudaEventElapsedTime(&t, e0, e1));
      ms.push_back(t);
    }
    std::sort(ms.begin(), ms.end());
    CK(cudaMemsetAsync(dsum, 0, 8, st));
    checksum_kernel<<<148 * 8, 256, 0, st>>>((const uint32_t*)buf, bytes / 4, dsum);
    unsigned long long sum = 0;
    CK(cudaMemcpyAsync(&sum, dsum, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    const double med = ms[ms.size() / 2];
    printf("{\"lib\": \"%s\", \"dtype\": \"%s\", \"src\": \"%s\", \"algo\": %d, \"bytes\": %zu, \"ms_med\": %.4f, "
           "\"ms_best\": %.4f, \"gbs\": %.1f, \"frac\": %.4f, \"checksum\": \"%016llx\"}\n",
           path.c_str(), dtype.c_str(), src.c_str(), algo, bytes, med, ms[0], bytes / med / 1e6,
           bytes / med / 1e6 / 6565.8, sum);
    fflush(stdout);
    CK(cudaFree(ws));
    // the library stays loaded: unloading a CUDA module at exit ordering is not worth testing here
  }
  return 0;
}
