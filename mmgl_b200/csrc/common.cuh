// Host-side helpers shared by all translation units of libmmgl_b200.so: error string, launch counter.
#pragma once
#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace mmgl {

void set_error(const char* fmt, ...);          // thread-local message, returned by mmgl_last_error_string()
extern std::atomic<int64_t> g_launch_count;    // every kernel launch of this library bumps it
int sm_count();                                // SM count of the current device (cached per device)

inline int check_launch(const char* what) {
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    return 1;
  }
  return 0;
}

#define MMGL_REQUIRE(cond, ...)      \
  do {                               \
    if (!(cond)) {                   \
      ::mmgl::set_error(__VA_ARGS__); \
      return 2;                      \
    }                                \
  } while (0)

#define MMGL_CUDA(call)                                                        \
  do {                                                                         \
    cudaError_t e__ = (call);                                                  \
    if (e__ != cudaSuccess) {                                                  \
      ::mmgl::set_error("%s failed: %s", #call, cudaGetErrorString(e__));      \
      return 3;                                                                \
    }                                                                          \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// ---- small device helpers -------------------------------------------------------------------
__device__ __forceinline__ float bf16lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16hi(uint32_t u) { return __uint_as_float(u & 0xFFFF0000u); }
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace mmgl
