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
// Make the device that owns `device_ptr` current for the calling thread (this library links its own CUDA runtime;
// PyTorch's autograd threads may not have bound a context yet, and driver calls such as cuTensorMapEncodeTiled
// need one).  Returns 0 or an error code with the message set.
int bind_device_of(const void* device_ptr, const char* who);

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

#define MMGL_BIND(ptr, who)                                   \
  do {                                                        \
    if (int rc__ = ::mmgl::bind_device_of((ptr), (who))) return rc__; \
  } while (0)

// ---- programmatic dependent launch (PDL) ----------------------------------------------------
// Kernels of this library are launched with cudaLaunchAttributeProgrammaticStreamSerialization: the next kernel on the
// stream may begin (block scheduling, barrier init, TMEM allocation, descriptor prefetch) while the tail of the previous
// one drains, instead of paying the ~2-3 us launch gap after every one of the ~800 kernels of a step.  Contract inside
// every kernel launched this way: pdl_launch() first (dependents may start their prologue), then pdl_wait() -- executed
// by EVERY thread -- before the first access to global memory: it returns once the preceding grid has completed and its
// writes are visible, so read-after-write AND write-after-read on recycled buffers stay ordered.  Both are no-ops under a
// normal launch.  MMGL_PDL=0 switches the attribute off (A/B control).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

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

// tanh of the scalar Flamingo gate (model/modelling_cross_attention.py:335, :359).  --use_fast_math turns tanhf into
// tanh.approx (2^-11 relative error); the gates are trained from 0.0 and their gradient carries 1 - tanh^2, so this stays
// at fp32 accuracy whatever the translation unit's flags: odd Taylor polynomial below 0.1 (error < 3e-11), otherwise
// (e^2x - 1) / (e^2x + 1) with an IEEE reciprocal (__expf: 2 ulp of a value >= 1.22).
__device__ __forceinline__ float tanh_precise(float x) {
  const float ax = fabsf(x);
  float t;
  if (ax < 0.1f) {
    const float x2 = ax * ax;
    t = ax * (1.f + x2 * (-0.333333333f + x2 * (0.133333333f + x2 * -0.0539682540f)));
  } else if (ax > 15.f) {
    t = 1.f;
  } else {
    const float e = __expf(2.f * ax);
    t = __fmul_rn(e - 1.f, __frcp_rn(e + 1.f));
  }
  return copysignf(t, x);
}

// ---- counter-based dropout mask ---------------------------------------------------------------
// 16 random bits per element from splitmix64 of (seed, row, col / 8); element (row, col) is KEPT iff its 16-bit lane
// >= thresh, thresh = round(p * 65536).  Stateless, so backward regenerates the forward mask from the seed alone
// (nn.functional.dropout call sites: model/modelling_cross_attention.py:332, :356).
__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t z) {
  z += 0x9e3779b97f4a7c15ULL;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
  return z ^ (z >> 31);
}
struct DropBits { uint64_t lo, hi; };
__host__ __device__ __forceinline__ DropBits dropout_bits(uint64_t seed, int64_t row, int64_t col8, int64_t groups_per_row) {
  const uint64_t g = static_cast<uint64_t>(row * groups_per_row + col8);
  DropBits b;
  b.lo = splitmix64(seed ^ (2 * g));
  b.hi = splitmix64(seed ^ (2 * g + 1));
  return b;
}
__host__ __device__ __forceinline__ bool dropout_keep(const DropBits& b, int e, uint32_t thresh) {
  const uint64_t w = (e < 4) ? b.lo : b.hi;
  return static_cast<uint32_t>((w >> (16 * (e & 3))) & 0xFFFFu) >= thresh;
}

}  // namespace mmgl
