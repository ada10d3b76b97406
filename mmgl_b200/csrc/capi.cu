// Library-level C ABI: version, thread-local error string, launch counter, device info.
#include <cstdlib>
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "../../include/mmgl_b200.h"
#include "common.cuh"

namespace mmgl {

static thread_local char g_err[512] = "";
std::atomic<int64_t> g_launch_count{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("MMGL_PDL"); return e == nullptr || atoi(e) != 0; }();
  return on;
}

int bind_device_of(const void* device_ptr, const char* who) {
  if (device_ptr == nullptr) { set_error("%s: null pointer", who); return 2; }
  cudaPointerAttributes at;
  cudaError_t e = cudaPointerGetAttributes(&at, device_ptr);
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("%s: cudaPointerGetAttributes failed: %s (no CUDA device? this library has no CPU path)", who,
              cudaGetErrorString(e));
    return 3;
  }
  if (at.type != cudaMemoryTypeDevice && at.type != cudaMemoryTypeManaged) {
    set_error("%s: expected a CUDA device pointer (this library has no CPU path)", who);
    return 2;
  }
  e = cudaSetDevice(at.device);
  if (e != cudaSuccess) { set_error("%s: cudaSetDevice(%d) failed: %s", who, at.device, cudaGetErrorString(e)); return 3; }
  return 0;
}

}  // namespace mmgl

extern "C" int mmgl_version(void) { return MMGL_ABI_VERSION; }
extern "C" const char* mmgl_last_error_string(void) { return mmgl::g_err; }
extern "C" int64_t mmgl_launch_count(void) { return mmgl::g_launch_count.load(std::memory_order_relaxed); }
