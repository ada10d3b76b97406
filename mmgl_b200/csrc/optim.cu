// AdamW on fp32 master weights with the bf16 shadow written in the same pass (round 2).
//
// Replaces, in the reference's train loop (language_modelling/run_generation.py:329-333 builds torch.optim.AdamW, :486
// calls optimizer.step()), torch's fused multi-tensor AdamW (1.45 ms per cfg2 step for 216.7 M parameters) plus the
// per-parameter fp32 -> bf16 conversion the kernels' operands need afterwards.  One read of (p, g, m, v), one write of
// (p, m, v, bf16 p): 30 bytes per parameter, HBM-bound.
//
// Arithmetic = torch.optim.AdamW (amsgrad = False, maximize = False), in fp32:
//   p *= 1 - lr * weight_decay;  m = b1 m + (1 - b1) g;  v = b2 v + (1 - b2) g^2
//   p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../../include/mmgl_b200.h"
#include "common.cuh"

namespace mmgl {
namespace {

struct AdamWConst {
  float decay;        // 1 - lr * weight_decay
  float b1, b2;       // betas
  float step_size;    // lr / (1 - b1^t)
  float inv_bc2_sqrt; // 1 / sqrt(1 - b2^t)
  float eps;
  float grad_scale;   // gradients are multiplied by this first (1 = none; e.g. 1 / accumulation steps)
};

__device__ __forceinline__ float adamw_one(float& p, float g, float& m, float& v, const AdamWConst& c) {
  g *= c.grad_scale;
  p *= c.decay;
  m = fmaf(c.b1, m, (1.f - c.b1) * g);
  v = fmaf(c.b2, v, (1.f - c.b2) * g * g);
  const float denom = sqrtf(v) * c.inv_bc2_sqrt + c.eps;
  p -= c.step_size * (m / denom);
  return p;
}

__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
             __nv_bfloat16* __restrict__ shadow, int64_t n, AdamWConst c, int vec) {
  pdl_launch();
  pdl_wait();
  const int64_t stride = (int64_t)gridDim.x * 256;
  if (vec) {
    const int64_t n4 = n >> 2;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += stride) {
      float4 pp = reinterpret_cast<float4*>(p)[i];
      const float4 gg = __ldg(reinterpret_cast<const float4*>(g) + i);
      float4 mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
      adamw_one(pp.x, gg.x, mm.x, vv.x, c); adamw_one(pp.y, gg.y, mm.y, vv.y, c);
      adamw_one(pp.z, gg.z, mm.z, vv.z, c); adamw_one(pp.w, gg.w, mm.w, vv.w, c);
      reinterpret_cast<float4*>(p)[i] = pp;
      reinterpret_cast<float4*>(m)[i] = mm;
      reinterpret_cast<float4*>(v)[i] = vv;
      if (shadow != nullptr) reinterpret_cast<uint2*>(shadow)[i] = make_uint2(pack_bf16(pp.x, pp.y), pack_bf16(pp.z, pp.w));
    }
    for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += stride) {   // tail (< 4 elements)
      float pp = p[i], mm = m[i], vv = v[i];
      adamw_one(pp, g[i], mm, vv, c);
      p[i] = pp; m[i] = mm; v[i] = vv;
      if (shadow != nullptr) shadow[i] = __float2bfloat16_rn(pp);
    }
  } else {
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += stride) {
      float pp = p[i], mm = m[i], vv = v[i];
      adamw_one(pp, g[i], mm, vv, c);
      p[i] = pp; m[i] = mm; v[i] = vv;
      if (shadow != nullptr) shadow[i] = __float2bfloat16_rn(pp);
    }
  }
}

}  // namespace
}  // namespace mmgl

using namespace mmgl;

extern "C" int mmgl_adamw_step(void* param, const void* grad, void* exp_avg, void* exp_avg_sq, void* shadow_bf16, int64_t n,
                               float lr, float beta1, float beta2, float eps, float weight_decay, int64_t step, float grad_scale,
                               void* stream_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  MMGL_REQUIRE(param && grad && exp_avg && exp_avg_sq && n > 0, "mmgl_adamw_step: bad arguments");
  MMGL_REQUIRE(step >= 1, "mmgl_adamw_step: step counts from 1 (the value AFTER this update, as torch.optim does)");
  MMGL_BIND(param, "mmgl_adamw_step");
  AdamWConst c;
  c.decay = 1.f - lr * weight_decay;
  c.b1 = beta1; c.b2 = beta2;
  const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
  c.step_size = (float)((double)lr / bc1);
  c.inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
  c.eps = eps;
  c.grad_scale = grad_scale;
  const int vec = aligned16(param) && aligned16(grad) && aligned16(exp_avg) && aligned16(exp_avg_sq) &&
                  (shadow_bf16 == nullptr || (reinterpret_cast<uintptr_t>(shadow_bf16) & 7) == 0);
  int64_t blocks = ((vec ? (n + 3) / 4 : n) + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  MMGL_CUDA(launch_pdl(adamw_kernel, dim3((unsigned)blocks), dim3(256), 0, s, (float*)param, (const float*)grad, (float*)exp_avg,
                       (float*)exp_avg_sq, (__nv_bfloat16*)shadow_bf16, n, c, vec));
  return check_launch("mmgl_adamw_step");
}
