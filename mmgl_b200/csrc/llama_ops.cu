// Elementwise pieces of the frozen Llama decoder layers around which the gated cross-attention blocks are interleaved
// (BASELINE configs[4]: Llama-2-7B + flamingo).  The reference has no Llama wrapper; the arithmetic
// restated here is HF transformers' (models/llama/modeling_llama.py): apply_rotary_pos_emb / rotate_half (:110-140) and
// LlamaMLP.forward (:155-165) down_proj(silu(gate_proj(x)) * up_proj(x)).  HBM-bound passes over bf16 activations:
// 16-byte vector accesses, grid sized in multiples of the SM count.
#include <cuda_bf16.h>

#include "../../include/mmgl_b200.h"
#include "common.cuh"

namespace mmgl {
namespace {

// x[row, s * H + h * d + i] for section s in [0, sections) (q, k of a fused Q|K|V buffer), i in [0, d / 2):
//   (x1, x2) = (x[i], x[i + d/2])  ->  (x1 cos - x2 sin, x2 cos + x1 sin),  cos / sin = table[pos][i], pos = row % seq.
// inverse = 1 applies the transposed rotation (sin -> -sin): the backward of the forward call.
__global__ void __launch_bounds__(256)
rope_kernel(__nv_bfloat16* __restrict__ x, int64_t ld, int64_t rows, int seq, int heads, int d, int sections,
            const float2* __restrict__ cs, float sign) {
  const int vec_per_head = d / 16;                       // 8-element vectors in half a head
  const int64_t per_row = (int64_t)sections * heads * vec_per_head;
  const int64_t total = rows * per_row;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = t / per_row;
    const int rem = (int)(t % per_row);
    const int sh = rem / vec_per_head, v = rem % vec_per_head;     // (section, head) index and vector inside the half head
    __nv_bfloat16* p1 = x + row * ld + (int64_t)sh * d + v * 8;
    __nv_bfloat16* p2 = p1 + d / 2;
    const float2* c = cs + ((int64_t)(row % seq) * (d / 2) + v * 8);
    uint4 a = *reinterpret_cast<const uint4*>(p1), b = *reinterpret_cast<const uint4*>(p2);
    uint32_t wa[4] = {a.x, a.y, a.z, a.w}, wb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 c0 = __ldg(c + 2 * e), c1 = __ldg(c + 2 * e + 1);
      const float x1l = bf16lo(wa[e]), x1h = bf16hi(wa[e]), x2l = bf16lo(wb[e]), x2h = bf16hi(wb[e]);
      wa[e] = pack_bf16(x1l * c0.x - sign * x2l * c0.y, x1h * c1.x - sign * x2h * c1.y);
      wb[e] = pack_bf16(x2l * c0.x + sign * x1l * c0.y, x2h * c1.x + sign * x1h * c1.y);
    }
    *reinterpret_cast<uint4*>(p1) = make_uint4(wa[0], wa[1], wa[2], wa[3]);
    *reinterpret_cast<uint4*>(p2) = make_uint4(wb[0], wb[1], wb[2], wb[3]);
  }
}

__device__ __forceinline__ float sigmoidf_(float v) { return __frcp_rn(1.f + __expf(-v)); }

// h = silu(g) * u with gu = [g | u] ([M, 2F], the output of one GEMM over the row-concatenated gate / up weights)
__global__ void __launch_bounds__(256)
swiglu_fwd_kernel(const __nv_bfloat16* __restrict__ gu, int64_t ldgu, __nv_bfloat16* __restrict__ h, int64_t ldh, int64_t m, int64_t f) {
  const int64_t per_row = f / 8, total = m * per_row;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = t / per_row, c = (t % per_row) * 8;
    const uint4 g = __ldg(reinterpret_cast<const uint4*>(gu + row * ldgu + c));
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(gu + row * ldgu + f + c));
    const uint32_t wg[4] = {g.x, g.y, g.z, g.w}, wu[4] = {u.x, u.y, u.z, u.w};
    uint32_t o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float gl = bf16lo(wg[e]), gh = bf16hi(wg[e]);
      o[e] = pack_bf16(gl * sigmoidf_(gl) * bf16lo(wu[e]), gh * sigmoidf_(gh) * bf16hi(wu[e]));
    }
    *reinterpret_cast<uint4*>(h + row * ldh + c) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// dg = dh * u * s * (1 + g * (1 - s)),  du = dh * g * s,  s = sigmoid(g);   dgu = [dg | du]
__global__ void __launch_bounds__(256)
swiglu_bwd_kernel(const __nv_bfloat16* __restrict__ gu, int64_t ldgu, const __nv_bfloat16* __restrict__ dh, int64_t lddh,
                  __nv_bfloat16* __restrict__ dgu, int64_t lddgu, int64_t m, int64_t f) {
  const int64_t per_row = f / 8, total = m * per_row;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = t / per_row, c = (t % per_row) * 8;
    const uint4 g = __ldg(reinterpret_cast<const uint4*>(gu + row * ldgu + c));
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(gu + row * ldgu + f + c));
    const uint4 d = __ldg(reinterpret_cast<const uint4*>(dh + row * lddh + c));
    const uint32_t wg[4] = {g.x, g.y, g.z, g.w}, wu[4] = {u.x, u.y, u.z, u.w}, wd[4] = {d.x, d.y, d.z, d.w};
    uint32_t og[4], ou[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float r[2][2];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const float gv = q ? bf16hi(wg[e]) : bf16lo(wg[e]), uv = q ? bf16hi(wu[e]) : bf16lo(wu[e]);
        const float dv = q ? bf16hi(wd[e]) : bf16lo(wd[e]);
        const float s = sigmoidf_(gv);
        r[0][q] = dv * uv * s * (1.f + gv * (1.f - s));
        r[1][q] = dv * gv * s;
      }
      og[e] = pack_bf16(r[0][0], r[0][1]);
      ou[e] = pack_bf16(r[1][0], r[1][1]);
    }
    *reinterpret_cast<uint4*>(dgu + row * lddgu + c) = make_uint4(og[0], og[1], og[2], og[3]);
    *reinterpret_cast<uint4*>(dgu + row * lddgu + f + c) = make_uint4(ou[0], ou[1], ou[2], ou[3]);
  }
}

inline unsigned grid_for(int64_t total) {
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 8;
  return (unsigned)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

}  // namespace
}  // namespace mmgl

using namespace mmgl;

extern "C" int mmgl_rope_inplace(void* x, int64_t ld, int64_t rows, int64_t seq, int64_t heads, int64_t head_dim,
                                 int64_t sections, const float* cos_sin, int32_t inverse, void* stream_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  MMGL_REQUIRE(x && cos_sin && rows > 0 && seq > 0 && heads > 0 && sections > 0, "mmgl_rope_inplace: bad arguments");
  MMGL_BIND(x, "mmgl_rope_inplace");
  MMGL_REQUIRE(head_dim % 16 == 0 && ld % 8 == 0 && aligned16(x) && (reinterpret_cast<uintptr_t>(cos_sin) & 7) == 0,
               "mmgl_rope_inplace: head_dim must be a multiple of 16, x 16-byte aligned with ld %% 8 == 0");
  MMGL_REQUIRE(sections * heads * head_dim <= ld, "mmgl_rope_inplace: sections * heads * head_dim exceeds the row pitch");
  const int64_t total = rows * sections * heads * (head_dim / 16);
  rope_kernel<<<grid_for(total), 256, 0, s>>>((__nv_bfloat16*)x, ld, rows, (int)seq, (int)heads, (int)head_dim, (int)sections,
                                              reinterpret_cast<const float2*>(cos_sin), inverse ? -1.f : 1.f);
  return check_launch("mmgl_rope_inplace");
}

extern "C" int mmgl_swiglu_fwd(const void* gu, int64_t ldgu, void* h, int64_t ldh, int64_t m, int64_t f, void* stream_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  MMGL_REQUIRE(gu && h && m > 0 && f > 0, "mmgl_swiglu_fwd: bad arguments");
  MMGL_BIND(gu, "mmgl_swiglu_fwd");
  MMGL_REQUIRE(f % 8 == 0 && ldgu % 8 == 0 && ldh % 8 == 0 && aligned16(gu) && aligned16(h),
               "mmgl_swiglu_fwd: F and the leading dimensions must be multiples of 8, pointers 16-byte aligned");
  swiglu_fwd_kernel<<<grid_for(m * (f / 8)), 256, 0, s>>>((const __nv_bfloat16*)gu, ldgu, (__nv_bfloat16*)h, ldh, m, f);
  return check_launch("mmgl_swiglu_fwd");
}

extern "C" int mmgl_swiglu_bwd(const void* gu, int64_t ldgu, const void* dh, int64_t lddh, void* dgu, int64_t lddgu, int64_t m,
                               int64_t f, void* stream_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  MMGL_REQUIRE(gu && dh && dgu && m > 0 && f > 0, "mmgl_swiglu_bwd: bad arguments");
  MMGL_BIND(gu, "mmgl_swiglu_bwd");
  MMGL_REQUIRE(f % 8 == 0 && ldgu % 8 == 0 && lddh % 8 == 0 && lddgu % 8 == 0 && aligned16(gu) && aligned16(dh) && aligned16(dgu),
               "mmgl_swiglu_bwd: F and the leading dimensions must be multiples of 8, pointers 16-byte aligned");
  swiglu_bwd_kernel<<<grid_for(m * (f / 8)), 256, 0, s>>>((const __nv_bfloat16*)gu, ldgu, (const __nv_bfloat16*)dh, lddh,
                                                          (__nv_bfloat16*)dgu, lddgu, m, f);
  return check_launch("mmgl_swiglu_bwd");
}
