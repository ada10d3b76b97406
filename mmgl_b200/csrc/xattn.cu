// Fused cross-attention core (forward + backward) over the packed neighbor bank.
//
//   O = softmax(max(Q K^T + mask, -FLT_MAX)) V       per (sample, head); Q pre-scaled by d^-1/2
//
// The FORWARD kernel lives in xattn_sm100.cu (tcgen05 + TMEM + TMA).  This file holds the C entry points and the
// BACKWARD kernel: one CTA per (sample, head, 64-column slice) recomputes P from the saved row statistics, keeps its
// dK / dV slice in registers across all query blocks and writes dQ per block; warp-level mma.sync tiles fed by
// ldmatrix (the five contractions of the backward are tiny, the kernel is HBM / latency bound at these shapes).
// Replaces the autograd backward of model/modelling_cross_attention.py:206-271.
#include <cfloat>
#include <cuda_bf16.h>

#include "../../include/mmgl_b200.h"
#include "common.cuh"

namespace mmgl {

constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* smem_ptr) {
  const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(smem_ptr));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const void* smem_ptr) {
  const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(smem_ptr));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
// D(16x8,f32) += A(16x16,bf16,row) * B(16x8,bf16,col)
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// Cooperative load of `rows` x D bf16 (global row pitch ld) into smem with pitch D+8; rows >= valid -> zeros.
template <int D, int NT>
__device__ __forceinline__ void load_tile(__nv_bfloat16* s, const __nv_bfloat16* g, int64_t ld, int rows, int valid) {
  constexpr int CH = D / 8;
  for (int c = threadIdx.x; c < rows * CH; c += NT) {
    const int r = c / CH, ch = c % CH;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (r < valid) v = __ldg(reinterpret_cast<const uint4*>(g + (int64_t)r * ld + ch * 8));
    *reinterpret_cast<uint4*>(s + r * (D + 8) + ch * 8) = v;
  }
}

// Scores for this warp's 16 query rows against all keys: s[nt][4], nt over 8-key tiles (NKT*2 of them).
template <int D, int NKT>
__device__ __forceinline__ void warp_scores(float (&s)[NKT * 2][4], const __nv_bfloat16* sQw, const __nv_bfloat16* sK,
                                            int nkp, int lane) {
#pragma unroll
  for (int nt = 0; nt < NKT * 2; ++nt) { s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f; }
#pragma unroll
  for (int kk = 0; kk < D / 16; ++kk) {
    uint32_t a[4];
    ldsm_x4(a, sQw + (lane % 16) * (D + 8) + kk * 16 + (lane / 16) * 8);
#pragma unroll
    for (int jt = 0; jt < NKT; ++jt) {
      if (jt * 16 < nkp) {
        uint32_t b[4];
        ldsm_x4(b, sK + (jt * 16 + (lane / 16) * 8 + (lane % 8)) * (D + 8) + kk * 16 + ((lane / 8) % 2) * 8);
        mma16816(s[2 * jt], a, b[0], b[1]);
        mma16816(s[2 * jt + 1], a, b[2], b[3]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------ backward
// grid (D/64, heads, batch), W warps.  Each CTA owns a 64-wide column slice of dQ / dK / dV for one
// (sample, head), loops over query blocks of R = 16*W rows and keeps its dK/dV slice in registers.
template <int D, int NKT, int W>
__global__ void __launch_bounds__(W * 32)
xattn_bwd_kernel(const __nv_bfloat16* __restrict__ d_o, int64_t lddo, const __nv_bfloat16* __restrict__ q, int64_t ldq,
                 const __nv_bfloat16* __restrict__ k, int64_t ldk, const __nv_bfloat16* __restrict__ v, int64_t ldv,
                 const __nv_bfloat16* __restrict__ o, int64_t ldo, const float* __restrict__ stats,
                 const uint8_t* __restrict__ mask, __nv_bfloat16* __restrict__ dq, int64_t lddq,
                 __nv_bfloat16* __restrict__ dk, int64_t lddk, __nv_bfloat16* __restrict__ dv, int64_t lddv,
                 int seq, int nk, int heads) {
  constexpr int NT = W * 32;
  constexpr int R = W * 16;
  constexpr int JT_PER_WARP = (NKT + W - 1) / W;
  constexpr int DP = D + 8;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int nkp = (nk + 15) & ~15;
  const int PP = NKT * 16 + 8;  // pitch of sP / sdS
  __nv_bfloat16* sK = reinterpret_cast<__nv_bfloat16*>(smem_raw);
  __nv_bfloat16* sV = sK + nkp * DP;
  __nv_bfloat16* sQ = sV + nkp * DP;
  __nv_bfloat16* sdO = sQ + R * DP;
  __nv_bfloat16* sP = sdO + R * DP;
  __nv_bfloat16* sdS = sP + R * PP;
  float* sMask = reinterpret_cast<float*>(sdS + R * PP);
  float* sDelta = sMask + NKT * 16;

  const int b = blockIdx.z, h = blockIdx.y, c0 = blockIdx.x * 64;  // column slice [c0, c0+64) of the head
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;

  load_tile<D, NT>(sK, k + (int64_t)b * nk * ldk + h * D, ldk, nkp, nk);
  load_tile<D, NT>(sV, v + (int64_t)b * nk * ldv + h * D, ldv, nkp, nk);
  for (int j = threadIdx.x; j < NKT * 16; j += NT)
    sMask[j] = (j < nk) ? (mask[(int64_t)b * nk + j] ? 0.f : -FLT_MAX) : -INFINITY;

  float acc_dk[JT_PER_WARP][8][4], acc_dv[JT_PER_WARP][8][4];
#pragma unroll
  for (int i = 0; i < JT_PER_WARP; ++i)
#pragma unroll
    for (int dt = 0; dt < 8; ++dt)
#pragma unroll
      for (int e = 0; e < 4; ++e) { acc_dk[i][dt][e] = 0.f; acc_dv[i][dt][e] = 0.f; }

  for (int r0 = 0; r0 < seq; r0 += R) {
    const int rows_valid = min(R, seq - r0);
    __syncthreads();  // previous block's readers of sQ/sdO/sP/sdS are done (also orders the K/V/mask fill)
    load_tile<D, NT>(sQ, q + ((int64_t)b * seq + r0) * ldq + h * D, ldq, R, rows_valid);
    // dO tile + delta_i = sum_d dO[i,d] * O[i,d]
    {
      constexpr int CH = D / 8;
      const __nv_bfloat16* gdo = d_o + ((int64_t)b * seq + r0) * lddo + h * D;
      const __nv_bfloat16* go = o + ((int64_t)b * seq + r0) * ldo + h * D;
      for (int c = threadIdx.x; c < R * CH; c += NT) {
        const int r = c / CH, ch = c % CH;
        uint4 vd = make_uint4(0, 0, 0, 0), vo = make_uint4(0, 0, 0, 0);
        if (r < rows_valid) {
          vd = __ldg(reinterpret_cast<const uint4*>(gdo + (int64_t)r * lddo + ch * 8));
          vo = __ldg(reinterpret_cast<const uint4*>(go + (int64_t)r * ldo + ch * 8));
        }
        *reinterpret_cast<uint4*>(sdO + r * DP + ch * 8) = vd;
        const uint32_t wd[4] = {vd.x, vd.y, vd.z, vd.w}, wo[4] = {vo.x, vo.y, vo.z, vo.w};
        float part = 0.f;
#pragma unroll
        for (int e = 0; e < 4; ++e) part += bf16lo(wd[e]) * bf16lo(wo[e]) + bf16hi(wd[e]) * bf16hi(wo[e]);
#pragma unroll
        for (int off = CH / 2; off > 0; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
        if (ch == 0) sDelta[r] = part;
      }
    }
    __syncthreads();

    // ---- phase 1: this warp's 16 query rows: P, dS -> smem ; dQ slice -> global
    {
      float s[NKT * 2][4];
      warp_scores<D, NKT>(s, sQ + warp * 16 * DP, sK, nkp, lane);
      const int ra = warp * 16 + g, rb = ra + 8;
      float m_a = 0.f, il_a = 0.f, m_b = 0.f, il_b = 0.f;
      if (r0 + ra < seq) { const float* st = stats + (((int64_t)b * heads + h) * seq + r0 + ra) * 2; m_a = st[0]; il_a = st[1]; }
      if (r0 + rb < seq) { const float* st = stats + (((int64_t)b * heads + h) * seq + r0 + rb) * 2; m_b = st[0]; il_b = st[1]; }
      const float del_a = sDelta[ra], del_b = sDelta[rb];
      // dO fragments of this warp's rows (A operand of dP = dO V^T)
      uint32_t ado[D / 16][4];
#pragma unroll
      for (int kk = 0; kk < D / 16; ++kk)
        ldsm_x4(ado[kk], sdO + (warp * 16 + lane % 16) * DP + kk * 16 + (lane / 16) * 8);
      float accq[8][4];
#pragma unroll
      for (int dt = 0; dt < 8; ++dt) { accq[dt][0] = accq[dt][1] = accq[dt][2] = accq[dt][3] = 0.f; }
#pragma unroll
      for (int jt = 0; jt < NKT; ++jt) {
        if (jt * 16 < nkp) {
          float dp[2][4];
#pragma unroll
          for (int e = 0; e < 4; ++e) { dp[0][e] = 0.f; dp[1][e] = 0.f; }
#pragma unroll
          for (int kk = 0; kk < D / 16; ++kk) {
            uint32_t bb[4];
            ldsm_x4(bb, sV + (jt * 16 + (lane / 16) * 8 + (lane % 8)) * DP + kk * 16 + ((lane / 8) % 2) * 8);
            mma16816(dp[0], ado[kk], bb[0], bb[1]);
            mma16816(dp[1], ado[kk], bb[2], bb[3]);
          }
          uint32_t ads[4];
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const int nt = 2 * jt + half;
            const float mk0 = sMask[nt * 8 + 2 * t], mk1 = sMask[nt * 8 + 2 * t + 1];
            float p[4];
            p[0] = exp2f((((mk0 == 0.f) ? fmaxf(s[nt][0], -FLT_MAX) : mk0) - m_a) * kLog2e) * il_a;
            p[1] = exp2f((((mk1 == 0.f) ? fmaxf(s[nt][1], -FLT_MAX) : mk1) - m_a) * kLog2e) * il_a;
            p[2] = exp2f((((mk0 == 0.f) ? fmaxf(s[nt][2], -FLT_MAX) : mk0) - m_b) * kLog2e) * il_b;
            p[3] = exp2f((((mk1 == 0.f) ? fmaxf(s[nt][3], -FLT_MAX) : mk1) - m_b) * kLog2e) * il_b;
            // masked entries tie in the reference's clamp max(S + mask, finfo.min): torch halves a tie's gradient
            const float h0 = (mk0 == 0.f) ? 1.f : 0.5f, h1 = (mk1 == 0.f) ? 1.f : 0.5f;
            const float ds0 = h0 * p[0] * (dp[half][0] - del_a), ds1 = h1 * p[1] * (dp[half][1] - del_a);
            const float ds2 = h0 * p[2] * (dp[half][2] - del_b), ds3 = h1 * p[3] * (dp[half][3] - del_b);
            const uint32_t pa = pack_bf16(p[0], p[1]), pb = pack_bf16(p[2], p[3]);
            const uint32_t da = pack_bf16(ds0, ds1), db = pack_bf16(ds2, ds3);
            *reinterpret_cast<uint32_t*>(sP + ra * PP + nt * 8 + 2 * t) = pa;
            *reinterpret_cast<uint32_t*>(sP + rb * PP + nt * 8 + 2 * t) = pb;
            *reinterpret_cast<uint32_t*>(sdS + ra * PP + nt * 8 + 2 * t) = da;
            *reinterpret_cast<uint32_t*>(sdS + rb * PP + nt * 8 + 2 * t) = db;
            ads[half * 2] = da; ads[half * 2 + 1] = db;
          }
          // dQ[:, c0:c0+64] += dS(16 x 16 keys) * K[keys, c0:c0+64]
#pragma unroll
          for (int dpair = 0; dpair < 4; ++dpair) {
            uint32_t bb[4];
            ldsm_x4_t(bb, sK + (jt * 16 + ((lane / 8) % 2) * 8 + (lane % 8)) * DP + c0 + dpair * 16 + (lane / 16) * 8);
            mma16816(accq[2 * dpair], ads, bb[0], bb[1]);
            mma16816(accq[2 * dpair + 1], ads, bb[2], bb[3]);
          }
        }
      }
      if (r0 + ra < seq) {
        __nv_bfloat16* qp = dq + ((int64_t)b * seq + r0 + ra) * lddq + h * D + c0 + 2 * t;
#pragma unroll
        for (int dt = 0; dt < 8; ++dt) *reinterpret_cast<uint32_t*>(qp + dt * 8) = pack_bf16(accq[dt][0], accq[dt][1]);
      }
      if (r0 + rb < seq) {
        __nv_bfloat16* qp = dq + ((int64_t)b * seq + r0 + rb) * lddq + h * D + c0 + 2 * t;
#pragma unroll
        for (int dt = 0; dt < 8; ++dt) *reinterpret_cast<uint32_t*>(qp + dt * 8) = pack_bf16(accq[dt][2], accq[dt][3]);
      }
    }
    __syncthreads();

    // ---- phase 2: dV[j-tile, slice] += P^T dO[:, slice] ; dK[j-tile, slice] += dS^T Q[:, slice]
#pragma unroll
    for (int i = 0; i < JT_PER_WARP; ++i) {
      const int jt = warp + i * W;
      if (jt * 16 < nkp) {
#pragma unroll 2
        for (int qk = 0; qk < R / 16; ++qk) {
          uint32_t ap[4], as_[4];
          const int krow = qk * 16 + (lane / 16) * 8 + (lane % 8);
          const int mcol = jt * 16 + ((lane / 8) % 2) * 8;
          ldsm_x4_t(ap, sP + krow * PP + mcol);
          ldsm_x4_t(as_, sdS + krow * PP + mcol);
#pragma unroll
          for (int dpair = 0; dpair < 4; ++dpair) {
            uint32_t bo[4], bq[4];
            const int brow = qk * 16 + ((lane / 8) % 2) * 8 + (lane % 8);
            const int bcol = c0 + dpair * 16 + (lane / 16) * 8;
            ldsm_x4_t(bo, sdO + brow * DP + bcol);
            ldsm_x4_t(bq, sQ + brow * DP + bcol);
            mma16816(acc_dv[i][2 * dpair], ap, bo[0], bo[1]);
            mma16816(acc_dv[i][2 * dpair + 1], ap, bo[2], bo[3]);
            mma16816(acc_dk[i][2 * dpair], as_, bq[0], bq[1]);
            mma16816(acc_dk[i][2 * dpair + 1], as_, bq[2], bq[3]);
          }
        }
      }
    }
  }

#pragma unroll
  for (int i = 0; i < JT_PER_WARP; ++i) {
    const int jt = warp + i * W;
    if (jt * 16 < nkp) {
      const int ja = jt * 16 + g, jb = ja + 8;
      if (ja < nk) {
        __nv_bfloat16* kp = dk + ((int64_t)b * nk + ja) * lddk + h * D + c0 + 2 * t;
        __nv_bfloat16* vp = dv + ((int64_t)b * nk + ja) * lddv + h * D + c0 + 2 * t;
#pragma unroll
        for (int dt = 0; dt < 8; ++dt) {
          *reinterpret_cast<uint32_t*>(kp + dt * 8) = pack_bf16(acc_dk[i][dt][0], acc_dk[i][dt][1]);
          *reinterpret_cast<uint32_t*>(vp + dt * 8) = pack_bf16(acc_dv[i][dt][0], acc_dv[i][dt][1]);
        }
      }
      if (jb < nk) {
        __nv_bfloat16* kp = dk + ((int64_t)b * nk + jb) * lddk + h * D + c0 + 2 * t;
        __nv_bfloat16* vp = dv + ((int64_t)b * nk + jb) * lddv + h * D + c0 + 2 * t;
#pragma unroll
        for (int dt = 0; dt < 8; ++dt) {
          *reinterpret_cast<uint32_t*>(kp + dt * 8) = pack_bf16(acc_dk[i][dt][2], acc_dk[i][dt][3]);
          *reinterpret_cast<uint32_t*>(vp + dt * 8) = pack_bf16(acc_dv[i][dt][2], acc_dv[i][dt][3]);
        }
      }
    }
  }
}

int xattn_fwd_tc(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, const uint8_t* mask,
                 void* o, int64_t ldo, float* stats, int64_t batch, int64_t seq, int64_t nk, int64_t heads, int64_t d,
                 cudaStream_t stream);  // xattn_sm100.cu (tcgen05 + TMA)

template <int D, int NKT, int W>
static int launch_bwd(const void* d_o, int64_t lddo, const void* q, int64_t ldq, const void* k, int64_t ldk,
                      const void* v, int64_t ldv, const void* o, int64_t ldo, const float* stats, const uint8_t* mask,
                      void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv, int64_t batch,
                      int64_t seq, int64_t nk, int64_t heads, cudaStream_t stream) {
  const int nkp = ((int)nk + 15) & ~15;
  constexpr int R = W * 16;
  const size_t smem = (size_t)(2 * nkp + 2 * R) * (D + 8) * 2 + (size_t)2 * R * (NKT * 16 + 8) * 2 +
                      (NKT * 16 + R) * 4;
  auto kern = xattn_bwd_kernel<D, NKT, W>;
  MMGL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(D / 64, (unsigned)heads, (unsigned)batch);
  kern<<<grid, W * 32, smem, stream>>>((const __nv_bfloat16*)d_o, lddo, (const __nv_bfloat16*)q, ldq,
                                       (const __nv_bfloat16*)k, ldk, (const __nv_bfloat16*)v, ldv,
                                       (const __nv_bfloat16*)o, ldo, stats, mask, (__nv_bfloat16*)dq, lddq,
                                       (__nv_bfloat16*)dk, lddk, (__nv_bfloat16*)dv, lddv, (int)seq, (int)nk,
                                       (int)heads);
  return check_launch("mmgl_xattn_bwd");
}

static int check_xattn_args(const char* who, int64_t batch, int64_t seq, int64_t nk, int64_t heads, int64_t d,
                            std::initializer_list<int64_t> lds, std::initializer_list<const void*> ptrs) {
  MMGL_REQUIRE(batch > 0 && seq > 0 && nk > 0 && heads > 0, "%s: empty problem", who);
  MMGL_REQUIRE(d == 64 || d == 128, "%s: head_dim must be 64 or 128 (got %lld)", who, (long long)d);
  MMGL_REQUIRE(nk <= 256, "%s: Nk must be <= 256 (got %lld)", who, (long long)nk);
  MMGL_REQUIRE(batch < 65536 && heads < 65536, "%s: batch/heads too large for the grid", who);
  for (int64_t ld : lds) MMGL_REQUIRE(ld % 8 == 0, "%s: leading dims must be multiples of 8", who);
  for (const void* p : ptrs) MMGL_REQUIRE(p != nullptr && aligned16(p), "%s: pointers must be non-null, 16B aligned", who);
  return 0;
}

}  // namespace mmgl

using namespace mmgl;

extern "C" int mmgl_xattn_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                              const uint8_t* mask, void* o, int64_t ldo, float* stats, int64_t batch, int64_t seq,
                              int64_t nk, int64_t heads, int64_t d, void* stream_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  MMGL_BIND(q, "mmgl_xattn_fwd");
  if (int rc = check_xattn_args("mmgl_xattn_fwd", batch, seq, nk, heads, d, {ldq, ldk, ldv, ldo}, {q, k, v, o}))
    return rc;
  MMGL_REQUIRE(mask != nullptr && stats != nullptr, "mmgl_xattn_fwd: null mask/stats");
  MMGL_REQUIRE(d == 64 || nk <= 128, "mmgl_xattn_fwd: head_dim 128 supports Nk <= 128");
  return xattn_fwd_tc(q, ldq, k, ldk, v, ldv, mask, o, ldo, stats, batch, seq, nk, heads, d, s);
  return 0;
}

extern "C" int mmgl_xattn_bwd(const void* d_o, int64_t lddo, const void* q, int64_t ldq, const void* k, int64_t ldk,
                              const void* v, int64_t ldv, const void* o, int64_t ldo, const float* stats,
                              const uint8_t* mask, void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv,
                              int64_t lddv, int64_t batch, int64_t seq, int64_t nk, int64_t heads, int64_t d,
                              void* stream_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  MMGL_BIND(q, "mmgl_xattn_bwd");
  if (int rc = check_xattn_args("mmgl_xattn_bwd", batch, seq, nk, heads, d, {lddo, ldq, ldk, ldv, ldo, lddq, lddk, lddv},
                                {d_o, q, k, v, o, dq, dk, dv}))
    return rc;
  MMGL_REQUIRE(mask != nullptr && stats != nullptr, "mmgl_xattn_bwd: null mask/stats");
#define BWD(D_, NKT_, W_)                                                                                         \
  return launch_bwd<D_, NKT_, W_>(d_o, lddo, q, ldq, k, ldk, v, ldv, o, ldo, stats, mask, dq, lddq, dk, lddk, dv, \
                                  lddv, batch, seq, nk, heads, s)
  MMGL_REQUIRE(nk <= 128, "mmgl_xattn_bwd: Nk must be <= 128 (got %lld)", (long long)nk);
  if (d == 64) {
    if (nk <= 64) BWD(64, 4, 4);
    BWD(64, 8, 8);
  } else {
    if (nk <= 64) BWD(128, 4, 4);
    BWD(128, 8, 8);
  }
#undef BWD
  return 0;
}
