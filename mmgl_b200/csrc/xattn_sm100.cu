// Fused cross-attention core, forward, on the 5th-generation tensor cores (tcgen05 + TMEM + TMA).
//
//   O = softmax(max(Q K^T + mask, -FLT_MAX)) V          per (sample, head); Q pre-scaled by d^-1/2
//
// One CTA = one 128-query tile of one (sample, head):
//   * thread 0 TMA-loads the Q tile and the head's whole K / V neighbor bank (Nk <= 256 rows, zero-filled tails)
//     into 128B-swizzled shared memory (cp.async.bulk.tensor, one mbarrier),
//   * thread 0 issues tcgen05.mma  S[128 x Nk] = Q K^T  into TMEM (fp32),
//   * the 128 threads each own one query row = one TMEM lane: tcgen05.ld the row, byte-mask + clamp, fp32 row max /
//     exp2 / sum entirely in registers (no shuffles, the row never leaves the thread), un-normalised P written as bf16
//     into swizzled shared memory = the A operand of the second MMA,
//   * thread 0 issues tcgen05.mma  O[128 x d] = P V  (V consumed MN-major straight from its row-major tile),
//   * epilogue: tcgen05.ld O, scale by 1/row-sum, bf16, head-interleaved store; (row max, 1/sum) saved for backward.
// Nothing of shape [S, Nk] touches HBM, the additive mask of the reference is never built, and the head split /
// merge copies do not exist.  Several CTAs share an SM (48 KB smem, 128 TMEM columns at Nk = d = 64), which is what
// hides the TMA -> MMA -> softmax -> MMA dependency chain of a single tile.
//
// Replaces model/modelling_cross_attention.py:176-177, 206-271 and _expand_mask (:68-79).
#include <cfloat>
#include <cuda.h>
#include <cuda_bf16.h>

#include "../../include/mmgl_b200.h"
#include "common.cuh"
#include "ptx.cuh"

namespace mmgl {

constexpr float kLog2eF = 1.4426950408889634f;

// 16-byte chunk `chunk` (0..7) of row `row` inside a 128B-swizzled [rows][64 bf16] slab
__device__ __forceinline__ uint32_t swz128(uint32_t slab_base, int row, int chunk) {
  return slab_base + row * 128 + (((chunk ^ (row & 7)) & 7) << 4);
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

template <int D>
__global__ void __launch_bounds__(128)
xattn_fwd_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                    const __grid_constant__ CUtensorMap map_v, const uint8_t* __restrict__ mask,
                    __nv_bfloat16* __restrict__ o, int64_t ldo, float* __restrict__ stats, int seq, int nk, int nkp,
                    int heads, uint32_t tmem_cols) {
  constexpr int DS = D / 64;  // 64-wide (128-byte) slabs along the head dim
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int ps = (nkp + 63) / 64;                       // slabs of P along the key dim
  uint8_t* sQ = smem;                                   // DS x [128][64]
  uint8_t* sK = sQ + DS * 16384;                        // DS x [nkp][64]   (K-major B operand of S = Q K^T)
  uint8_t* sV = sK + DS * nkp * 128;                    // DS x [nkp][64]   (MN-major B operand of O = P V)
  uint8_t* sP = sV + DS * nkp * 128;                    // ps x [128][64]
  float* sMask = reinterpret_cast<float*>(sP + ps * 16384);          // [ps * 64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sMask + ps * 64);     // load, mma1, mma2
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 3);

  const int b = blockIdx.z, h = blockIdx.y, r0 = blockIdx.x * 128;
  const int warp = threadIdx.x >> 5, tid = threadIdx.x;

  if (tid == 0) {
    tma_prefetch_desc(&map_q); tma_prefetch_desc(&map_k); tma_prefetch_desc(&map_v);
    mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init(&bars[2], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_dyn(tmem_ptr, tmem_cols);
  for (int j = tid; j < ps * 64; j += 128)
    sMask[j] = (j < nk) ? (mask[(int64_t)b * nk + j] ? 0.f : -FLT_MAX) : -INFINITY;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (tid == 0) {
    mbar_arrive_expect_tx(&bars[0], DS * (16384 + 2 * nkp * 128));
#pragma unroll
    for (int j = 0; j < DS; ++j) {
      tma_load_2d(sQ + j * 16384, &map_q, &bars[0], h * D + 64 * j, b * seq + r0);
      tma_load_2d(sK + j * nkp * 128, &map_k, &bars[0], h * D + 64 * j, b * nk);
      tma_load_2d(sV + j * nkp * 128, &map_v, &bars[0], h * D + 64 * j, b * nk);
    }
    mbar_wait(&bars[0], 0);
    tc_fence_after();
    // S[128 x nkp] = Q K^T : A, B both K-major
    const uint32_t idesc = make_idesc_bf16(128, nkp, 0, 0);
    const uint32_t aq = smem_u32(sQ), bk = smem_u32(sK);
#pragma unroll
    for (int k = 0; k < D / 16; ++k) {
      const uint64_t da = make_smem_desc(aq + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024);
      const uint64_t db = make_smem_desc(bk + (k >> 2) * nkp * 128 + (k & 3) * 32, 16, 1024);
      umma_f16_ss(tmem_base, da, db, idesc, k != 0 ? 1u : 0u);
    }
    umma_commit(&bars[1]);
  }
  __syncwarp();

  // ---- softmax: this thread owns query row `tid` (TMEM lane tid)
  mbar_wait(&bars[1], 0);
  tc_fence_after();
  const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
  const int nchunks = (nkp + 31) / 32;
  float mx = -FLT_MAX;
  for (int c = 0; c < nchunks; ++c) {
    uint32_t r[32];
    tmem_ld_32x32(lane_addr + c * 32, r);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float mk = sMask[c * 32 + j];
      const float x = (mk == 0.f) ? fmaxf(__uint_as_float(r[j]), -FLT_MAX) : mk;   // max(s + mask, finfo.min)
      mx = fmaxf(mx, x);
    }
  }
  float sum = 0.f;
  const uint32_t p_base = smem_u32(sP);
  for (int c = 0; c < nchunks; ++c) {
    uint32_t r[32];
    tmem_ld_32x32(lane_addr + c * 32, r);
    tmem_ld_wait();
    float p[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float mk = sMask[c * 32 + j];
      const float x = (mk == 0.f) ? fmaxf(__uint_as_float(r[j]), -FLT_MAX) : mk;
      p[j] = exp2f((x - mx) * kLog2eF);
      sum += p[j];
    }
    const uint32_t slab = p_base + (c >> 1) * 16384;
#pragma unroll
    for (int g = 0; g < 4; ++g)
      st_shared_v4(swz128(slab, tid, (c & 1) * 4 + g), pack_bf16(p[8 * g], p[8 * g + 1]),
                   pack_bf16(p[8 * g + 2], p[8 * g + 3]), pack_bf16(p[8 * g + 4], p[8 * g + 5]),
                   pack_bf16(p[8 * g + 6], p[8 * g + 7]));
  }
  const float inv = 1.f / sum;
  fence_proxy_async();   // generic-proxy smem writes (P) -> visible to the tensor core's async proxy
  tc_fence_before();
  __syncthreads();

  if (tid == 0) {
    tc_fence_after();
    // O[128 x D] = P V : A = P K-major (K = keys), B = V MN-major ([key rows][d cols] as loaded)
    const uint32_t idesc = make_idesc_bf16(128, D, 0, 1);
    const uint32_t bv = smem_u32(sV);
    for (int k = 0; k < nkp / 16; ++k) {
      const uint64_t da = make_smem_desc(p_base + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024);
      const uint64_t db = make_smem_desc(bv + k * 2048, nkp * 128, 1024);
      umma_f16_ss(tmem_base + ps * 64, da, db, idesc, k != 0 ? 1u : 0u);
    }
    umma_commit(&bars[2]);
  }
  __syncwarp();
  mbar_wait(&bars[2], 0);
  tc_fence_after();

  const int row = r0 + tid;
  const uint32_t o_addr = lane_addr + ps * 64;
#pragma unroll
  for (int c = 0; c < D / 32; ++c) {
    uint32_t r[32];
    tmem_ld_32x32(o_addr + c * 32, r);
    tmem_ld_wait();
    if (row < seq) {
      __nv_bfloat16* op = o + ((int64_t)b * seq + row) * ldo + h * D + c * 32;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint4 v;
        v.x = pack_bf16(__uint_as_float(r[8 * g]) * inv, __uint_as_float(r[8 * g + 1]) * inv);
        v.y = pack_bf16(__uint_as_float(r[8 * g + 2]) * inv, __uint_as_float(r[8 * g + 3]) * inv);
        v.z = pack_bf16(__uint_as_float(r[8 * g + 4]) * inv, __uint_as_float(r[8 * g + 5]) * inv);
        v.w = pack_bf16(__uint_as_float(r[8 * g + 6]) * inv, __uint_as_float(r[8 * g + 7]) * inv);
        *reinterpret_cast<uint4*>(op + 8 * g) = v;
      }
    }
  }
  if (row < seq) {
    float* st = stats + (((int64_t)b * heads + h) * seq + row) * 2;
    *reinterpret_cast<float2*>(st) = make_float2(mx, inv);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc_dyn(tmem_base, tmem_cols);
}

int make_tensor_map_2d(CUtensorMap* out, const void* ptr, uint64_t d0, uint64_t d1, uint64_t ld, uint32_t b0, uint32_t b1);

template <int D>
static int launch_fwd_tc(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                         const uint8_t* mask, void* o, int64_t ldo, float* stats, int64_t batch, int64_t seq, int64_t nk,
                         int64_t heads, cudaStream_t stream) {
  const int nkp = ((int)nk + 15) & ~15;
  const int ps = (nkp + 63) / 64;
  constexpr int DS = D / 64;
  const uint32_t need_cols = (uint32_t)(ps * 64 + D);
  uint32_t tmem_cols = 32;
  while (tmem_cols < need_cols) tmem_cols <<= 1;
  const size_t smem = 1024 + (size_t)DS * 16384 + 2 * (size_t)DS * nkp * 128 + (size_t)ps * 16384 + ps * 64 * 4 + 64;
  CUtensorMap mq, mk, mv;
  int rc;
  if ((rc = make_tensor_map_2d(&mq, q, (uint64_t)(heads * D), (uint64_t)(batch * seq), (uint64_t)ldq, 64, 128))) return rc;
  if ((rc = make_tensor_map_2d(&mk, k, (uint64_t)(heads * D), (uint64_t)(batch * nk), (uint64_t)ldk, 64, (uint32_t)nkp))) return rc;
  if ((rc = make_tensor_map_2d(&mv, v, (uint64_t)(heads * D), (uint64_t)(batch * nk), (uint64_t)ldv, 64, (uint32_t)nkp))) return rc;
  auto kern = xattn_fwd_tc_kernel<D>;
  MMGL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)((seq + 127) / 128), (unsigned)heads, (unsigned)batch);
  kern<<<grid, 128, smem, stream>>>(mq, mk, mv, mask, (__nv_bfloat16*)o, ldo, stats, (int)seq, (int)nk, nkp, (int)heads,
                                    tmem_cols);
  return check_launch("mmgl_xattn_fwd");
}

int xattn_fwd_tc(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, const uint8_t* mask,
                 void* o, int64_t ldo, float* stats, int64_t batch, int64_t seq, int64_t nk, int64_t heads, int64_t d,
                 cudaStream_t stream) {
  if (d == 64) return launch_fwd_tc<64>(q, ldq, k, ldk, v, ldv, mask, o, ldo, stats, batch, seq, nk, heads, stream);
  return launch_fwd_tc<128>(q, ldq, k, ldk, v, ldv, mask, o, ldo, stats, batch, seq, nk, heads, stream);
}

// ------------------------------------------------------------------------------------------------ backward
// One CTA per (sample, head): K and V stay resident in shared memory, dK / dV accumulate in TMEM across all 128-query
// tiles.  Per tile (thread 0 issues, the 128 threads own one query row = one TMEM lane each):
//     S  = Q K^T,  dP = dO V^T                      (tcgen05.mma, both into TMEM)
//     P  = exp2(S~ - m) / l,  dS = P (dP - delta)   (registers; S~ = masked/clamped scores, (m, 1/l) from forward,
//                                                    delta = rowsum(dO * O); masked entries get the reference's 1/2)
//     P, dS -> bf16 -> swizzled shared memory; the SAME bytes serve as K-major A operand (dS K) and, read MN-major,
//     as the transposed A operand of the two reductions over queries:
//     dQ = dS K,   dV += P^T dO,   dK += dS^T Q     (tcgen05.mma; Q / dO / K tiles reused MN-major as B operands)
// Nothing of shape [S, Nk] touches HBM.  Replaces the autograd backward of model/modelling_cross_attention.py:206-271.
template <int D>
__global__ void __launch_bounds__(128)
xattn_bwd_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_do,
                    const __grid_constant__ CUtensorMap map_k, const __grid_constant__ CUtensorMap map_v,
                    const __nv_bfloat16* __restrict__ o, int64_t ldo, const __nv_bfloat16* __restrict__ d_o, int64_t lddo,
                    const float* __restrict__ stats, const uint8_t* __restrict__ mask,
                    __nv_bfloat16* __restrict__ dq, int64_t lddq, __nv_bfloat16* __restrict__ dk, int64_t lddk,
                    __nv_bfloat16* __restrict__ dv, int64_t lddv, int seq, int nk, int nkp, int heads,
                    uint32_t tmem_cols, int alias_dq) {
  constexpr int DS = D / 64;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int ps = (nkp + 63) / 64;
  uint8_t* sK = smem;                         // DS x [nkp][64]
  uint8_t* sV = sK + DS * nkp * 128;          // DS x [nkp][64]
  uint8_t* sQ = sV + DS * nkp * 128;          // DS x [128][64]
  uint8_t* sdO = sQ + DS * 16384;             // DS x [128][64]
  uint8_t* sP = sdO + DS * 16384;             // ps x [128][64]
  uint8_t* sdS = sP + ps * 16384;             // ps x [128][64]  (+ one slab of slack: M = 128 reads a 2nd key atom)
  float* sMask = reinterpret_cast<float*>(sdS + (ps + 1) * 16384);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sMask + ps * 64);   // kv, q, a, b
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 4);

  const int h = blockIdx.x, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, tid = threadIdx.x;
  const int ntiles = (seq + 127) / 128;
  const int nchunks = (nkp + 31) / 32;

  if (tid == 0) {
    tma_prefetch_desc(&map_q); tma_prefetch_desc(&map_do); tma_prefetch_desc(&map_k); tma_prefetch_desc(&map_v);
    for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_dyn(tmem_ptr, tmem_cols);
  for (int j = tid; j < ps * 64; j += 128)
    sMask[j] = (j < nk) ? (mask[(int64_t)b * nk + j] ? 0.f : -FLT_MAX) : -INFINITY;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t cS = 0, cdP = ps * 64, cdQ = alias_dq ? 0 : 2 * ps * 64;
  const uint32_t cdV = (alias_dq ? 2 * ps * 64 : 2 * ps * 64 + D), cdK = cdV + D;
  const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
  const uint32_t aQ = smem_u32(sQ), adO = smem_u32(sdO), aK = smem_u32(sK), aV = smem_u32(sV), aP = smem_u32(sP),
                 adS = smem_u32(sdS);

  if (tid == 0) {
    mbar_arrive_expect_tx(&bars[0], 2 * DS * nkp * 128);
#pragma unroll
    for (int j = 0; j < DS; ++j) {
      tma_load_2d(sK + j * nkp * 128, &map_k, &bars[0], h * D + 64 * j, b * nk);
      tma_load_2d(sV + j * nkp * 128, &map_v, &bars[0], h * D + 64 * j, b * nk);
    }
  }

  for (int t = 0; t < ntiles; ++t) {
    const int r0 = t * 128;
    const uint32_t par = t & 1;
    if (tid == 0) {
      mbar_arrive_expect_tx(&bars[1], 2 * DS * 16384);
#pragma unroll
      for (int j = 0; j < DS; ++j) {
        tma_load_2d(sQ + j * 16384, &map_q, &bars[1], h * D + 64 * j, b * seq + r0);
        tma_load_2d(sdO + j * 16384, &map_do, &bars[1], h * D + 64 * j, b * seq + r0);
      }
    }
    // row statistics and delta = sum_d dO * O for this thread's query row (overlaps the TMA)
    const int row = r0 + tid;
    const bool row_ok = row < seq;
    float m = 0.f, inv = 0.f, delta = 0.f;
    if (row_ok) {
      const float2 st = *reinterpret_cast<const float2*>(stats + (((int64_t)b * heads + h) * seq + row) * 2);
      m = st.x; inv = st.y;
      const uint4* po = reinterpret_cast<const uint4*>(o + ((int64_t)b * seq + row) * ldo + h * D);
      const uint4* pd = reinterpret_cast<const uint4*>(d_o + ((int64_t)b * seq + row) * lddo + h * D);
#pragma unroll
      for (int i = 0; i < D / 8; ++i) {
        const uint4 vo = __ldg(po + i), vd = __ldg(pd + i);
        const uint32_t wo[4] = {vo.x, vo.y, vo.z, vo.w}, wd[4] = {vd.x, vd.y, vd.z, vd.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) delta += bf16lo(wo[e]) * bf16lo(wd[e]) + bf16hi(wo[e]) * bf16hi(wd[e]);
      }
    }
    if (tid == 0) {
      if (t == 0) mbar_wait(&bars[0], 0);
      mbar_wait(&bars[1], par);
      tc_fence_after();
      const uint32_t idesc = make_idesc_bf16(128, nkp, 0, 0);
#pragma unroll
      for (int k = 0; k < D / 16; ++k) {   // S = Q K^T
        const uint64_t da = make_smem_desc(aQ + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024);
        const uint64_t db = make_smem_desc(aK + (k >> 2) * nkp * 128 + (k & 3) * 32, 16, 1024);
        umma_f16_ss(tmem_base + cS, da, db, idesc, k != 0 ? 1u : 0u);
      }
#pragma unroll
      for (int k = 0; k < D / 16; ++k) {   // dP = dO V^T
        const uint64_t da = make_smem_desc(adO + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024);
        const uint64_t db = make_smem_desc(aV + (k >> 2) * nkp * 128 + (k & 3) * 32, 16, 1024);
        umma_f16_ss(tmem_base + cdP, da, db, idesc, k != 0 ? 1u : 0u);
      }
      umma_commit(&bars[2]);
    }
    __syncwarp();
    mbar_wait(&bars[2], par);
    tc_fence_after();
    for (int c = 0; c < nchunks; ++c) {
      uint32_t rs[32], rp[32];
      tmem_ld_32x32(lane_addr + cS + c * 32, rs);
      tmem_ld_32x32(lane_addr + cdP + c * 32, rp);
      tmem_ld_wait();
      uint32_t pk[16], dk16[16];
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        float pv[2], dsv[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const float mk = sMask[c * 32 + j + e];
          const float x = (mk == 0.f) ? fmaxf(__uint_as_float(rs[j + e]), -FLT_MAX) : mk;
          const float p = row_ok ? exp2f((x - m) * kLog2eF) * inv : 0.f;
          // masked entries tie in the reference's clamp max(S + mask, finfo.min): torch halves a tie's gradient
          const float half = (mk == 0.f) ? 1.f : 0.5f;
          pv[e] = p;
          dsv[e] = row_ok ? half * p * (__uint_as_float(rp[j + e]) - delta) : 0.f;
        }
        pk[j >> 1] = pack_bf16(pv[0], pv[1]);
        dk16[j >> 1] = pack_bf16(dsv[0], dsv[1]);
      }
      const uint32_t slabP = aP + (c >> 1) * 16384, slabS = adS + (c >> 1) * 16384;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        st_shared_v4(swz128(slabP, tid, (c & 1) * 4 + g), pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]);
        st_shared_v4(swz128(slabS, tid, (c & 1) * 4 + g), dk16[4 * g], dk16[4 * g + 1], dk16[4 * g + 2], dk16[4 * g + 3]);
      }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      {  // dQ[128 x D] = dS K : A = dS K-major (K = keys), B = K tile read MN-major
        const uint32_t idesc = make_idesc_bf16(128, D, 0, 1);
        for (int k = 0; k < nkp / 16; ++k) {
          const uint64_t da = make_smem_desc(adS + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024);
          const uint64_t db = make_smem_desc(aK + k * 2048, nkp * 128, 1024);
          umma_f16_ss(tmem_base + cdQ, da, db, idesc, k != 0 ? 1u : 0u);
        }
      }
      {  // dV[keys x D] += P^T dO ; dK[keys x D] += dS^T Q : A read MN-major (M = keys), B = dO / Q tiles read MN-major
        const uint32_t idesc = make_idesc_bf16(128, D, 1, 1);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint64_t da = make_smem_desc(aP + k * 2048, 16384, 1024);
          const uint64_t db = make_smem_desc(adO + k * 2048, 16384, 1024);
          umma_f16_ss(tmem_base + cdV, da, db, idesc, (t | k) != 0 ? 1u : 0u);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint64_t da = make_smem_desc(adS + k * 2048, 16384, 1024);
          const uint64_t db = make_smem_desc(aQ + k * 2048, 16384, 1024);
          umma_f16_ss(tmem_base + cdK, da, db, idesc, (t | k) != 0 ? 1u : 0u);
        }
      }
      umma_commit(&bars[3]);
    }
    __syncwarp();
    mbar_wait(&bars[3], par);
    tc_fence_after();
#pragma unroll
    for (int c = 0; c < D / 32; ++c) {
      uint32_t r[32];
      tmem_ld_32x32(lane_addr + cdQ + c * 32, r);
      tmem_ld_wait();
      if (row_ok) {
        __nv_bfloat16* qp = dq + ((int64_t)b * seq + row) * lddq + h * D + c * 32;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 v;
          v.x = pack_bf16(__uint_as_float(r[8 * g]), __uint_as_float(r[8 * g + 1]));
          v.y = pack_bf16(__uint_as_float(r[8 * g + 2]), __uint_as_float(r[8 * g + 3]));
          v.z = pack_bf16(__uint_as_float(r[8 * g + 4]), __uint_as_float(r[8 * g + 5]));
          v.w = pack_bf16(__uint_as_float(r[8 * g + 6]), __uint_as_float(r[8 * g + 7]));
          *reinterpret_cast<uint4*>(qp + 8 * g) = v;
        }
      }
    }
    tc_fence_before();
    __syncthreads();   // dQ (aliasing S) fully read, Q / dO / P / dS tiles free for the next tile
    tc_fence_after();
  }

  // dK / dV rows: TMEM lane = key index
#pragma unroll
  for (int which = 0; which < 2; ++which) {
    const uint32_t col = which ? cdK : cdV;
    __nv_bfloat16* base = which ? dk : dv;
    const int64_t ld = which ? lddk : lddv;
#pragma unroll
    for (int c = 0; c < D / 32; ++c) {
      uint32_t r[32];
      tmem_ld_32x32(lane_addr + col + c * 32, r);
      tmem_ld_wait();
      if (tid < nk) {
        __nv_bfloat16* kp = base + ((int64_t)b * nk + tid) * ld + h * D + c * 32;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 v;
          v.x = pack_bf16(__uint_as_float(r[8 * g]), __uint_as_float(r[8 * g + 1]));
          v.y = pack_bf16(__uint_as_float(r[8 * g + 2]), __uint_as_float(r[8 * g + 3]));
          v.z = pack_bf16(__uint_as_float(r[8 * g + 4]), __uint_as_float(r[8 * g + 5]));
          v.w = pack_bf16(__uint_as_float(r[8 * g + 6]), __uint_as_float(r[8 * g + 7]));
          *reinterpret_cast<uint4*>(kp + 8 * g) = v;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc_dyn(tmem_base, tmem_cols);
}

template <int D>
static int launch_bwd_tc(const void* d_o, int64_t lddo, const void* q, int64_t ldq, const void* k, int64_t ldk,
                         const void* v, int64_t ldv, const void* o, int64_t ldo, const float* stats, const uint8_t* mask,
                         void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv, int64_t batch,
                         int64_t seq, int64_t nk, int64_t heads, cudaStream_t stream) {
  const int nkp = ((int)nk + 15) & ~15;
  const int ps = (nkp + 63) / 64;
  constexpr int DS = D / 64;
  const int alias_dq = (ps * 64 >= D) ? 1 : 0;
  const uint32_t need_cols = (uint32_t)(2 * ps * 64 + (alias_dq ? 2 : 3) * D);
  if (need_cols > 512) { set_error("mmgl_xattn_bwd: Nk = %lld with head_dim %d needs more than 512 TMEM columns", (long long)nk, D); return 2; }
  uint32_t tmem_cols = 32;
  while (tmem_cols < need_cols) tmem_cols <<= 1;
  const size_t smem = 1024 + 2 * (size_t)DS * nkp * 128 + 2 * (size_t)DS * 16384 + (size_t)(2 * ps + 1) * 16384 + ps * 64 * 4 + 64;
  CUtensorMap mq, mdo, mk, mv;
  int rc;
  if ((rc = make_tensor_map_2d(&mq, q, (uint64_t)(heads * D), (uint64_t)(batch * seq), (uint64_t)ldq, 64, 128))) return rc;
  if ((rc = make_tensor_map_2d(&mdo, d_o, (uint64_t)(heads * D), (uint64_t)(batch * seq), (uint64_t)lddo, 64, 128))) return rc;
  if ((rc = make_tensor_map_2d(&mk, k, (uint64_t)(heads * D), (uint64_t)(batch * nk), (uint64_t)ldk, 64, (uint32_t)nkp))) return rc;
  if ((rc = make_tensor_map_2d(&mv, v, (uint64_t)(heads * D), (uint64_t)(batch * nk), (uint64_t)ldv, 64, (uint32_t)nkp))) return rc;
  auto kern = xattn_bwd_tc_kernel<D>;
  MMGL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)heads, (unsigned)batch);
  kern<<<grid, 128, smem, stream>>>(mq, mdo, mk, mv, (const __nv_bfloat16*)o, ldo, (const __nv_bfloat16*)d_o, lddo, stats,
                                    mask, (__nv_bfloat16*)dq, lddq, (__nv_bfloat16*)dk, lddk, (__nv_bfloat16*)dv, lddv,
                                    (int)seq, (int)nk, nkp, (int)heads, tmem_cols, alias_dq);
  return check_launch("mmgl_xattn_bwd");
}

int xattn_bwd_tc(const void* d_o, int64_t lddo, const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                 int64_t ldv, const void* o, int64_t ldo, const float* stats, const uint8_t* mask, void* dq, int64_t lddq,
                 void* dk, int64_t lddk, void* dv, int64_t lddv, int64_t batch, int64_t seq, int64_t nk, int64_t heads,
                 int64_t d, cudaStream_t stream) {
  if (d == 64)
    return launch_bwd_tc<64>(d_o, lddo, q, ldq, k, ldk, v, ldv, o, ldo, stats, mask, dq, lddq, dk, lddk, dv, lddv, batch, seq,
                             nk, heads, stream);
  return launch_bwd_tc<128>(d_o, lddo, q, ldq, k, ldk, v, ldv, o, ldo, stats, mask, dq, lddq, dk, lddk, dv, lddv, batch, seq,
                            nk, heads, stream);
}

}  // namespace mmgl
