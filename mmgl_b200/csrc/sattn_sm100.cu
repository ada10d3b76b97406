// Causal / key-padded self-attention of the frozen decoder layers, tcgen05 + TMEM + TMA.
//
//   O = softmax(max(scale * Q K^T + causal + key-padding, finfo.min)) V      per (sample, head)
//
// Reference: MPTAttention self branch, model/modelling_cross_attention.py:201-275 with the additive mask built at
// :455-476 (causal AND key-not-padding).  Sequences here are short (S <= ~1200), so the kernels use 128 x 128 score blocks
// and a TWO-PASS softmax instead of online rescaling: pass 1 recomputes nothing but the row maximum (S = Q K_j^T per key
// block), pass 2 recomputes S, exponentiates against the final maximum and accumulates O += P_j V_j in TMEM.  One extra
// QK^T per block buys the absence of any accumulator correction; the tensor pipe is far from the limit at these sizes.
//
//   forward : CTA = (128-query tile, head, sample); K/V blocks double-buffered by TMA; only blocks j <= i when causal
//   backward: dQ kernel   CTA = (query tile i): loops key blocks j, dQ_i += dS_ij K_j            (accumulates in TMEM)
//             dK/dV kernel CTA = (key block j): loops query tiles i, dV_j += P_ij^T dO_i, dK_j += dS_ij^T Q_i (TMEM)
//             both recompute S and dP = dO V^T from Q, K, V, dO and the saved row statistics (m, 1/l).
// Each of the 128 threads owns one score row = one TMEM lane, so the softmax needs no shuffles.
#include <cfloat>
#include <cuda.h>
#include <cuda_bf16.h>

#include "../../include/mmgl_b200.h"
#include "common.cuh"
#include "ptx.cuh"

namespace mmgl {

int make_tensor_map_2d(CUtensorMap* out, const void* ptr, uint64_t d0, uint64_t d1, uint64_t ld, uint32_t b0, uint32_t b1);

namespace {

constexpr float kL2E = 1.4426950408889634f;

__device__ __forceinline__ uint32_t swz(uint32_t slab_base, int row, int chunk) {
  return slab_base + row * 128 + (((chunk ^ (row & 7)) & 7) << 4);
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// masked score of (query row, key col): scale * s where attended, -FLT_MAX where masked (the reference's finfo.min after
// the clamp), -inf for keys beyond the sequence (they do not exist in the reference)
__device__ __forceinline__ float masked_score(float s, float scale, uint8_t key_flag, int key, int row, int causal) {
  if (key_flag != 0) return key_flag == 1 ? -FLT_MAX : -INFINITY;
  if (causal && key > row) return -FLT_MAX;
  return fmaxf(s * scale, -FLT_MAX);
}
// key flags of one 128-key block into smem: 0 = attend, 1 = padding key (masked), 2 = beyond the sequence
__device__ __forceinline__ void load_key_flags(uint8_t* dst, const uint8_t* key_mask, int b, int seq, int k0, int tid) {
  const int key = k0 + tid;
  uint8_t f = 2;
  if (key < seq) f = (key_mask == nullptr || key_mask[(int64_t)b * seq + key]) ? 0 : 1;
  dst[tid] = f;
}
__device__ __forceinline__ void store_row_bf16(__nv_bfloat16* dst, const uint32_t (&r)[32], float mul) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    uint4 v;
    v.x = pack_bf16(__uint_as_float(r[8 * g]) * mul, __uint_as_float(r[8 * g + 1]) * mul);
    v.y = pack_bf16(__uint_as_float(r[8 * g + 2]) * mul, __uint_as_float(r[8 * g + 3]) * mul);
    v.z = pack_bf16(__uint_as_float(r[8 * g + 4]) * mul, __uint_as_float(r[8 * g + 5]) * mul);
    v.w = pack_bf16(__uint_as_float(r[8 * g + 6]) * mul, __uint_as_float(r[8 * g + 7]) * mul);
    *reinterpret_cast<uint4*>(dst + 8 * g) = v;
  }
}

// issue S[128 x 128] = A[128 x D] * B[128 x D]^T (both K-major tiles of DS 64-wide slabs) into TMEM column `col`
template <int D>
__device__ __forceinline__ void mma_qk(uint32_t tmem, uint32_t a_base, uint32_t b_base) {
  const uint32_t idesc = make_idesc_bf16(128, 128, 0, 0);
#pragma unroll
  for (int k = 0; k < D / 16; ++k) {
    const uint64_t da = make_smem_desc(a_base + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024);
    const uint64_t db = make_smem_desc(b_base + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024);
    umma_f16_ss(tmem, da, db, idesc, k != 0 ? 1u : 0u);
  }
}
// issue C[128 x D] (+)= A[128 x 128] * B[128 x D]: A K-major (2 slabs of 64), B a [128 rows][D] tile read MN-major
template <int D>
__device__ __forceinline__ void mma_pv(uint32_t tmem, uint32_t a_base, uint32_t b_base, bool accumulate) {
  const uint32_t idesc = make_idesc_bf16(128, D, 0, 1);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const uint64_t da = make_smem_desc(a_base + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024);
    const uint64_t db = make_smem_desc(b_base + k * 2048, 16384, 1024);
    umma_f16_ss(tmem, da, db, idesc, (accumulate || k != 0) ? 1u : 0u);
  }
}
// issue C[128 x D] (+)= A^T * B: A a [128 rows][128] tile read MN-major (transposed), B a [128 rows][D] tile read MN-major
template <int D>
__device__ __forceinline__ void mma_tn(uint32_t tmem, uint32_t a_base, uint32_t b_base, bool accumulate) {
  const uint32_t idesc = make_idesc_bf16(128, D, 1, 1);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const uint64_t da = make_smem_desc(a_base + k * 2048, 16384, 1024);
    const uint64_t db = make_smem_desc(b_base + k * 2048, 16384, 1024);
    umma_f16_ss(tmem, da, db, idesc, (accumulate || k != 0) ? 1u : 0u);
  }
}
template <int D>
__device__ __forceinline__ void tma_tile(uint8_t* dst, const CUtensorMap* map, uint64_t* bar, int col0, int row0) {
#pragma unroll
  for (int j = 0; j < D / 64; ++j) tma_load_2d(dst + j * 16384, map, bar, col0 + 64 * j, row0);
}

// ------------------------------------------------------------------------------------------------ forward
template <int D>
__global__ void __launch_bounds__(128)
sattn_fwd_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                 const __grid_constant__ CUtensorMap map_v, const uint8_t* __restrict__ key_mask,
                 __nv_bfloat16* __restrict__ o, int64_t ldo, float* __restrict__ stats, int seq, int heads, float scale,
                 int causal, uint32_t tmem_cols) {
  constexpr int TB = (D / 64) * 16384;   // bytes of one [128][D] tile
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + TB;          // 2 buffers
  uint8_t* sV = sK + 2 * TB;      // 2 buffers
  uint8_t* sP = sV + 2 * TB;      // [128][128] bf16 = 2 slabs
  uint8_t* sFlag = sP + 32768;   // [2][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sFlag + 256);   // q, k0, k1, v0, v1, s, o
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 7);

  const int ntiles = (seq + 127) / 128;
  const int qt = ntiles - 1 - (int)blockIdx.x;   // heavy (late) query tiles first
  const int h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, tid = threadIdx.x;
  const int r0 = qt * 128, row = r0 + tid;
  const int nblk = causal ? qt + 1 : ntiles;
  const int colq = h * D, rowb = b * seq;

  if (tid == 0) {
    tma_prefetch_desc(&map_q); tma_prefetch_desc(&map_k); tma_prefetch_desc(&map_v);
    for (int i = 0; i < 7; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_dyn(tmem_ptr, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
  const uint32_t cO = 128;
  uint32_t kuse[2] = {0, 0}, vuse[2] = {0, 0}, sphase = 0, ophase = 0;   // thread 0 / all threads phase counters

  if (tid == 0) {
    mbar_arrive_expect_tx(&bars[0], TB);
    tma_tile<D>(sQ, &map_q, &bars[0], colq, rowb + r0);
    mbar_arrive_expect_tx(&bars[1], TB);
    tma_tile<D>(sK, &map_k, &bars[1], colq, rowb);
  }
  // ---------------- pass 1: row maximum (raw_mx: unscaled maximum over the unmasked interior blocks)
  float mx = -FLT_MAX, raw_mx = -FLT_MAX;
  for (int j = 0; j < nblk; ++j) {
    const int buf = j & 1;
    load_key_flags(sFlag + buf * 128, key_mask, b, seq, j * 128, tid);
    if (tid == 0) {
      if (j + 1 < nblk) {
        mbar_arrive_expect_tx(&bars[1 + (buf ^ 1)], TB);
        tma_tile<D>(sK + (buf ^ 1) * TB, &map_k, &bars[1 + (buf ^ 1)], colq, rowb + (j + 1) * 128);
      }
      if (j == 0) mbar_wait(&bars[0], 0);
      mbar_wait(&bars[1 + buf], kuse[buf] & 1); kuse[buf]++;
      tc_fence_after();
      mma_qk<D>(tmem_base, smem_u32(sQ), smem_u32(sK + buf * TB));
      umma_commit(&bars[5]);
    }
    // key flags visible; a block needs the per-element mask path only if it holds a padding / out-of-range key, or
    // is the diagonal block of a causal problem -- interior blocks take the 1-instruction-per-score path
    const bool masked = __syncthreads_or(sFlag[buf * 128 + tid] != 0) || (causal && j == qt);
    mbar_wait(&bars[5], sphase & 1); sphase++;
    tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t r[32];
      tmem_ld_32x32(lane_addr + c * 32, r);
      tmem_ld_wait();
      if (masked) {
#pragma unroll
        for (int e = 0; e < 32; ++e)
          mx = fmaxf(mx, masked_score(__uint_as_float(r[e]), scale, sFlag[buf * 128 + c * 32 + e], j * 128 + c * 32 + e, row, causal));
      } else {
#pragma unroll
        for (int e = 0; e < 32; ++e) raw_mx = fmaxf(raw_mx, __uint_as_float(r[e]));
      }
    }
    tc_fence_before();
    __syncthreads();   // S fully read before the next block's MMA overwrites it
    tc_fence_after();
  }
  // ---------------- pass 2: P = exp(x - max), O += P V
  if (tid == 0) {
    const int b0 = 0;
    mbar_arrive_expect_tx(&bars[1 + b0], TB);
    tma_tile<D>(sK + b0 * TB, &map_k, &bars[1 + b0], colq, rowb);
    mbar_arrive_expect_tx(&bars[3 + b0], TB);
    tma_tile<D>(sV + b0 * TB, &map_v, &bars[3 + b0], colq, rowb);
  }
  mx = fmaxf(mx, fmaxf(raw_mx * scale, -FLT_MAX));   // scale > 0
  const float c1 = scale * kL2E, mxc = mx * kL2E;
  float sum = 0.f;
  const uint32_t p_base = smem_u32(sP);
  for (int j = 0; j < nblk; ++j) {
    const int buf = j & 1;
    load_key_flags(sFlag + buf * 128, key_mask, b, seq, j * 128, tid);
    if (tid == 0) {
      if (j + 1 < nblk) {
        mbar_arrive_expect_tx(&bars[1 + (buf ^ 1)], TB);
        tma_tile<D>(sK + (buf ^ 1) * TB, &map_k, &bars[1 + (buf ^ 1)], colq, rowb + (j + 1) * 128);
        mbar_arrive_expect_tx(&bars[3 + (buf ^ 1)], TB);
        tma_tile<D>(sV + (buf ^ 1) * TB, &map_v, &bars[3 + (buf ^ 1)], colq, rowb + (j + 1) * 128);
      }
      mbar_wait(&bars[1 + buf], kuse[buf] & 1); kuse[buf]++;
      tc_fence_after();
      mma_qk<D>(tmem_base, smem_u32(sQ), smem_u32(sK + buf * TB));
      umma_commit(&bars[5]);
    }
    const bool masked = __syncthreads_or(sFlag[buf * 128 + tid] != 0) || (causal && j == qt);
    mbar_wait(&bars[5], sphase & 1); sphase++;
    tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t r[32];
      tmem_ld_32x32(lane_addr + c * 32, r);
      tmem_ld_wait();
      uint32_t pk[16];
      if (masked) {
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          const float x0 = masked_score(__uint_as_float(r[e]), scale, sFlag[buf * 128 + c * 32 + e], j * 128 + c * 32 + e, row, causal);
          const float x1 = masked_score(__uint_as_float(r[e + 1]), scale, sFlag[buf * 128 + c * 32 + e + 1], j * 128 + c * 32 + e + 1, row, causal);
          const float p0 = exp2f((x0 - mx) * kL2E), p1 = exp2f((x1 - mx) * kL2E);
          sum += p0 + p1;
          pk[e >> 1] = pack_bf16(p0, p1);
        }
      } else {
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          const float p0 = exp2f(fmaf(__uint_as_float(r[e]), c1, -mxc)), p1 = exp2f(fmaf(__uint_as_float(r[e + 1]), c1, -mxc));
          sum += p0 + p1;
          pk[e >> 1] = pack_bf16(p0, p1);
        }
      }
      const uint32_t slab = p_base + (c >> 1) * 16384;
#pragma unroll
      for (int g = 0; g < 4; ++g) sts128(swz(slab, tid, (c & 1) * 4 + g), pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      mbar_wait(&bars[3 + buf], vuse[buf] & 1); vuse[buf]++;
      tc_fence_after();
      mma_pv<D>(tmem_base + cO, p_base, smem_u32(sV + buf * TB), j != 0);
      umma_commit(&bars[6]);
    }
    __syncwarp();
    mbar_wait(&bars[6], ophase & 1); ophase++;   // P tile, V buffer and the S columns are free again
    tc_fence_after();
  }
  const float inv = 1.f / sum;
#pragma unroll
  for (int c = 0; c < D / 32; ++c) {
    uint32_t r[32];
    tmem_ld_32x32(lane_addr + cO + c * 32, r);
    tmem_ld_wait();
    if (row < seq) store_row_bf16(o + ((int64_t)rowb + row) * ldo + colq + c * 32, r, inv);
  }
  if (row < seq) *reinterpret_cast<float2*>(stats + (((int64_t)b * heads + h) * seq + row) * 2) = make_float2(mx, inv);
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc_dyn(tmem_base, tmem_cols);
}

// per-row backward inputs of this thread: (max, 1/sum) and delta = rowsum(dO * O)
template <int D>
__device__ __forceinline__ void row_stats_delta(const float* stats, const __nv_bfloat16* o, int64_t ldo,
                                                const __nv_bfloat16* d_o, int64_t lddo, int b, int h, int heads, int seq,
                                                int row, float& m, float& inv, float& delta) {
  m = 0.f; inv = 0.f; delta = 0.f;
  if (row >= seq) return;
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

// P and dS of this thread's row for one 128-key block: reads S (col 0) and dP (col 128) from TMEM, writes bf16 tiles.
// kMasked = false is the interior-block path (no padding key, not the causal diagonal, no ragged query tile).
template <bool kWriteP, bool kMasked>
__device__ __forceinline__ void softmax_grad_block(uint32_t lane_addr, const uint8_t* flags, int key0, int row, bool row_ok,
                                                   int causal, float scale, float m, float inv, float delta,
                                                   uint32_t p_base, uint32_t ds_base, int tid) {
  const float c1 = scale * kL2E, mc = m * kL2E;
#pragma unroll 1
  for (int c = 0; c < 4; ++c) {
    uint32_t rs[32], rp[32];
    tmem_ld_32x32(lane_addr + c * 32, rs);
    tmem_ld_32x32(lane_addr + 128 + c * 32, rp);
    tmem_ld_wait();
    uint32_t pk[16], dk[16];
#pragma unroll
    for (int e = 0; e < 32; e += 2) {
      float pv[2], dv[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        float p;
        if (kMasked) {
          const float x = masked_score(__uint_as_float(rs[e + u]), scale, flags[c * 32 + e + u], key0 + c * 32 + e + u, row, causal);
          p = row_ok ? exp2f((x - m) * kL2E) * inv : 0.f;
        } else {
          p = exp2f(fmaf(__uint_as_float(rs[e + u]), c1, -mc)) * inv;
        }
        pv[u] = p;
        dv[u] = p * (__uint_as_float(rp[e + u]) - delta);
      }
      pk[e >> 1] = pack_bf16(pv[0], pv[1]);
      dk[e >> 1] = pack_bf16(dv[0], dv[1]);
    }
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      if (kWriteP) sts128(swz(p_base + (c >> 1) * 16384, tid, (c & 1) * 4 + g), pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]);
      sts128(swz(ds_base + (c >> 1) * 16384, tid, (c & 1) * 4 + g), dk[4 * g], dk[4 * g + 1], dk[4 * g + 2], dk[4 * g + 3]);
    }
  }
}

// ------------------------------------------------------------------------------------------------ backward: dQ
template <int D>
__global__ void __launch_bounds__(128)
sattn_bwd_dq_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_do,
                    const __grid_constant__ CUtensorMap map_k, const __grid_constant__ CUtensorMap map_v,
                    const uint8_t* __restrict__ key_mask, const __nv_bfloat16* __restrict__ o, int64_t ldo,
                    const __nv_bfloat16* __restrict__ d_o, int64_t lddo, const float* __restrict__ stats,
                    __nv_bfloat16* __restrict__ dq, int64_t lddq, int seq, int heads, float scale, int causal,
                    uint32_t tmem_cols) {
  constexpr int TB = (D / 64) * 16384;
  constexpr int NB = (D == 64) ? 2 : 1;   // K / V buffers
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sdO = sQ + TB;
  uint8_t* sK = sdO + TB;
  uint8_t* sV = sK + NB * TB;
  uint8_t* sdS = sV + NB * TB;     // 2 slabs
  uint8_t* sFlag = sdS + 32768;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sFlag + 256);   // q, kv0, kv1, a, b
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 5);

  const int ntiles = (seq + 127) / 128;
  const int qt = ntiles - 1 - (int)blockIdx.x;
  const int h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, tid = threadIdx.x;
  const int r0 = qt * 128, row = r0 + tid;
  const bool row_ok = row < seq;
  const int nblk = causal ? qt + 1 : ntiles;
  const int colq = h * D, rowb = b * seq;

  if (tid == 0) {
    tma_prefetch_desc(&map_q); tma_prefetch_desc(&map_do); tma_prefetch_desc(&map_k); tma_prefetch_desc(&map_v);
    for (int i = 0; i < 5; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_dyn(tmem_ptr, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
  const uint32_t cdQ = 256;
  uint32_t kvuse[2] = {0, 0}, aphase = 0, bphase = 0;

  if (tid == 0) {
    mbar_arrive_expect_tx(&bars[0], 2 * TB);
    tma_tile<D>(sQ, &map_q, &bars[0], colq, rowb + r0);
    tma_tile<D>(sdO, &map_do, &bars[0], colq, rowb + r0);
    mbar_arrive_expect_tx(&bars[1], 2 * TB);
    tma_tile<D>(sK, &map_k, &bars[1], colq, rowb);
    tma_tile<D>(sV, &map_v, &bars[1], colq, rowb);
  }
  float m, inv, delta;
  row_stats_delta<D>(stats, o, ldo, d_o, lddo, b, h, heads, seq, row, m, inv, delta);

  for (int j = 0; j < nblk; ++j) {
    const int buf = (NB == 2) ? (j & 1) : 0;
    load_key_flags(sFlag + (j & 1) * 128, key_mask, b, seq, j * 128, tid);
    if (tid == 0) {
      if (NB == 2 && j + 1 < nblk) {
        mbar_arrive_expect_tx(&bars[1 + (buf ^ 1)], 2 * TB);
        tma_tile<D>(sK + (buf ^ 1) * TB, &map_k, &bars[1 + (buf ^ 1)], colq, rowb + (j + 1) * 128);
        tma_tile<D>(sV + (buf ^ 1) * TB, &map_v, &bars[1 + (buf ^ 1)], colq, rowb + (j + 1) * 128);
      }
      if (j == 0) mbar_wait(&bars[0], 0);
      mbar_wait(&bars[1 + buf], kvuse[buf] & 1); kvuse[buf]++;
      tc_fence_after();
      mma_qk<D>(tmem_base, smem_u32(sQ), smem_u32(sK + buf * TB));           // S
      mma_qk<D>(tmem_base + 128, smem_u32(sdO), smem_u32(sV + buf * TB));    // dP = dO V^T
      umma_commit(&bars[3]);
    }
    const bool masked = __syncthreads_or(sFlag[(j & 1) * 128 + tid] != 0) || (causal && j == qt) || (r0 + 128 > seq);
    mbar_wait(&bars[3], aphase & 1); aphase++;
    tc_fence_after();
    if (masked)
      softmax_grad_block<false, true>(lane_addr, sFlag + (j & 1) * 128, j * 128, row, row_ok, causal, scale, m, inv, delta, 0,
                                      smem_u32(sdS), tid);
    else
      softmax_grad_block<false, false>(lane_addr, sFlag + (j & 1) * 128, j * 128, row, row_ok, causal, scale, m, inv, delta,
                                       0, smem_u32(sdS), tid);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      mma_pv<D>(tmem_base + cdQ, smem_u32(sdS), smem_u32(sK + buf * TB), j != 0);   // dQ += dS K_j
      umma_commit(&bars[4]);
      if (NB == 1 && j + 1 < nblk) {   // single buffer: reload K / V once this block's MMAs have drained
        mbar_wait(&bars[4], bphase & 1);
        mbar_arrive_expect_tx(&bars[1], 2 * TB);
        tma_tile<D>(sK, &map_k, &bars[1], colq, rowb + (j + 1) * 128);
        tma_tile<D>(sV, &map_v, &bars[1], colq, rowb + (j + 1) * 128);
      }
    }
    __syncwarp();
    mbar_wait(&bars[4], bphase & 1); bphase++;
    tc_fence_after();
  }
#pragma unroll
  for (int c = 0; c < D / 32; ++c) {
    uint32_t r[32];
    tmem_ld_32x32(lane_addr + cdQ + c * 32, r);
    tmem_ld_wait();
    if (row_ok) store_row_bf16(dq + ((int64_t)rowb + row) * lddq + colq + c * 32, r, scale);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc_dyn(tmem_base, tmem_cols);
}

// ------------------------------------------------------------------------------------------------ backward: dK, dV
template <int D>
__global__ void __launch_bounds__(128)
sattn_bwd_dkv_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_do,
                     const __grid_constant__ CUtensorMap map_k, const __grid_constant__ CUtensorMap map_v,
                     const uint8_t* __restrict__ key_mask, const __nv_bfloat16* __restrict__ o, int64_t ldo,
                     const __nv_bfloat16* __restrict__ d_o, int64_t lddo, const float* __restrict__ stats,
                     __nv_bfloat16* __restrict__ dk, int64_t lddk, __nv_bfloat16* __restrict__ dv, int64_t lddv,
                     int seq, int heads, float scale, int causal, uint32_t tmem_cols) {
  constexpr int TB = (D / 64) * 16384;
  constexpr int NB = (D == 64) ? 2 : 1;   // Q / dO buffers
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sK = smem;
  uint8_t* sV = sK + TB;
  uint8_t* sQ = sV + TB;
  uint8_t* sdO = sQ + NB * TB;
  uint8_t* sP = sdO + NB * TB;     // 2 slabs
  uint8_t* sdS = sP + 32768;       // 2 slabs
  uint8_t* sFlag = sdS + 32768;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sFlag + 128);   // kv, q0, q1, a, b
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 5);

  const int ntiles = (seq + 127) / 128;
  const int kb = blockIdx.x;   // key block; early key blocks see the most query tiles and come first
  const int h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, tid = threadIdx.x;
  const int i0 = causal ? kb : 0;
  const int colq = h * D, rowb = b * seq;

  if (tid == 0) {
    tma_prefetch_desc(&map_q); tma_prefetch_desc(&map_do); tma_prefetch_desc(&map_k); tma_prefetch_desc(&map_v);
    for (int i = 0; i < 5; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_dyn(tmem_ptr, tmem_cols);
  load_key_flags(sFlag, key_mask, b, seq, kb * 128, tid);
  tc_fence_before();
  const bool blk_flagged = __syncthreads_or(sFlag[tid] != 0);
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
  const uint32_t cdV = 256, cdK = 256 + D;
  uint32_t quse[2] = {0, 0}, aphase = 0, bphase = 0;

  if (tid == 0) {
    mbar_arrive_expect_tx(&bars[0], 2 * TB);
    tma_tile<D>(sK, &map_k, &bars[0], colq, rowb + kb * 128);
    tma_tile<D>(sV, &map_v, &bars[0], colq, rowb + kb * 128);
    mbar_arrive_expect_tx(&bars[1], 2 * TB);
    tma_tile<D>(sQ, &map_q, &bars[1], colq, rowb + i0 * 128);
    tma_tile<D>(sdO, &map_do, &bars[1], colq, rowb + i0 * 128);
  }
  for (int i = i0; i < ntiles; ++i) {
    const int it = i - i0;
    const int buf = (NB == 2) ? (it & 1) : 0;
    const int row = i * 128 + tid;
    const bool row_ok = row < seq;
    float m, inv, delta;
    row_stats_delta<D>(stats, o, ldo, d_o, lddo, b, h, heads, seq, row, m, inv, delta);
    if (tid == 0) {
      if (NB == 2 && i + 1 < ntiles) {
        mbar_arrive_expect_tx(&bars[1 + (buf ^ 1)], 2 * TB);
        tma_tile<D>(sQ + (buf ^ 1) * TB, &map_q, &bars[1 + (buf ^ 1)], colq, rowb + (i + 1) * 128);
        tma_tile<D>(sdO + (buf ^ 1) * TB, &map_do, &bars[1 + (buf ^ 1)], colq, rowb + (i + 1) * 128);
      }
      if (it == 0) mbar_wait(&bars[0], 0);
      mbar_wait(&bars[1 + buf], quse[buf] & 1); quse[buf]++;
      tc_fence_after();
      mma_qk<D>(tmem_base, smem_u32(sQ + buf * TB), smem_u32(sK));          // S  = Q_i K_j^T
      mma_qk<D>(tmem_base + 128, smem_u32(sdO + buf * TB), smem_u32(sV));   // dP = dO_i V_j^T
      umma_commit(&bars[3]);
    }
    __syncwarp();
    mbar_wait(&bars[3], aphase & 1); aphase++;
    tc_fence_after();
    if (blk_flagged || (causal && i == kb) || (i * 128 + 128 > seq))
      softmax_grad_block<true, true>(lane_addr, sFlag, kb * 128, row, row_ok, causal, scale, m, inv, delta, smem_u32(sP),
                                     smem_u32(sdS), tid);
    else
      softmax_grad_block<true, false>(lane_addr, sFlag, kb * 128, row, row_ok, causal, scale, m, inv, delta, smem_u32(sP),
                                      smem_u32(sdS), tid);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      mma_tn<D>(tmem_base + cdV, smem_u32(sP), smem_u32(sdO + buf * TB), it != 0);    // dV_j += P^T dO_i
      mma_tn<D>(tmem_base + cdK, smem_u32(sdS), smem_u32(sQ + buf * TB), it != 0);    // dK_j += dS^T Q_i
      umma_commit(&bars[4]);
      if (NB == 1 && i + 1 < ntiles) {
        mbar_wait(&bars[4], bphase & 1);
        mbar_arrive_expect_tx(&bars[1], 2 * TB);
        tma_tile<D>(sQ, &map_q, &bars[1], colq, rowb + (i + 1) * 128);
        tma_tile<D>(sdO, &map_do, &bars[1], colq, rowb + (i + 1) * 128);
      }
    }
    __syncwarp();
    mbar_wait(&bars[4], bphase & 1); bphase++;
    tc_fence_after();
  }
  const int key = kb * 128 + tid;
#pragma unroll
  for (int c = 0; c < D / 32; ++c) {
    uint32_t r[32];
    tmem_ld_32x32(lane_addr + cdV + c * 32, r);
    tmem_ld_wait();
    if (key < seq) store_row_bf16(dv + ((int64_t)rowb + key) * lddv + colq + c * 32, r, 1.f);
    tmem_ld_32x32(lane_addr + cdK + c * 32, r);
    tmem_ld_wait();
    if (key < seq) store_row_bf16(dk + ((int64_t)rowb + key) * lddk + colq + c * 32, r, scale);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc_dyn(tmem_base, tmem_cols);
}

struct Maps { CUtensorMap q, k, v, d_o; };

int build_maps(Maps& mp, const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, const void* d_o,
               int64_t lddo, int64_t batch, int64_t seq, int64_t heads, int d) {
  int rc;
  const uint64_t cols = (uint64_t)(heads * d), rows = (uint64_t)(batch * seq);
  if ((rc = make_tensor_map_2d(&mp.q, q, cols, rows, (uint64_t)ldq, 64, 128))) return rc;
  if ((rc = make_tensor_map_2d(&mp.k, k, cols, rows, (uint64_t)ldk, 64, 128))) return rc;
  if ((rc = make_tensor_map_2d(&mp.v, v, cols, rows, (uint64_t)ldv, 64, 128))) return rc;
  if (d_o != nullptr && (rc = make_tensor_map_2d(&mp.d_o, d_o, cols, rows, (uint64_t)lddo, 64, 128))) return rc;
  return 0;
}

template <int D>
int launch_fwd(const Maps& mp, const uint8_t* key_mask, void* o, int64_t ldo, float* stats, int64_t batch, int64_t seq,
               int64_t heads, float scale, int causal, cudaStream_t stream) {
  constexpr int TB = (D / 64) * 16384;
  const size_t smem = 5 * (size_t)TB + 32768 + 256 + 64;
  const uint32_t tmem_cols = 256;   // S 128 + O D
  auto kern = sattn_fwd_kernel<D>;
  MMGL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)((seq + 127) / 128), (unsigned)heads, (unsigned)batch);
  kern<<<grid, 128, smem, stream>>>(mp.q, mp.k, mp.v, key_mask, (__nv_bfloat16*)o, ldo, stats, (int)seq, (int)heads, scale,
                                    causal, tmem_cols);
  return check_launch("mmgl_sattn_fwd");
}

template <int D>
int launch_bwd(const Maps& mp, const uint8_t* key_mask, const void* o, int64_t ldo, const void* d_o, int64_t lddo,
               const float* stats, void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv, int64_t batch,
               int64_t seq, int64_t heads, float scale, int causal, cudaStream_t stream) {
  constexpr int TB = (D / 64) * 16384;
  constexpr int NB = (D == 64) ? 2 : 1;
  dim3 grid((unsigned)((seq + 127) / 128), (unsigned)heads, (unsigned)batch);
  {
    const size_t smem = (2 + 2 * NB) * (size_t)TB + 32768 + 256 + 64;
    auto kern = sattn_bwd_dq_kernel<D>;
    MMGL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, 128, smem, stream>>>(mp.q, mp.d_o, mp.k, mp.v, key_mask, (const __nv_bfloat16*)o, ldo,
                                      (const __nv_bfloat16*)d_o, lddo, stats, (__nv_bfloat16*)dq, lddq, (int)seq, (int)heads,
                                      scale, causal, 512u);
    if (int rc = check_launch("mmgl_sattn_bwd(dq)")) return rc;
  }
  {
    const size_t smem = (2 + 2 * NB) * (size_t)TB + 65536 + 128 + 64;
    auto kern = sattn_bwd_dkv_kernel<D>;
    MMGL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, 128, smem, stream>>>(mp.q, mp.d_o, mp.k, mp.v, key_mask, (const __nv_bfloat16*)o, ldo,
                                      (const __nv_bfloat16*)d_o, lddo, stats, (__nv_bfloat16*)dk, lddk, (__nv_bfloat16*)dv, lddv,
                                      (int)seq, (int)heads, scale, causal, 512u);
    return check_launch("mmgl_sattn_bwd(dkv)");
  }
}

int check_args(const char* who, int64_t batch, int64_t seq, int64_t heads, int64_t d) {
  MMGL_REQUIRE(batch > 0 && seq > 0 && heads > 0, "%s: empty problem", who);
  MMGL_REQUIRE(d == 64 || d == 128, "%s: head_dim must be 64 or 128 (got %lld)", who, (long long)d);
  MMGL_REQUIRE(batch < 65536 && heads < 65536, "%s: batch/heads too large for the grid", who);
  return 0;
}

}  // namespace
}  // namespace mmgl

using namespace mmgl;

extern "C" int mmgl_sattn_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                              const uint8_t* key_mask, void* o, int64_t ldo, float* stats, int64_t batch, int64_t seq,
                              int64_t heads, int64_t d, float scale, int32_t causal, void* stream_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  MMGL_BIND(q, "mmgl_sattn_fwd");
  if (int rc = check_args("mmgl_sattn_fwd", batch, seq, heads, d)) return rc;
  MMGL_REQUIRE(k && v && o && stats && aligned16(q) && aligned16(k) && aligned16(v) && aligned16(o) && ldq % 8 == 0 &&
               ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0, "mmgl_sattn_fwd: pointers must be 16B aligned, leading dims %% 8 == 0");
  Maps mp;
  if (int rc = build_maps(mp, q, ldq, k, ldk, v, ldv, nullptr, 0, batch, seq, heads, (int)d)) return rc;
  if (d == 64) return launch_fwd<64>(mp, key_mask, o, ldo, stats, batch, seq, heads, scale, causal, s);
  return launch_fwd<128>(mp, key_mask, o, ldo, stats, batch, seq, heads, scale, causal, s);
}

extern "C" int mmgl_sattn_bwd(const void* d_o, int64_t lddo, const void* q, int64_t ldq, const void* k, int64_t ldk,
                              const void* v, int64_t ldv, const void* o, int64_t ldo, const float* stats,
                              const uint8_t* key_mask, void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv,
                              int64_t lddv, int64_t batch, int64_t seq, int64_t heads, int64_t d, float scale,
                              int32_t causal, void* stream_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  MMGL_BIND(q, "mmgl_sattn_bwd");
  if (int rc = check_args("mmgl_sattn_bwd", batch, seq, heads, d)) return rc;
  MMGL_REQUIRE(d_o && k && v && o && stats && dq && dk && dv, "mmgl_sattn_bwd: null pointer");
  MMGL_REQUIRE(aligned16(d_o) && aligned16(q) && aligned16(k) && aligned16(v) && aligned16(o) && aligned16(dq) &&
               aligned16(dk) && aligned16(dv), "mmgl_sattn_bwd: pointers must be 16B aligned");
  MMGL_REQUIRE(lddo % 8 == 0 && ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0 && lddq % 8 == 0 &&
               lddk % 8 == 0 && lddv % 8 == 0, "mmgl_sattn_bwd: leading dims must be multiples of 8");
  Maps mp;
  if (int rc = build_maps(mp, q, ldq, k, ldk, v, ldv, d_o, lddo, batch, seq, heads, (int)d)) return rc;
  if (d == 64)
    return launch_bwd<64>(mp, key_mask, o, ldo, d_o, lddo, stats, dq, lddq, dk, lddk, dv, lddv, batch, seq, heads, scale, causal, s);
  return launch_bwd<128>(mp, key_mask, o, ldo, d_o, lddo, stats, dq, lddq, dk, lddk, dv, lddv, batch, seq, heads, scale, causal, s);
}
