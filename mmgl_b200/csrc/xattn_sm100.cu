// Fused cross-attention core, forward, on the 5th-generation tensor cores (tcgen05 + TMEM + TMA).
//
//   O = softmax(max(Q K^T + mask, -FLT_MAX)) V          per (sample, head); Q pre-scaled by d^-1/2
//
// One 128-query tile of one (sample, head) at a time (persistent CTAs walk runs of tiles, see the kernel):
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

#ifdef MMGL_TRACE
__device__ unsigned long long g_xtrace[1024];
#define XTR(idx_var, tag)                                                         \
  do {                                                                            \
    if (blockIdx.x == 0 && threadIdx.x == 0 && (idx_var) < 511) {                 \
      g_xtrace[2 * (idx_var)] = (unsigned long long)(tag);                        \
      g_xtrace[2 * (idx_var) + 1] = clock64();                                    \
      ++(idx_var);                                                                \
    }                                                                             \
  } while (0)
#else
#define XTR(idx_var, tag) do { } while (0)
#endif

constexpr float kLog2eF = 1.4426950408889634f;

// 16-byte chunk `chunk` (0..7) of row `row` inside a 128B-swizzled [rows][64 bf16] slab
__device__ __forceinline__ uint32_t swz128(uint32_t slab_base, int row, int chunk) {
  return slab_base + row * 128 + (((chunk ^ (row & 7)) & 7) << 4);
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// Persistent form (round 2): the grid is a few CTAs per SM and every CTA walks a CONTIGUOUS run of (sample, head, query tile)
// items, so the tensor-map fetch, barrier set-up, TMEM allocation and the mask row are paid once per CTA instead of once per
// tile, the head's K / V tiles stay in shared memory across its query tiles, the NEXT tile's Q is requested as soon as the
// S = Q K^T MMA has consumed the current one, and S of the next tile is issued right behind P V of this one (separate TMEM
// columns) so it completes while this tile's output is being written.
// Attend bits of sample b's bank row, one 32-key word per chunk: the byte mask is read once per (CTA, sample) with one
// coalesced load per chunk and a ballot.  (Round 1 kept a float mask row in
// shared memory behind a pointer whose address space the compiler could not see: 2 x Nk generic LD per thread and tile.)
// (kept in shared memory as nchunks words: one broadcast LDS per 32 keys)
__device__ __forceinline__ void load_mask_words(const uint8_t* __restrict__ mask, int b, int nk, int nchunks, uint32_t* saw) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int c = warp; c < nchunks; c += (blockDim.x >> 5)) {
    const int key = c * 32 + lane;
    const uint32_t w = __ballot_sync(0xffffffffu, key < nk && mask[(int64_t)b * nk + key] != 0);
    if (lane == 0) saw[c] = w;
  }
}
__device__ __forceinline__ uint32_t low_bits32(int n) { return n <= 0 ? 0u : (n >= 32 ? 0xffffffffu : ((1u << n) - 1u)); }

template <int D>
__global__ void __launch_bounds__(128)
xattn_fwd_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                    const __grid_constant__ CUtensorMap map_v, const uint8_t* __restrict__ mask,
                    __nv_bfloat16* __restrict__ o, int64_t ldo, float* __restrict__ stats, int seq, int nk, int nkp,
                    int heads, int batch, int items_per_cta, uint32_t tmem_cols) {
  constexpr int DS = D / 64;  // 64-wide (128-byte) slabs along the head dim
  extern __shared__ __align__(1024) uint8_t smem[];
  const int ps = (nkp + 63) / 64;                       // slabs of P along the key dim
  constexpr int NQ = 2;                                 // Q tiles in flight: the kernel is bound by HBM latency x bytes in flight
  uint8_t* sQ = smem;                                   // NQ x DS x [128][64]
  uint8_t* sK = sQ + NQ * DS * 16384;                   // DS x [nkp][64]   (K-major B operand of S = Q K^T)
  uint8_t* sV = sK + DS * nkp * 128;                    // DS x [nkp][64]   (MN-major B operand of O = P V)
  uint8_t* sP = sV + DS * nkp * 128;                    // ps x [128][64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + ps * 16384);     // mma1, mma2, kv, q[NQ]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 3 + NQ);
  uint32_t* saw = tmem_ptr + 2;                                      // [8] attend words of the current sample
  // barrier roles: bars[1] S done, bars[2] P V done, bars[0] K / V landed, bars[3 + slot] Q tile of that ring slot landed

  // lane-0 broadcast: warp-uniform for ptxas, so warp 0 can issue TMA / MMA from uniform registers (no ELECT / R2UR waterfall)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), tid = threadIdx.x;
  const int ntiles = (seq + 127) / 128;
  const int n_items = batch * heads * ntiles;
  const int item0 = blockIdx.x * items_per_cta;
  const int item_end = min(n_items, item0 + items_per_cta);
  if (item0 >= item_end) return;                        // (whole CTA, before any barrier)
  // item -> (sample, head, tile): tiles of one (sample, head) are consecutive
  auto bh_of = [&](int item) { return item / ntiles; };

  if (tid == 0) {
    tma_prefetch_desc(&map_q); tma_prefetch_desc(&map_k); tma_prefetch_desc(&map_v);
    for (int i = 0; i < 3 + NQ; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_dyn(tmem_ptr, tmem_cols);
  const int nchunks = (nkp + 31) / 32;
  pdl_launch();
  pdl_wait();   // barrier init / TMEM allocation overlapped the previous kernel's tail; global memory from here on
  load_mask_words(mask, bh_of(item0) / heads, nk, nchunks, saw);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);
  const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
  const uint32_t p_base = smem_u32(sP);

  // thread 0 helpers.  Tile `it` (local index) lives in Q ring slot it % NQ; its barrier completes once per use of the slot.
  auto load_q = [&](int item, int it) {
    const int bh = bh_of(item), t = item % ntiles, slot = it % NQ;
    mbar_arrive_expect_tx(&bars[3 + slot], DS * 16384);
#pragma unroll
    for (int j = 0; j < DS; ++j)
      tma_load_2d(sQ + (slot * DS + j) * 16384, &map_q, &bars[3 + slot], (bh % heads) * D + 64 * j, (bh / heads) * seq + t * 128);
  };
  auto load_kv = [&](int item) {
    const int bh = bh_of(item);
    mbar_arrive_expect_tx(&bars[0], 2 * DS * nkp * 128);
#pragma unroll
    for (int j = 0; j < DS; ++j) {
      tma_load_2d(sK + j * nkp * 128, &map_k, &bars[0], (bh % heads) * D + 64 * j, (bh / heads) * nk);
      tma_load_2d(sV + j * nkp * 128, &map_v, &bars[0], (bh % heads) * D + 64 * j, (bh / heads) * nk);
    }
  };
  // descriptor bases are built once; the issuing thread only adds (byte offset >> 4) per instruction
  const uint64_t desc_q = make_smem_desc(smem_u32(sQ), 16, 1024), desc_k = make_smem_desc(smem_u32(sK), 16, 1024);
  const uint64_t desc_p = make_smem_desc(p_base, 16, 1024), desc_v = make_smem_desc(smem_u32(sV), nkp * 128, 1024);
  const uint32_t idesc_s = make_idesc_bf16(128, nkp, 0, 0);
  const uint64_t kslab = (uint64_t)((nkp * 128) >> 4);       // next 64-column slab of the K tile
  auto issue_s = [&](int it) {   // S[128 x nkp] = Q K^T : A, B both K-major
    const uint64_t dq = desc_q + (uint64_t)(((it % NQ) * DS * 16384) >> 4);
#pragma unroll
    for (int k = 0; k < D / 16; ++k)
      umma_f16_ss(tmem_base, dq + (uint64_t)(((k >> 2) * 16384 + (k & 3) * 32) >> 4),
                  desc_k + (uint64_t)(k >> 2) * kslab + (uint64_t)(((k & 3) * 32) >> 4), idesc_s, k != 0 ? 1u : 0u);
    umma_commit(&bars[1]);
  };

  int xi = 0; (void)xi;
  XTR(xi, 1);
  uint32_t kv_uses = 0;       // warp 0: completed K / V loads (phase of bars[0])
  if (warp == 0) {            // the whole warp waits; one elected lane issues the TMA / MMA instructions
    if (elect_one()) {
      load_kv(item0);
      for (int a = 0; a < NQ && item0 + a < item_end; ++a) load_q(item0 + a, a);
    }
    __syncwarp();
    mbar_wait(&bars[0], kv_uses & 1); ++kv_uses;
    mbar_wait(&bars[3], 0);
    XTR(xi, 2);
    tc_fence_after();
    if (elect_one()) issue_s(0);
    __syncwarp();
  }

  for (int item = item0, it = 0; item < item_end; ++item, ++it) {
    const uint32_t par = it & 1;
    const int bh = bh_of(item), b = bh / heads, h = bh % heads, r0 = (item % ntiles) * 128;
    const bool has_next = item + 1 < item_end;
    const bool next_same_bh = has_next && bh_of(item + 1) == bh;

    // ---- softmax: this thread owns query row `tid` (TMEM lane tid)
    mbar_wait(&bars[1], par);
    XTR(xi, 100 + it * 10);
    tc_fence_after();
    if (warp == 0 && item + NQ < item_end) {   // S has consumed this slot: refill it NQ tiles ahead
      if (elect_one()) load_q(item + NQ, it + NQ);
      __syncwarp();
    }
    // pass 1: row maximum over the attended keys (-FLT_MAX when the row attends nothing: the clamp of the reference then
    // makes every existing key's score finfo.min, i.e. uniform attention over the bank)
    float mx = -FLT_MAX;
#pragma unroll 1
    for (int c = 0; c < nchunks; ++c) {
      uint32_t r[32];
      tmem_ld_32x32(lane_addr + c * 32, r);
      const uint32_t w = saw[c];
      tmem_ld_wait();
      float m4[4] = {-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX};
      if (w == 0xffffffffu) {      // warp-uniform: all 32 keys attended
#pragma unroll
        for (int j = 0; j < 32; ++j) m4[j & 3] = fmaxf(m4[j & 3], __uint_as_float(r[j]));
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) m4[j & 3] = fmaxf(m4[j & 3], ((w >> j) & 1u) ? __uint_as_float(r[j]) : -FLT_MAX);
      }
      mx = fmaxf(fmaxf(mx, fmaxf(m4[0], m4[1])), fmaxf(m4[2], m4[3]));
    }
    // pass 2: p = 2^((max(s, finfo.min) - mx) log2 e) on attended keys; masked existing keys: 1 when the row attends nothing,
    // else exactly 0 (exp of -3.4e38 - mx); zero-filled tail keys: 0
    const bool none = !(mx > -FLT_MAX);
    const float c1 = none ? 0.f : kLog2eF, off = none ? 0.f : mx * kLog2eF, pmask = none ? 1.f : 0.f;
    float sum = 0.f;
#pragma unroll 1
    for (int c = 0; c < nchunks; ++c) {
      uint32_t r[32];
      tmem_ld_32x32(lane_addr + c * 32, r);
      const uint32_t w = saw[c];
      const uint32_t ex = low_bits32(nk - c * 32);      // keys of the chunk that exist
      tmem_ld_wait();
      float p[32];
      float s4[4] = {0.f, 0.f, 0.f, 0.f};
      if (w == 0xffffffffu) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          p[j] = exp2f(fmaf(fmaxf(__uint_as_float(r[j]), -FLT_MAX), c1, -off));
          s4[j & 3] += p[j];
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float e = exp2f(fmaf(fmaxf(__uint_as_float(r[j]), -FLT_MAX), c1, -off));
          p[j] = ((w >> j) & 1u) ? e : (((ex >> j) & 1u) ? pmask : 0.f);
          s4[j & 3] += p[j];
        }
      }
      sum += (s4[0] + s4[1]) + (s4[2] + s4[3]);
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
    XTR(xi, 101 + it * 10);
    __syncthreads();       // P complete; every thread has finished reading S (and, from the previous tile, O)
    XTR(xi, 102 + it * 10);

    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {
        // O[128 x D] = P V : A = P K-major (K = keys), B = V MN-major ([key rows][d cols] as loaded)
        constexpr uint32_t idesc = make_idesc_bf16(128, D, 0, 1);
        for (int k4 = 0; k4 < nkp / 16; k4 += 4) {        // one 64-key slab of P per outer step
          const uint64_t da = desc_p + (uint64_t)(((k4 >> 2) * 16384) >> 4), db = desc_v + (uint64_t)((k4 * 2048) >> 4);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            if (k4 + kk < nkp / 16)
              umma_f16_ss(tmem_base + ps * 64, da + (uint64_t)((kk * 32) >> 4), db + (uint64_t)((kk * 2048) >> 4), idesc,
                          (k4 + kk) != 0 ? 1u : 0u);
        }
        umma_commit(&bars[2]);
      }
      __syncwarp();
      if (has_next) {
        if (!next_same_bh) {                   // new head: its K / V replace this one's once P V has read them
          mbar_wait(&bars[2], par);
          if (elect_one()) load_kv(item + 1);
          __syncwarp();
          mbar_wait(&bars[0], kv_uses & 1); ++kv_uses;
        }
        XTR(xi, 103 + it * 10);
        mbar_wait(&bars[3 + (it + 1) % NQ], ((it + 1) / NQ) & 1);
        XTR(xi, 104 + it * 10);
        tc_fence_after();
        if (elect_one()) issue_s(it + 1);      // runs behind P V on the tensor pipe, ready when the epilogue below is done
        __syncwarp();
      }
    }
    if (has_next && bh_of(item + 1) / heads != b) {    // next item belongs to another sample: its mask row
      load_mask_words(mask, bh_of(item + 1) / heads, nk, nchunks, saw);
      __syncthreads();
    }
    mbar_wait(&bars[2], par);
    XTR(xi, 105 + it * 10);
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
    XTR(xi, 106 + it * 10);
    tc_fence_before();       // the next tile's P V overwrites the O columns after the next block barrier
  }
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
  const size_t smem = 2 * (size_t)DS * 16384 + 2 * (size_t)DS * nkp * 128 + (size_t)ps * 16384 + 128;
  CUtensorMap mq, mk, mv;
  int rc;
  if ((rc = make_tensor_map_2d(&mq, q, (uint64_t)(heads * D), (uint64_t)(batch * seq), (uint64_t)ldq, 64, 128))) return rc;
  if ((rc = make_tensor_map_2d(&mk, k, (uint64_t)(heads * D), (uint64_t)(batch * nk), (uint64_t)ldk, 64, (uint32_t)nkp))) return rc;
  if ((rc = make_tensor_map_2d(&mv, v, (uint64_t)(heads * D), (uint64_t)(batch * nk), (uint64_t)ldv, 64, (uint32_t)nkp))) return rc;
  auto kern = xattn_fwd_tc_kernel<D>;
  MMGL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // persistent: as many CTAs as fit per SM (shared memory / TMEM columns), each walking a contiguous run of items; runs are
  // whole (sample, head)s when that costs at most one extra wave, so K / V are loaded once per head
  const int ntiles = (int)((seq + 127) / 128);
  const int64_t n_items = batch * heads * ntiles;
  int per_sm = (int)(512 / tmem_cols);
  const int by_smem = (int)((227 * 1024) / (smem + 1024));
  if (by_smem < per_sm) per_sm = by_smem;
  if (per_sm < 1) per_sm = 1;
  const int64_t slots = (int64_t)sm_count() * per_sm;
  int64_t per_cta = (n_items + slots - 1) / slots;
  if (per_cta < 1) per_cta = 1;
  if (per_cta < ntiles && ntiles <= 2 * per_cta) per_cta = ntiles;   // (longer runs simply straddle heads: K / V reload at the seam)
  const int64_t grid = (n_items + per_cta - 1) / per_cta;
  MMGL_CUDA(launch_pdl(kern, dim3((unsigned)grid), dim3(128), smem, stream, mq, mk, mv, mask, (__nv_bfloat16*)o, ldo, stats, (int)seq,
                       (int)nk, nkp, (int)heads, (int)batch, (int)per_cta, tmem_cols));
  return check_launch("mmgl_xattn_fwd");
}

#ifdef MMGL_TRACE
extern "C" int mmgl_debug_trace_xattn(unsigned long long* host_dst) {
  cudaDeviceSynchronize();
  return (int)cudaMemcpyFromSymbol(host_dst, g_xtrace, sizeof(g_xtrace));
}
#endif

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
  extern __shared__ __align__(1024) uint8_t smem[];
  const int ps = (nkp + 63) / 64;
  uint8_t* sK = smem;                         // DS x [nkp][64]
  uint8_t* sV = sK + DS * nkp * 128;          // DS x [nkp][64]
  uint8_t* sQ = sV + DS * nkp * 128;          // DS x [128][64]
  uint8_t* sdO = sQ + DS * 16384;             // DS x [128][64]
  uint8_t* sP = sdO + DS * 16384;             // ps x [128][64]
  uint8_t* sdS = sP + ps * 16384;             // ps x [128][64]  (+ one slab of slack: M = 128 reads a 2nd key atom)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sdS + (ps + 1) * 16384);   // kv, q, a, b
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 4);
  uint32_t* saw = tmem_ptr + 2;                                           // [8] attend words of this sample

  const int h = blockIdx.x, b = blockIdx.y;
  // lane-0 broadcast: warp-uniform for ptxas, so warp 0 can issue TMA / MMA from uniform registers (no ELECT / R2UR waterfall)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), tid = threadIdx.x;
  const int ntiles = (seq + 127) / 128;
  const int nchunks = (nkp + 31) / 32;

  if (tid == 0) {
    tma_prefetch_desc(&map_q); tma_prefetch_desc(&map_do); tma_prefetch_desc(&map_k); tma_prefetch_desc(&map_v);
    for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_dyn(tmem_ptr, tmem_cols);
  pdl_launch();
  pdl_wait();
  load_mask_words(mask, b, nk, nchunks, saw);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);
  const uint32_t cS = 0, cdP = ps * 64, cdQ = alias_dq ? 0 : 2 * ps * 64;
  const uint32_t cdV = (alias_dq ? 2 * ps * 64 : 2 * ps * 64 + D), cdK = cdV + D;
  const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
  const uint32_t aQ = smem_u32(sQ), adO = smem_u32(sdO), aK = smem_u32(sK), aV = smem_u32(sV), aP = smem_u32(sP),
                 adS = smem_u32(sdS);

  if (warp == 0) {
    if (elect_one()) {
      mbar_arrive_expect_tx(&bars[0], 2 * DS * nkp * 128);
#pragma unroll
      for (int j = 0; j < DS; ++j) {
        tma_load_2d(sK + j * nkp * 128, &map_k, &bars[0], h * D + 64 * j, b * nk);
        tma_load_2d(sV + j * nkp * 128, &map_v, &bars[0], h * D + 64 * j, b * nk);
      }
    }
    __syncwarp();
  }

  for (int t = 0; t < ntiles; ++t) {
    const int r0 = t * 128;
    const uint32_t par = t & 1;
    if (warp == 0) {
      if (elect_one()) {
        mbar_arrive_expect_tx(&bars[1], 2 * DS * 16384);
#pragma unroll
        for (int j = 0; j < DS; ++j) {
          tma_load_2d(sQ + j * 16384, &map_q, &bars[1], h * D + 64 * j, b * seq + r0);
          tma_load_2d(sdO + j * 16384, &map_do, &bars[1], h * D + 64 * j, b * seq + r0);
        }
      }
      __syncwarp();
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
    if (warp == 0) {
      if (t == 0) mbar_wait(&bars[0], 0);
      mbar_wait(&bars[1], par);
      tc_fence_after();
      if (elect_one()) {
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
    }
    mbar_wait(&bars[2], par);
    tc_fence_after();
    // P = 2^((max(s, finfo.min) - m) log2 e) / l on attended keys; masked existing keys: 1 / l when the row attends nothing,
    // else exactly 0; tail keys 0.  Masked entries tie in the reference's clamp max(S + mask, finfo.min): torch halves a
    // tie's gradient, hence the 1/2 on their dS.
    const bool none = !(m > -FLT_MAX);
    const float c1 = none ? 0.f : kLog2eF, off = none ? 0.f : m * kLog2eF, inv_ok = row_ok ? inv : 0.f;
    const float pmask = none ? inv_ok : 0.f;
#pragma unroll 1
    for (int c = 0; c < nchunks; ++c) {
      uint32_t rs[32], rp[32];
      tmem_ld_32x32(lane_addr + cS + c * 32, rs);
      tmem_ld_32x32(lane_addr + cdP + c * 32, rp);
      const uint32_t w = saw[c];
      const uint32_t ex = low_bits32(nk - c * 32);
      tmem_ld_wait();
      uint32_t pk[16], dk16[16];
      const bool full = w == 0xffffffffu;      // warp-uniform
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        float pv[2], dsv[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          float p = exp2f(fmaf(fmaxf(__uint_as_float(rs[j + e]), -FLT_MAX), c1, -off)) * inv_ok;
          float half = 1.f;
          if (!full) {
            const bool att = (w >> (j + e)) & 1u;
            p = att ? p : (((ex >> (j + e)) & 1u) ? pmask : 0.f);
            half = att ? 1.f : 0.5f;
          }
          pv[e] = p;
          dsv[e] = half * p * (__uint_as_float(rp[j + e]) - delta);
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
    if (warp == 0 && elect_one()) {
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
  const size_t smem = 2 * (size_t)DS * nkp * 128 + 2 * (size_t)DS * 16384 + (size_t)(2 * ps + 1) * 16384 + 96;
  CUtensorMap mq, mdo, mk, mv;
  int rc;
  if ((rc = make_tensor_map_2d(&mq, q, (uint64_t)(heads * D), (uint64_t)(batch * seq), (uint64_t)ldq, 64, 128))) return rc;
  if ((rc = make_tensor_map_2d(&mdo, d_o, (uint64_t)(heads * D), (uint64_t)(batch * seq), (uint64_t)lddo, 64, 128))) return rc;
  if ((rc = make_tensor_map_2d(&mk, k, (uint64_t)(heads * D), (uint64_t)(batch * nk), (uint64_t)ldk, 64, (uint32_t)nkp))) return rc;
  if ((rc = make_tensor_map_2d(&mv, v, (uint64_t)(heads * D), (uint64_t)(batch * nk), (uint64_t)ldv, 64, (uint32_t)nkp))) return rc;
  auto kern = xattn_bwd_tc_kernel<D>;
  MMGL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)heads, (unsigned)batch);
  MMGL_CUDA(launch_pdl(kern, grid, dim3(128), smem, stream, mq, mdo, mk, mv, (const __nv_bfloat16*)o, ldo, (const __nv_bfloat16*)d_o,
                       lddo, stats, mask, (__nv_bfloat16*)dq, lddq, (__nv_bfloat16*)dk, lddk, (__nv_bfloat16*)dv, lddv, (int)seq,
                       (int)nk, nkp, (int)heads, tmem_cols, alias_dq));
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
