// Self-attention of the decoder / encoder layers (SURVEY 8f row f1 and the concat path, row a7), tcgen05 + TMEM + TMA.
//
//   O = dropout(softmax(max(scale * Q K^T + rel_bias + causal + key-padding, finfo.min))) V      per (sample, head)
//
// References: MPTAttention self branch, model/modelling_cross_attention.py:201-275 with the additive mask built at
// :455-476 (causal AND key-not-padding); the HF T5 / OPT attention that model/modelling_self_attention.py:332 runs
// (T5: no 1/sqrt(d) scaling, additive relative-position bias, dropout on the probabilities; queries and keys of
// different lengths in the decoder's cross-attention).  Sequences here are short (S <= ~1200), so the kernels use
// 128 x 128 score blocks and a TWO-PASS softmax instead of online rescaling: pass 1 computes only the row maximum
// (S = Q K_j^T per key block), pass 2 recomputes S, exponentiates against the final maximum and accumulates
// O += P_j V_j in TMEM.  One extra QK^T per block buys the absence of any accumulator correction.
//
//   forward : CTA = TWO 128-query tiles of one (head, sample) in ping-pong: 16 softmax warps (8 per tile, two threads per
//             score row, each thread = one TMEM lane and half of the columns), two MMA-issuing warps (one per tile) and one
//             TMA warp.  The K / V blocks stream once through a 3-stage ring for both tiles.  Pass 1 double-buffers S in TMEM;
//             pass 2 works on 64-key half blocks with three S buffers (two at head_dim 128) and two P slabs per tile, so the
//             tensor core runs ahead of the exponentials.  Only blocks at or below the diagonal are visited when causal (tiles
//             are paired heavy-with-next-heavy, heavy pairs first).  A packed variable-length batch (cu_seqlens) is supported.
//   backward: dQ kernel   item = (query tile i): loops key blocks j, dQ_i += dS_ij K_j            (accumulates in TMEM)
//             dK/dV kernel item = (key block j): loops query tiles i, dV_j += P_ij^T dO_i, dK_j += dS_ij^T Q_i (TMEM)
//             both recompute S and dP = dO V^T from Q, K, V, dO and the saved row statistics (m, 1/l); 512 threads = four per
//             score row; persistent CTAs walk the items, heavy first.
// Masks are bit masks: one 32-bit word per 32 keys (attend = exists AND not padding), built once per CTA; the causal
// limit of the diagonal block is a per-thread shift.  Interior blocks (all 128 keys attended, not diagonal, no bias)
// take a 3-instruction-per-score path.
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

// Development trace (compiled in only with -DMMGL_TRACE): CTA (0,0,0) stamps (tag, clock) pairs per role into a global
// buffer read back by mmgl_debug_trace(); tools/attn_trace.py prints the timeline.
#ifdef MMGL_TRACE
__device__ unsigned long long g_trace[4 * 1024];
#define TR(role, idx_var, tag)                                                          \
  do {                                                                                  \
    if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (idx_var) < 511) {     \
      g_trace[(role) * 1024 + 2 * (idx_var)] = (unsigned long long)(tag);               \
      g_trace[(role) * 1024 + 2 * (idx_var) + 1] = clock64();                           \
      ++(idx_var);                                                                      \
    }                                                                                   \
  } while (0)
#else
#define TR(role, idx_var, tag) do { } while (0)
#endif

struct AttnParams {
  const uint8_t* key_mask;   // [B, seq_k] or null
  const float* rel_bias;     // [heads, seq_q + seq_k - 1] or null: bias(h, row, key) = rel_bias[h][key - row + seq_q - 1]
  int seq_q, seq_k, heads, causal;
  int batch;                 // backward kernels: persistent CTAs walk (tile, head, sample) work items
  float* d_rel_bias;         // dQ kernel only, or null: gradient of rel_bias, += over (sample, row) with atomics
  const int32_t* cu_seqlens; // forward only, or null: sample b = packed rows [cu[b], cu[b+1]) (variable-length batch)
  int coff;                  // causal: key allowed iff key <= row + coff, coff = seq_k - seq_q (bottom-right aligned; prefix K/V)
  float scale;
  uint32_t drop_thresh;      // 0: no dropout; else round(p * 65536)
  float drop_scale;          // 1 / (1 - p)
  uint64_t drop_seed;
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t swz(uint32_t slab_base, int row, int chunk) {
  return slab_base + row * 128 + (((chunk ^ (row & 7)) & 7) << 4);
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// the low n bits set (n may be <= 0 or >= 32)
__device__ __forceinline__ uint32_t low_bits(int n) { return n <= 0 ? 0u : (n >= 32 ? 0xffffffffu : ((1u << n) - 1u)); }

// attend bits of the whole key range of sample b: word w covers keys [32w, 32w + 32); bit = key exists and is not padding.
// All threads of the CTA must call (full warps).
__device__ __forceinline__ void build_key_bits(uint32_t* kbits, const uint8_t* key_mask, int b, int seq_k, int nwords) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int w = warp; w < nwords; w += nwarps) {
    const int key = w * 32 + lane;
    const bool a = key < seq_k && (key_mask == nullptr || key_mask[(int64_t)b * seq_k + key] != 0);
    const uint32_t bits = __ballot_sync(0xffffffffu, a);
    if (lane == 0) kbits[w] = bits;
  }
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
// keep bits (bit e = element e kept) of 32 consecutive keys starting at key0 (multiple of 32) of dropout row `drow`
__device__ __forceinline__ uint32_t keep_word(const AttnParams& p, int64_t drow, int key0, int64_t groups_per_row) {
  uint32_t w = 0;
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const DropBits bits = dropout_bits(p.drop_seed, drow, (key0 >> 3) + g, groups_per_row);
#pragma unroll
    for (int e = 0; e < 8; ++e) w |= dropout_keep(bits, e, p.drop_thresh) ? (1u << (8 * g + e)) : 0u;
  }
  return w;
}

// MMA issue from precomputed descriptor bases (descriptor of addr + delta = descriptor of addr + (delta >> 4)): the issuing
// thread adds compile-time constants instead of building two descriptors per instruction.
//   mma_qk_desc: S[128 x 128] = A[128 x D] * B[128 x D]^T, both K-major tiles of D / 64 slabs of [128][64]
//   mma_pv_desc: C[128 x D] (+)= A[128 x 128] * B[128 x D], A K-major (2 slabs), B a [128 rows][D] tile read MN-major (LBO 16384)
__host__ __device__ constexpr uint64_t kslab_off(int k) { return (uint64_t)(((k >> 2) * 16384 + (k & 3) * 32) >> 4); }
template <int D>
__device__ __forceinline__ void mma_qk_desc(uint32_t tmem, uint64_t da0, uint64_t db0) {
  const uint32_t idesc = make_idesc_bf16(128, 128, 0, 0);
#pragma unroll
  for (int k = 0; k < D / 16; ++k) umma_f16_ss(tmem, da0 + kslab_off(k), db0 + kslab_off(k), idesc, k != 0 ? 1u : 0u);
}
template <int D>
__device__ __forceinline__ void mma_pv_desc(uint32_t tmem, uint64_t da0, uint64_t db0, bool accumulate) {
  const uint32_t idesc = make_idesc_bf16(128, D, 0, 1);
#pragma unroll
  for (int k = 0; k < 8; ++k)
    umma_f16_ss(tmem, da0 + kslab_off(k), db0 + (uint64_t)((k * 2048) >> 4), idesc, (accumulate || k != 0) ? 1u : 0u);
}
// C[128 x D] (+)= A^T * B from descriptor bases: A a [128 rows][128] tile read MN-major (LBO 16384), B a [128 rows][D] tile
// read MN-major (LBO 16384); 8 k-steps of 16 rows (2048 bytes)
template <int D>
__device__ __forceinline__ void mma_tn_desc(uint32_t tmem, uint64_t da0, uint64_t db0, bool accumulate) {
  const uint32_t idesc = make_idesc_bf16(128, D, 1, 1);
#pragma unroll
  for (int k = 0; k < 8; ++k)
    umma_f16_ss(tmem, da0 + (uint64_t)((k * 2048) >> 4), db0 + (uint64_t)((k * 2048) >> 4), idesc, (accumulate || k != 0) ? 1u : 0u);
}
// half-block (64 keys) forms for the double-buffered pass 2 of the forward kernel:
//   S[128 x 64] = Q[128 x D] * Khalf[64 x D]^T          (db0 = descriptor of the half's first key row)
//   O[128 x D] (+)= P[128 x 64] * Vhalf[64 x D]          (da0 = P slab, db0 = descriptor of the half's first V row, MN-major)
template <int D>
__device__ __forceinline__ void mma_qk_half(uint32_t tmem, uint64_t da0, uint64_t db0) {
  const uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
#pragma unroll
  for (int k = 0; k < D / 16; ++k) umma_f16_ss(tmem, da0 + kslab_off(k), db0 + kslab_off(k), idesc, k != 0 ? 1u : 0u);
}
template <int D>
__device__ __forceinline__ void mma_pv_half(uint32_t tmem, uint64_t da0, uint64_t db0, bool accumulate) {
  const uint32_t idesc = make_idesc_bf16(128, D, 0, 1);
#pragma unroll
  for (int k = 0; k < 4; ++k)
    umma_f16_ss(tmem, da0 + (uint64_t)((k * 32) >> 4), db0 + (uint64_t)((k * 2048) >> 4), idesc, (accumulate || k != 0) ? 1u : 0u);
}
template <int D>
__device__ __forceinline__ void tma_tile(uint8_t* dst, const CUtensorMap* map, uint64_t* bar, int col0, int row0) {
#pragma unroll
  for (int j = 0; j < D / 64; ++j) tma_load_2d(dst + j * 16384, map, bar, col0 + 64 * j, row0);
}


// ---- per-chunk softmax pieces (one thread = one score row, rc = 32 consecutive keys of it) ------------------------------
// pass 1: running maxima.  raw_mx: unscaled, over attended keys (no-bias paths); mx: natural units (bias path).
// full = every lane of the warp attends all 32 keys.  bk points at the bias of this chunk's first key, valid up to [lim].
template <bool kBias>
__device__ __forceinline__ void max_chunk(const uint32_t (&rc)[32], uint32_t m, bool full, const float* bk, int lim,
                                          float scale, float& mx, float& raw_mx) {
  // four independent running maxima: with one or two warps per scheduler a single dependent chain of 32 is latency-bound
  float a[4] = {-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX};
  if (kBias) {
#pragma unroll
    for (int e = 0; e < 32; ++e) {
      const float x = fmaf(__uint_as_float(rc[e]), scale, __ldg(bk + min(e, lim)));
      a[e & 3] = fmaxf(a[e & 3], (m >> e) & 1u ? x : -FLT_MAX);
    }
    mx = fmaxf(fmaxf(mx, fmaxf(a[0], a[1])), fmaxf(a[2], a[3]));
  } else if (full) {
#pragma unroll
    for (int e = 0; e < 32; e += 2) a[(e >> 1) & 3] = fmaxf(a[(e >> 1) & 3], fmaxf(__uint_as_float(rc[e]), __uint_as_float(rc[e + 1])));
    raw_mx = fmaxf(fmaxf(raw_mx, fmaxf(a[0], a[1])), fmaxf(a[2], a[3]));
  } else {
#pragma unroll
    for (int e = 0; e < 32; ++e) a[e & 3] = fmaxf(a[e & 3], (m >> e) & 1u ? __uint_as_float(rc[e]) : -FLT_MAX);
    raw_mx = fmaxf(fmaxf(raw_mx, fmaxf(a[0], a[1])), fmaxf(a[2], a[3]));
  }
}
// pass 2: rc <- p = 2^(s * c1 + bias * bsc - mxc) on attended keys, 0 elsewhere; returns the chunk's sum.
// empty = no lane of the warp attends any of the 32 keys.
template <bool kBias>
__device__ __forceinline__ float exp_chunk(uint32_t (&rc)[32], uint32_t m, bool full, bool empty, const float* bk, int lim,
                                           float c1, float mxc, float bsc) {
  float sum[4] = {0.f, 0.f, 0.f, 0.f};
  if (empty) {
#pragma unroll
    for (int e = 0; e < 32; ++e) rc[e] = 0u;
  } else if (kBias) {
#pragma unroll
    for (int e = 0; e < 32; ++e) {
      float pe = ex2(fmaf(__uint_as_float(rc[e]), c1, fmaf(__ldg(bk + min(e, lim)), bsc, -mxc)));
      pe = (m >> e) & 1u ? pe : 0.f;
      sum[e & 3] += pe;
      rc[e] = __float_as_uint(pe);
    }
  } else if (full) {
#pragma unroll
    for (int e = 0; e < 32; ++e) {
      const float pe = ex2(fmaf(__uint_as_float(rc[e]), c1, -mxc));
      sum[e & 3] += pe;
      rc[e] = __float_as_uint(pe);
    }
  } else {
#pragma unroll
    for (int e = 0; e < 32; ++e) {
      float pe = ex2(fmaf(__uint_as_float(rc[e]), c1, -mxc));
      pe = (m >> e) & 1u ? pe : 0.f;
      sum[e & 3] += pe;
      rc[e] = __float_as_uint(pe);
    }
  }
  return (sum[0] + sum[1]) + (sum[2] + sum[3]);
}
// 32 fp32 values of one row -> bf16, into the 128B-swizzled [128][128] tile (2 slabs of 64 columns) at chunk c
__device__ __forceinline__ void store_chunk_bf16(uint32_t tile_base, int tid, int c, const uint32_t (&rc)[32]) {
  const uint32_t slab = tile_base + (c >> 1) * 16384;
#pragma unroll
  for (int g = 0; g < 4; ++g)
    sts128(swz(slab, tid, (c & 1) * 4 + g),
           pack_bf16(__uint_as_float(rc[8 * g]), __uint_as_float(rc[8 * g + 1])),
           pack_bf16(__uint_as_float(rc[8 * g + 2]), __uint_as_float(rc[8 * g + 3])),
           pack_bf16(__uint_as_float(rc[8 * g + 4]), __uint_as_float(rc[8 * g + 5])),
           pack_bf16(__uint_as_float(rc[8 * g + 6]), __uint_as_float(rc[8 * g + 7])));
}

// ------------------------------------------------------------------------------------------------ forward
// barrier indices of the forward kernel
template <int NS> struct FwdBars {
  static constexpr int qfull = 0;            // [2]  Q tile t landed
  static constexpr int kfull = 2;            // [NS] K stage landed
  static constexpr int kfree = 2 + NS;       // [NS] every MMA reading the K stage has completed
  static constexpr int vfull = 2 + 2 * NS;   // [NS]
  static constexpr int vfree = 2 + 3 * NS;   // [NS]
  static constexpr int sfull = 2 + 4 * NS;   // [2 buffers][2 tiles]  S_t = Q_t K_j^T complete in TMEM buffer (index 2 * buf + t)
  static constexpr int sfree = 6 + 4 * NS;   // [2 buffers][2 tiles]  pass 1: tile t's 128 threads have read that S buffer
  static constexpr int pfull = 10 + 4 * NS;  // [2 buffers][2 tiles]  pass 2: tile t's P half-block is in shared memory (128 arrivals)
  static constexpr int pfree = 14 + 4 * NS;  // [2 buffers][2 tiles]  P V of that half-block complete: P buffer reusable, O_t updated
  static constexpr int s2full = 18 + 4 * NS; // [3 buffers][2 tiles]  pass 2: S of a half block complete in TMEM buffer (index 2 * buf + t)
  static constexpr int count = 24 + 4 * NS;
};

template <int D, bool kBias, bool kDrop>
__global__ void __launch_bounds__(608, 1)
sattn_fwd_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                 const __grid_constant__ CUtensorMap map_v, const __grid_constant__ AttnParams p,
                 __nv_bfloat16* __restrict__ o, int64_t ldo, float* __restrict__ stats) {
  constexpr int TB = (D / 64) * 16384;   // bytes of one [128][D] tile
  constexpr int NS = (D == 64) ? 3 : 1;  // K / V ring stages
  using B = FwdBars<NS>;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;                    // 2 tiles
  uint8_t* sK = sQ + 2 * TB;             // NS stages
  uint8_t* sV = sK + NS * TB;            // NS stages
  uint8_t* sP = sV + NS * TB;            // 2 x [128][128] bf16 (2 slabs each)
  const int nbk_max = (p.seq_k + 127) / 128;                       // layout: sized for the longest sample
  uint32_t* kbits = reinterpret_cast<uint32_t*>(sP + 2 * 32768);   // 4 * nbk words
  uint64_t* bars = reinterpret_cast<uint64_t*>(kbits + 4 * nbk_max + ((4 * nbk_max) & 1));
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + B::count);
  float* xch = reinterpret_cast<float*>(tmem_ptr + 4);   // [2 tiles][2 column halves][128 rows]: row max, then row sum

  const int h = blockIdx.y, b = blockIdx.z;
  // this sample's lengths and first rows: uniform batch, or a packed variable-length batch (cu_seqlens)
  int sq = p.seq_q, sk = p.seq_k, rowq = b * p.seq_q, rowk = b * p.seq_k;
  if (p.cu_seqlens != nullptr) {
    const int c0 = p.cu_seqlens[b], c1 = p.cu_seqlens[b + 1];
    sq = sk = c1 - c0;
    rowq = rowk = c0;
  }
  const int nbk = (sk + 127) / 128;
  const int ntq = (sq + 127) / 128;
  if (ntq - 1 - 2 * (int)blockIdx.x < 0) return;   // a shorter sample has no such tile pair (whole CTA, before any barrier)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int tri = 0; (void)tri;
  if (threadIdx.x == 0) TR(0, tri, 1);
  // tile 0 of the pair is the later (heavier when causal) query tile; tile 1 the one before it (absent if < 0)
  const int qt0 = ntq - 1 - 2 * (int)blockIdx.x;
  const int qts[2] = {qt0, qt0 - 1};
  // key blocks a query tile visits: all of them, or (causal) those holding a key <= its last row + coff
  auto blocks_of = [&](int qt) { return p.causal ? min(nbk, (qt * 128 + 127 + p.coff) / 128 + 1) : nbk; };
  const int nblk[2] = {blocks_of(qt0), qt0 - 1 < 0 ? 0 : blocks_of(qt0 - 1)};
  const int nbmax = nblk[0];
  const int colq = h * D;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_q); tma_prefetch_desc(&map_k); tma_prefetch_desc(&map_v);
    for (int i = 0; i < B::count; ++i) {
      const bool wide = (i >= B::sfree && i < B::sfree + 4) || (i >= B::pfull && i < B::pfull + 4);
      const bool two = (i >= B::kfree && i < B::kfree + NS) || (i >= B::vfree && i < B::vfree + NS);
      mbar_init(&bars[i], wide ? 256 : (two ? 2 : 1));
    }
    fence_barrier_init();
  }
  if (warp == 16) tmem_alloc<512>(tmem_ptr);
  build_key_bits(kbits, p.key_mask, b, sk, 4 * nbk);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  if (threadIdx.x == 0) TR(0, tri, 2);

  if (warp == 18) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      for (int t = 0; t < 2; ++t)
        if (nblk[t] > 0) {
          mbar_arrive_expect_tx(&bars[B::qfull + t], TB);
          tma_tile<D>(sQ + t * TB, &map_q, &bars[B::qfull + t], colq, rowq + qts[t] * 128);
        }
      for (int n = 0; n < 2 * nbmax; ++n) {           // K block stream: pass 1 then pass 2
        const int j = n < nbmax ? n : n - nbmax, s = n % NS;
        if (n >= NS) mbar_wait(&bars[B::kfree + s], ((n / NS) - 1) & 1);
        mbar_arrive_expect_tx(&bars[B::kfull + s], TB);
        tma_tile<D>(sK + s * TB, &map_k, &bars[B::kfull + s], colq, rowk + j * 128);
        TR(3, tri, 100 + n);
        if (n >= nbmax) {                               // pass 2: V block j rides along
          const int sv = j % NS;
          if (j >= NS) mbar_wait(&bars[B::vfree + sv], ((j / NS) - 1) & 1);
          mbar_arrive_expect_tx(&bars[B::vfull + sv], TB);
          tma_tile<D>(sV + sv * TB, &map_v, &bars[B::vfull + sv], colq, rowk + j * 128);
        }
      }
    }
    __syncwarp();
  } else if (warp >= 16) {
    // ------------------------------------------------------------------ MMA issuers: warp 16 -> tile 0, warp 17 -> tile 1
    const int t = warp - 16;
    const int nb = nblk[t];
    if (lane == 0 && nb > 0) {
      const uint64_t dq = make_smem_desc(smem_u32(sQ + t * TB), 16, 1024);
      const uint64_t dp = make_smem_desc(smem_u32(sP + t * 32768), 16, 1024);
      const uint64_t dk0 = make_smem_desc(smem_u32(sK), 16, 1024);
      const uint64_t dv0 = make_smem_desc(smem_u32(sV), 16384, 1024);
      const uint32_t cS = tmem_base + t * 128;
      // a K / V stage is released by two arrivals, one per tile; the only tile using a block arrives twice
      auto release = [&](uint64_t* bar, int j) {
        umma_commit(bar);
        if (j >= nblk[1]) umma_commit(bar);
      };
      mbar_wait(&bars[B::qfull + t], 0);
      if (t == 0) TR(2, tri, 10);
      // pass 1: S_t(j) for the row maxima, double-buffered (block j at column offset (j & 1) * 256; the O columns are
      // not in use yet), so the tensor core runs one block ahead of the max reduction
      for (int j = 0; j < nb; ++j) {
        const int s = j % NS;
        mbar_wait(&bars[B::kfull + s], (j / NS) & 1);
        if (j >= 2) mbar_wait(&bars[B::sfree + 2 * (j & 1) + t], ((j >> 1) - 1) & 1);
        tc_fence_after();
        if (t == 0) TR(2, tri, 100 + j);
        mma_qk_desc<D>(cS + (j & 1) * 256, dq, dk0 + (uint64_t)((s * TB) >> 4));
        umma_commit(&bars[B::sfull + 2 * (j & 1) + t]);
        release(&bars[B::kfree + s], j);
        if (t == 0) TR(2, tri, 150 + j);
      }
      // pass 2 prologue: S_t(0) once BOTH tiles' pass-1 reads are done (tile 0's O columns alias tile 1's second S buffer
      // and vice versa); barrier phases are consumed in order
      for (int tt = 0; tt < 2; ++tt)
        for (int jj = max(nblk[tt] - 2, 0); jj < nblk[tt]; ++jj)
          mbar_wait(&bars[B::sfree + 2 * (jj & 1) + tt], (jj >> 1) & 1);
      // Pass 2 works on HALF blocks (64 keys) with NSB S buffers (TMEM) and two P slabs (shared memory) per tile: half block x
      // = half (x & 1) of key block x >> 1 lives in S buffer x % NSB and P slab x & 1.  S(jj + NSB) is issued right after
      // P V(jj), so the softmax warps always find the next scores ready and the hand-off latency is off the critical path.
      constexpr int NSB = (D == 64) ? 3 : 2;                 // S buffers per tile in pass 2 (TMEM: 2 * NSB * 64 + 2 * D <= 512)
      const uint32_t cS2 = tmem_base + t * (NSB * 64);       // + v * 64
      const uint32_t cO2 = tmem_base + 2 * NSB * 64 + t * D;
      const uint64_t dp2[2] = {dp, dp + (uint64_t)(16384 >> 4)};
      const int nsub = 2 * nb;
      // S of half block x (half x & 1 of key block x >> 1) into S buffer x % NSB; the first half waits for its K stage, the
      // second releases it
      auto issue_s = [&](int x) {
        const int jn = x >> 1, hk = x & 1, n = nbmax + jn, st = n % NS;
        if (hk == 0) mbar_wait(&bars[B::kfull + st], (n / NS) & 1);
        tc_fence_after();
        mma_qk_half<D>(cS2 + (x % NSB) * 64, dq, dk0 + (uint64_t)((st * TB + hk * 8192) >> 4));
        umma_commit(&bars[B::s2full + 2 * (x % NSB) + t]);
        if (hk == 1) release(&bars[B::kfree + st], jn);
      };
      if (t == 0) TR(2, tri, 198);
      for (int x = 0; x < NSB && x < nsub; ++x) issue_s(x);   // the tensor core starts NSB half blocks ahead of the softmax
      if (t == 0) TR(2, tri, 199);
      for (int jj = 0; jj < nsub; ++jj) {
        const int pp = jj & 1, j = jj >> 1, sv = j % NS;
        mbar_wait(&bars[B::pfull + 2 * pp + t], (jj >> 1) & 1);
        if (t == 0) TR(2, tri, 200 + jj);
        if (pp == 0) mbar_wait(&bars[B::vfull + sv], (j / NS) & 1);
        tc_fence_after();
        mma_pv_half<D>(cO2, dp2[pp], dv0 + (uint64_t)((sv * TB + pp * 8192) >> 4), jj != 0);
        umma_commit(&bars[B::pfree + 2 * pp + t]);
        if (pp == 1) release(&bars[B::vfree + sv], j);
        if (jj + NSB < nsub) issue_s(jj + NSB);               // into the S buffer the softmax warps have just finished reading
        if (t == 0) TR(2, tri, 300 + jj);
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ softmax warps: TWO threads per score row.
    // warp w: tile t = (w / 4) & 1, column half hf = w / 8, TMEM lane quarter w & 3; thread (row, hf) owns the 32-key chunks
    // 2 hf, 2 hf + 1 of every 128-key block in pass 1 and chunk hf of every 64-key half block in pass 2, so four warps per
    // scheduler hide each other's TMEM / MUFU latency.  Row max and row sum of the two halves meet in shared memory.
    const int t = (warp >> 2) & 1, hf = warp >> 3;
    const int nb = nblk[t];
    if (nb > 0) {
      const int qt = qts[t];
      const int tid = ((warp & 3) << 5) | lane;           // row in tile
      const int row = qt * 128 + tid;
      const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
      constexpr int NSB = (D == 64) ? 3 : 2;               // S buffers per tile in pass 2
      const uint32_t cS = lane_addr + t * 128;             // pass 1: + (j & 1) * 256
      const uint32_t cS2 = lane_addr + t * (NSB * 64);     // pass 2: + v * 64
      const uint32_t cO = lane_addr + 2 * NSB * 64 + t * D;
      const float scale = p.scale;
      constexpr bool has_bias = kBias;
      // bias of (row, key) = bias_row[key]  (bias_row points at the entry of key 0; negative offsets are valid memory)
      const float* bias_row = has_bias ? p.rel_bias + (int64_t)h * (sq + sk - 1) + (sq - 1 - min(row, sq - 1)) : nullptr;
      const float* bias0 = has_bias ? bias_row : nullptr;
      float* xrow = xch + (t * 2) * 128 + tid;            // [hf * 128] apart
      // ---------------- pass 1: row maximum (S double-buffered in TMEM: block j lives at column offset (j & 1) * 256)
      float mx = -FLT_MAX;     // natural units (bias path)
      float raw_mx = -FLT_MAX; // unscaled (paths without bias)
      for (int j = 0; j < nb; ++j) {
        const bool diag = p.causal && j * 128 + 127 > qt * 128 + p.coff;   // the block holds keys beyond some row's limit
        const uint32_t cSj = cS + (j & 1) * 256 + hf * 64;
        mbar_wait(&bars[B::sfull + 2 * (j & 1) + t], (j >> 1) & 1);
        tc_fence_after();
        if (threadIdx.x == 0) TR(0, tri, 100 + j);
        uint32_t ra[32], rb[32];
        tmem_ld_32x32(cSj, ra);
        tmem_ld_32x32(cSj + 32, rb);
#pragma unroll
        for (int w = 0; w < 2; ++w) {
          const int c = 2 * hf + w;
          const int key0 = j * 128 + c * 32;
          uint32_t m = kbits[4 * j + c];
          if (diag) m &= low_bits(row + p.coff - key0 + 1);
          const bool full = __all_sync(0xffffffffu, m == 0xffffffffu);
          const float* bk = has_bias ? bias0 + min(key0, sk - 1) : nullptr;
          const int lim = max(sk - 1 - key0, 0);
          if (w == 0) tmem_ld_wait();
          if (w == 0) max_chunk<kBias>(ra, m, full, bk, lim, scale, mx, raw_mx);
          else max_chunk<kBias>(rb, m, full, bk, lim, scale, mx, raw_mx);
        }
        tc_fence_before();
        mbar_arrive(&bars[B::sfree + 2 * (j & 1) + t]);
        if (threadIdx.x == 0) TR(0, tri, 150 + j);
      }
      if (raw_mx > -FLT_MAX) mx = fmaxf(mx, fmaxf(raw_mx * scale, -FLT_MAX));   // scale > 0
      // the two column halves of a row exchange their maxima (named barrier per tile: 256 threads)
      xrow[hf * 128] = mx;
      asm volatile("bar.sync %0, 256;" ::"r"(1 + t) : "memory");
      mx = fmaxf(mx, xrow[(hf ^ 1) * 128]);
      asm volatile("bar.sync %0, 256;" ::"r"(1 + t) : "memory");   // both read before the sums overwrite the slots
      // A row with no attended key at all (mx == -FLT_MAX): the reference's clamp makes every existing key's score
      // finfo.min, i.e. uniform attention -> p = 1 on the existing keys of the visited blocks.
      const bool none = !(mx > -FLT_MAX);
      const float c1 = none ? 0.f : scale * kL2E, mxc = none ? 0.f : mx * kL2E, bsc = none ? 0.f : kL2E;
      // ---------------- pass 2: P = exp(x - max), O += P V, on half blocks (64 keys = chunks 2u, 2u+1 of key block jj >> 1) with
      // NSB S buffers in TMEM (3 at head_dim 64) and two P slabs in shared memory: while this tile's warps exponentiate half
      // block jj, the tensor core already holds S(jj + 1) (and S(jj + 2)) and is free to run P V(jj - 1) and S(jj + NSB), so the
      // softmax -> MMA -> softmax hand-off latency stays off the critical path.  This thread owns chunk 2u + hf.
      float sum = 0.f;
      const uint32_t p_base = smem_u32(sP + t * 32768);
      const int64_t drow = ((int64_t)b * p.heads + h) * sq + row;
      const int64_t dgroups = (sk + 7) >> 3;
      const int nsub = 2 * nb;
      for (int jj = 0; jj < nsub; ++jj) {
        const int u = jj & 1, j = jj >> 1;
        const bool diag = p.causal && j * 128 + 127 > qt * 128 + p.coff;   // the block holds keys beyond some row's limit
        const int v = jj % NSB;                           // S buffer of this half block
        mbar_wait(&bars[B::s2full + 2 * v + t], (jj / NSB) & 1);
        tc_fence_after();
        if (threadIdx.x == 0) TR(0, tri, 200 + jj);
        uint32_t rc[32];
        tmem_ld_32x32(cS2 + v * 64 + hf * 32, rc);
        const int c = 2 * u + hf;                        // 32-key chunk of the 128-key block
        const int key0 = j * 128 + c * 32;
        uint32_t m = kbits[4 * j + c];
        if (diag) m &= low_bits(row + p.coff - key0 + 1);
        if (none) m = low_bits(sk - key0);
        const bool full = __all_sync(0xffffffffu, m == 0xffffffffu);
        const bool empty = __all_sync(0xffffffffu, m == 0u);
        const float* bk = has_bias ? bias0 + min(key0, sk - 1) : nullptr;
        const int lim = max(sk - 1 - key0, 0);
        tmem_ld_wait();
        sum += exp_chunk<kBias>(rc, m, full, empty, bk, lim, c1, mxc, bsc);
        if (kDrop) {
          const uint32_t keep = keep_word(p, drow, key0, dgroups);
#pragma unroll
          for (int e = 0; e < 32; ++e) rc[e] = (keep >> e) & 1u ? __float_as_uint(__uint_as_float(rc[e]) * p.drop_scale) : 0u;
        }
        if (jj >= 2) mbar_wait(&bars[B::pfree + 2 * u + t], ((jj >> 1) - 1) & 1);   // P V(jj - 2) done: buffer reusable
        // P buffer u is one 64-column slab; this thread's chunk fills its 16-byte columns 4 hf .. 4 hf + 3
        store_chunk_bf16(p_base + u * 16384, tid, hf, rc);
        fence_proxy_async();
        tc_fence_before();
        mbar_arrive(&bars[B::pfull + 2 * u + t]);
        if (threadIdx.x == 0) TR(0, tri, 250 + jj);
      }
      // the two halves of a row add their sums
      xrow[hf * 128] = sum;
      asm volatile("bar.sync %0, 256;" ::"r"(1 + t) : "memory");
      sum += xrow[(hf ^ 1) * 128];
      // all of O_t accumulated: the last P V of each buffer (completion is in issue order, so buffer 1's last implies all)
      mbar_wait(&bars[B::pfree + 0 + t], ((nsub - 2) >> 1) & 1);
      mbar_wait(&bars[B::pfree + 2 + t], ((nsub - 1) >> 1) & 1);
      tc_fence_after();
      if (threadIdx.x == 0) TR(0, tri, 900);
      const float inv = 1.f / sum;
      // O read-out: this thread stores the D / 2 columns [hf * D / 2, (hf + 1) * D / 2) of its row
#pragma unroll
      for (int c = 0; c < D / 64; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(cO + hf * (D / 2) + c * 32, r);
        tmem_ld_wait();
        if (row < sq) store_row_bf16(o + ((int64_t)rowq + row) * ldo + colq + hf * (D / 2) + c * 32, r, inv);
      }
      if (row < sq && stats != nullptr && hf == 0)
        *reinterpret_cast<float2*>(stats + (((int64_t)b * p.heads + h) * p.seq_q + row) * 2) = make_float2(mx, inv);
    }
  }
  if (threadIdx.x == 0) TR(0, tri, 990);
  tc_fence_before();
  __syncthreads();
  if (warp == 16) tmem_dealloc<512>(tmem_base);
  if (threadIdx.x == 0) TR(0, tri, 999);
}

// ------------------------------------------------------------------------------------------------ backward
// Both backward kernels run 512 threads per CTA: FOUR threads per score row (thread = (row, quarter); quarter = the
// 32-key chunk of every 128-key block it owns), so one block costs each thread a single 32-element chunk and 16 warps
// hide each other's TMEM / MUFU latency.  Warp w = 4 * quarter + row / 32 reads TMEM lanes 32 * (w & 3) as required.
//
// delta[row] = rowsum(dO . O): computed once by the dQ kernel (which sees every query row exactly once), parked in a
// caller-provided fp32 workspace [B, nh, seq_q] and read back by the dK/dV kernel (launched after it on the same stream).
template <int D>
__device__ __forceinline__ float delta_partial(const __nv_bfloat16* o, int64_t ldo, const __nv_bfloat16* d_o, int64_t lddo,
                                               int64_t grow, int col0) {
  // this thread's quarter of the row: D / 4 columns starting at col0
  const uint4* po = reinterpret_cast<const uint4*>(o + grow * ldo + col0);
  const uint4* pd = reinterpret_cast<const uint4*>(d_o + grow * lddo + col0);
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < D / 32; ++i) {
    const uint4 vo = __ldg(po + i), vd = __ldg(pd + i);
    const uint32_t wo[4] = {vo.x, vo.y, vo.z, vo.w}, wd[4] = {vd.x, vd.y, vd.z, vd.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) acc += bf16lo(wo[e]) * bf16lo(wd[e]) + bf16hi(wo[e]) * bf16hi(wd[e]);
  }
  return acc;
}

// P~ and dS of this thread's 32-key chunk c of one 128-key block: reads S (col 32c) and dP (col 128 + 32c) from TMEM,
// writes bf16 into the swizzled [128][128] tiles.
//   P~  = keep/(1-p) . P      (tile for dV += P~^T dO; equals P without dropout)
//   dS  = P . (keep/(1-p) . dP - delta)
// mw: attend bits of the chunk for this row (key bits AND causal limit; the existing keys when the row attends nothing).
template <bool kWriteP, bool kBias, bool kDrop>
__device__ __forceinline__ void softmax_grad_chunk(const AttnParams& p, uint32_t lane_addr, int c, uint32_t mw,
                                                   const float* bias_row, int key0, bool row_ok, float m, float inv,
                                                   float delta, int64_t drow, uint32_t p_base, uint32_t ds_base, int row_in_tile,
                                                   float* bias_bins = nullptr) {
  const bool flat = !(m > -FLT_MAX) || !row_ok;   // no attended key (uniform row) or a row beyond the sequence (p = 0)
  const float c1 = flat ? 0.f : p.scale * kL2E, mc = flat ? 0.f : m * kL2E, bsc = flat ? 0.f : kL2E;
  const float inv_ok = row_ok ? inv : 0.f;
  uint32_t rs[32], rp[32];
  tmem_ld_32x32(lane_addr + c * 32, rs);
  tmem_ld_32x32(lane_addr + 128 + c * 32, rp);
  uint32_t keep = 0xffffffffu;
  if (kDrop) keep = keep_word(p, drow, key0, (p.seq_k + 7) >> 3);
  const float ks = kDrop ? p.drop_scale : 1.f;
  const int lim = max(p.seq_k - 1 - key0, 0);
  const float* bk = kBias ? bias_row + min(key0, p.seq_k - 1) : nullptr;
  tmem_ld_wait();
  uint32_t pk[16], dk[16];
#pragma unroll
  for (int e = 0; e < 32; e += 2) {
    float pv[2], dv[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      float off = -mc;
      if (kBias) off = fmaf(__ldg(bk + min(e + u, lim)), bsc, -mc);
      float pr = ex2(fmaf(__uint_as_float(rs[e + u]), c1, off)) * inv_ok;
      pr = (mw >> (e + u)) & 1u ? pr : 0.f;
      if (kDrop) {
        const float kmul = (keep >> (e + u)) & 1u ? ks : 0.f;
        pv[u] = pr * kmul;
        dv[u] = pr * fmaf(__uint_as_float(rp[e + u]), kmul, -delta);
      } else {
        pv[u] = pr;
        dv[u] = pr * (__uint_as_float(rp[e + u]) - delta);
      }
    }
    pk[e >> 1] = pack_bf16(pv[0], pv[1]);
    dk[e >> 1] = pack_bf16(dv[0], dv[1]);
    if (kBias && bias_bins != nullptr) {   // d bias(key - row) += dS: one bin per diagonal of the 128 x 128 block
      atomicAdd(bias_bins + (c * 32 + e - row_in_tile + 127), dv[0]);
      atomicAdd(bias_bins + (c * 32 + e + 1 - row_in_tile + 127), dv[1]);
    }
  }
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    if (kWriteP) sts128(swz(p_base + (c >> 1) * 16384, row_in_tile, (c & 1) * 4 + g), pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]);
    sts128(swz(ds_base + (c >> 1) * 16384, row_in_tile, (c & 1) * 4 + g), dk[4 * g], dk[4 * g + 1], dk[4 * g + 2], dk[4 * g + 3]);
  }
}

// attend word of chunk c of key block j for query row (tile qt, row_in_tile)
__device__ __forceinline__ uint32_t row_word(const AttnParams& p, uint32_t kword, int j, int qt, int row_in_tile, int c, bool none) {
  uint32_t m = kword;
  if (p.causal) m &= low_bits(qt * 128 + row_in_tile + p.coff - (j * 128 + c * 32) + 1);
  if (none) m = low_bits(p.seq_k - (j * 128 + c * 32));   // uniform over the existing keys of the visited blocks
  return m;
}

// ------------------------------------------------------------------------------------------------ backward: dQ
template <int D, bool kBias, bool kDrop>
__global__ void __launch_bounds__(512, 1)
sattn_bwd_dq_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_do,
                    const __grid_constant__ CUtensorMap map_k, const __grid_constant__ CUtensorMap map_v,
                    const __grid_constant__ AttnParams p, const __nv_bfloat16* __restrict__ o, int64_t ldo,
                    const __nv_bfloat16* __restrict__ d_o, int64_t lddo, const float* __restrict__ stats,
                    __nv_bfloat16* __restrict__ dq, int64_t lddq, float* __restrict__ delta_ws) {
  constexpr int TB = (D / 64) * 16384;
  constexpr int NB = (D == 64) ? 2 : 1;   // K / V buffers
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sdO = sQ + TB;
  uint8_t* sK = sdO + TB;
  uint8_t* sV = sK + NB * TB;
  uint8_t* sdS = sV + NB * TB;     // 2 slabs
  const int nbk = (p.seq_k + 127) / 128;
  float* sDelta = reinterpret_cast<float*>(sdS + 32768);   // [4][128] partial row sums
  float* sBins = sDelta + 512;                             // [256] diagonal sums of dS (gradient of the relative-position bias)
  uint32_t* kbits = reinterpret_cast<uint32_t*>(sBins + 256);
  uint64_t* bars = reinterpret_cast<uint64_t*>(kbits + 4 * nbk);   // q, kv0, kv1, a, b
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 5);

  const int ntq = (p.seq_q + 127) / 128;
  const int warp = threadIdx.x >> 5, tid = threadIdx.x;
  const int rit = tid & 127, quarter = tid >> 7;          // row in tile, owned chunk

  if (tid == 0) {
    tma_prefetch_desc(&map_q); tma_prefetch_desc(&map_do); tma_prefetch_desc(&map_k); tma_prefetch_desc(&map_v);
    for (int i = 0; i < 5; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_ptr);
  const bool want_dbias = kBias && p.d_rel_bias != nullptr;
  if (tid < 256) sBins[tid] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  const uint32_t cdQ = 256;
  uint32_t kvuse[2] = {0, 0}, aphase = 0, bphase = 0, qphase = 0;
  // descriptor bases built once: the issuing thread only adds constants per MMA (it is also a worker: issue time is on the
  // critical path of every block)
  const uint64_t desc_q = make_smem_desc(smem_u32(sQ), 16, 1024), desc_do = make_smem_desc(smem_u32(sdO), 16, 1024);
  const uint64_t desc_k = make_smem_desc(smem_u32(sK), 16, 1024), desc_v = make_smem_desc(smem_u32(sV), 16, 1024);
  const uint64_t desc_ds = make_smem_desc(smem_u32(sdS), 16, 1024), desc_kmn = make_smem_desc(smem_u32(sK), 16384, 1024);

  // Persistent CTA: work items (query tile, head, sample), late (heavier when causal) tiles first, dealt round-robin so
  // every CTA gets a mix; TMEM, barriers and the tensor-map fetch are paid once per CTA instead of once per tile.
  const int hb = p.heads * p.batch;
  const int n_items = ntq * hb;
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int qt = ntq - 1 - item / hb;
    const int h = (item % hb) % p.heads, b = (item % hb) / p.heads;
    const int r0 = qt * 128, row = r0 + rit;
    const bool row_ok = row < p.seq_q;
    const int nblk = p.causal ? min(nbk, (qt * 128 + 127 + p.coff) / 128 + 1) : nbk;
    const int colq = h * D, rowq = b * p.seq_q, rowk = b * p.seq_k;

    if (tid == 0) {   // nothing of the previous item is still reading these buffers (its last MMA was waited for)
      mbar_arrive_expect_tx(&bars[0], 2 * TB);
      tma_tile<D>(sQ, &map_q, &bars[0], colq, rowq + r0);
      tma_tile<D>(sdO, &map_do, &bars[0], colq, rowq + r0);
      mbar_arrive_expect_tx(&bars[1], 2 * TB);
      tma_tile<D>(sK, &map_k, &bars[1], colq, rowk);
      tma_tile<D>(sV, &map_v, &bars[1], colq, rowk);
    }
    build_key_bits(kbits, p.key_mask, b, p.seq_k, 4 * nbk);
    sDelta[quarter * 128 + rit] = row_ok ? delta_partial<D>(o, ldo, d_o, lddo, (int64_t)rowq + row, colq + quarter * (D / 4)) : 0.f;
    __syncthreads();
    float m = 0.f, inv = 0.f;
    const float delta = (sDelta[rit] + sDelta[128 + rit]) + (sDelta[256 + rit] + sDelta[384 + rit]);
    if (row_ok) {
      const float2 st = *reinterpret_cast<const float2*>(stats + (((int64_t)b * p.heads + h) * p.seq_q + row) * 2);
      m = st.x; inv = st.y;
      if (quarter == 0) delta_ws[((int64_t)b * p.heads + h) * p.seq_q + row] = delta;
    }
    const bool none = !(m > -FLT_MAX);
    const float* bias_row = kBias ? p.rel_bias + (int64_t)h * (p.seq_q + p.seq_k - 1) + (p.seq_q - 1 - min(row, p.seq_q - 1)) : nullptr;
    const int64_t drow = ((int64_t)b * p.heads + h) * p.seq_q + row;

    for (int j = 0; j < nblk; ++j) {
      const int buf = (NB == 2) ? (j & 1) : 0;
      if (tid == 0) {
        if (j == 0) { mbar_wait(&bars[0], qphase & 1); }
        mbar_wait(&bars[1 + buf], kvuse[buf] & 1); kvuse[buf]++;
        tc_fence_after();
        const uint64_t off = (uint64_t)((buf * TB) >> 4);
        mma_qk_desc<D>(tmem_base, desc_q, desc_k + off);            // S
        mma_qk_desc<D>(tmem_base + 128, desc_do, desc_v + off);     // dP = dO V^T
        umma_commit(&bars[3]);
        if (NB == 2 && j + 1 < nblk) {   // the next block's tiles, after this block's MMAs are on their way
          mbar_arrive_expect_tx(&bars[1 + (buf ^ 1)], 2 * TB);
          tma_tile<D>(sK + (buf ^ 1) * TB, &map_k, &bars[1 + (buf ^ 1)], colq, rowk + (j + 1) * 128);
          tma_tile<D>(sV + (buf ^ 1) * TB, &map_v, &bars[1 + (buf ^ 1)], colq, rowk + (j + 1) * 128);
        }
      }
      const uint32_t mw = row_word(p, kbits[4 * j + quarter], j, qt, rit, quarter, none);
      __syncwarp();
      mbar_wait(&bars[3], aphase & 1); aphase++;
      tc_fence_after();
      softmax_grad_chunk<false, kBias, kDrop>(p, lane_addr, quarter, mw, bias_row, j * 128 + quarter * 32, row_ok, m, inv,
                                              delta, drow, 0, smem_u32(sdS), rit, want_dbias ? sBins : nullptr);
      fence_proxy_async();
      tc_fence_before();
      __syncthreads();
      if (want_dbias) {   // flush this block's 255 diagonals: bin i holds key - row = j * 128 - r0 + i - 127
        if (tid < 255) {
          const float v = sBins[tid];
          const int idx = j * 128 - r0 + tid - 127 + p.seq_q - 1;
          if (v != 0.f && idx >= 0 && idx < p.seq_q + p.seq_k - 1)
            atomicAdd(p.d_rel_bias + (int64_t)h * (p.seq_q + p.seq_k - 1) + idx, v);
          sBins[tid] = 0.f;
        }
        __syncthreads();
      }
      if (tid == 0) {
        tc_fence_after();
        mma_pv_desc<D>(tmem_base + cdQ, desc_ds, desc_kmn + (uint64_t)((buf * TB) >> 4), j != 0);   // dQ += dS K_j
        umma_commit(&bars[4]);
        if (NB == 1 && j + 1 < nblk) {   // single buffer: reload K / V once this block's MMAs have drained
          mbar_wait(&bars[4], bphase & 1);
          mbar_arrive_expect_tx(&bars[1], 2 * TB);
          tma_tile<D>(sK, &map_k, &bars[1], colq, rowk + (j + 1) * 128);
          tma_tile<D>(sV, &map_v, &bars[1], colq, rowk + (j + 1) * 128);
        }
      }
      __syncwarp();
      mbar_wait(&bars[4], bphase & 1); bphase++;
      tc_fence_after();
    }
    qphase++;
    // dQ tile: quarter q stores columns [q * D/4, (q + 1) * D/4)
    if (D == 64) {
      uint32_t r[16];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                     "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                   : "r"(lane_addr + cdQ + quarter * 16) : "memory");
      tmem_ld_wait();
      if (row_ok) {
        __nv_bfloat16* dst = dq + ((int64_t)rowq + row) * lddq + colq + quarter * 16;
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          uint4 v;
          v.x = pack_bf16(__uint_as_float(r[8 * g]) * p.scale, __uint_as_float(r[8 * g + 1]) * p.scale);
          v.y = pack_bf16(__uint_as_float(r[8 * g + 2]) * p.scale, __uint_as_float(r[8 * g + 3]) * p.scale);
          v.z = pack_bf16(__uint_as_float(r[8 * g + 4]) * p.scale, __uint_as_float(r[8 * g + 5]) * p.scale);
          v.w = pack_bf16(__uint_as_float(r[8 * g + 6]) * p.scale, __uint_as_float(r[8 * g + 7]) * p.scale);
          *reinterpret_cast<uint4*>(dst + 8 * g) = v;
        }
      }
    } else {
      uint32_t r[32];
      tmem_ld_32x32(lane_addr + cdQ + quarter * 32, r);
      tmem_ld_wait();
      if (row_ok) store_row_bf16(dq + ((int64_t)rowq + row) * lddq + colq + quarter * 32, r, p.scale);
    }
    tc_fence_before();   // the next item's S / dP / dQ MMAs overwrite TMEM this item's threads have just read
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// ------------------------------------------------------------------------------------------------ backward: dK, dV
template <int D, bool kBias, bool kDrop>
__global__ void __launch_bounds__(512, 1)
sattn_bwd_dkv_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_do,
                     const __grid_constant__ CUtensorMap map_k, const __grid_constant__ CUtensorMap map_v,
                     const __grid_constant__ AttnParams p, const float* __restrict__ stats,
                     const float* __restrict__ delta_ws, __nv_bfloat16* __restrict__ dk, int64_t lddk,
                     __nv_bfloat16* __restrict__ dv, int64_t lddv) {
  constexpr int TB = (D / 64) * 16384;
  constexpr int NB = (D == 64) ? 2 : 1;   // Q / dO buffers
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sK = smem;
  uint8_t* sV = sK + TB;
  uint8_t* sQ = sV + TB;
  uint8_t* sdO = sQ + NB * TB;
  uint8_t* sP = sdO + NB * TB;     // 2 slabs
  uint8_t* sdS = sP + 32768;       // 2 slabs
  uint32_t* kbits4 = reinterpret_cast<uint32_t*>(sdS + 32768);   // the 4 words of this key block
  uint64_t* bars = reinterpret_cast<uint64_t*>(kbits4 + 4);     // kv, q0, q1, a, b
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 5);

  const int ntq = (p.seq_q + 127) / 128;
  const int nbk = (p.seq_k + 127) / 128;
  const int warp = threadIdx.x >> 5, tid = threadIdx.x, lane = threadIdx.x & 31;
  const int rit = tid & 127, quarter = tid >> 7;

  if (tid == 0) {
    tma_prefetch_desc(&map_q); tma_prefetch_desc(&map_do); tma_prefetch_desc(&map_k); tma_prefetch_desc(&map_v);
    for (int i = 0; i < 5; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  const uint32_t cdV = 256, cdK = 256 + D;
  uint32_t quse[2] = {0, 0}, aphase = 0, bphase = 0, kphase = 0;
  const uint64_t desc_k = make_smem_desc(smem_u32(sK), 16, 1024), desc_v = make_smem_desc(smem_u32(sV), 16, 1024);
  const uint64_t desc_q = make_smem_desc(smem_u32(sQ), 16, 1024), desc_do = make_smem_desc(smem_u32(sdO), 16, 1024);
  const uint64_t desc_p_mn = make_smem_desc(smem_u32(sP), 16384, 1024), desc_ds_mn = make_smem_desc(smem_u32(sdS), 16384, 1024);
  const uint64_t desc_q_mn = make_smem_desc(smem_u32(sQ), 16384, 1024), desc_do_mn = make_smem_desc(smem_u32(sdO), 16384, 1024);

  // Persistent CTA over work items (key block, head, sample); early key blocks (seen by the most query tiles when
  // causal) first, dealt round-robin.
  const int hb = p.heads * p.batch;
  const int n_items = nbk * hb;
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
  const int kb = item / hb;
  const int h = (item % hb) % p.heads, b = (item % hb) / p.heads;
  const int i0 = p.causal ? max(0, (kb * 128 - p.coff) / 128) : 0;   // first query tile with a row that may see this key block
  const int colq = h * D, rowq = b * p.seq_q, rowk = b * p.seq_k;
  if (tid == 0) {
    mbar_arrive_expect_tx(&bars[0], 2 * TB);
    tma_tile<D>(sK, &map_k, &bars[0], colq, rowk + kb * 128);
    tma_tile<D>(sV, &map_v, &bars[0], colq, rowk + kb * 128);
    mbar_arrive_expect_tx(&bars[1], 2 * TB);
    tma_tile<D>(sQ, &map_q, &bars[1], colq, rowq + i0 * 128);
    tma_tile<D>(sdO, &map_do, &bars[1], colq, rowq + i0 * 128);
  }
  if (tid < 128) {
    const int key = kb * 128 + tid;
    const bool a = key < p.seq_k && (p.key_mask == nullptr || p.key_mask[(int64_t)b * p.seq_k + key] != 0);
    const uint32_t bits = __ballot_sync(0xffffffffu, a);
    if (lane == 0) kbits4[warp] = bits;
  }
  __syncthreads();
  const uint32_t kword = kbits4[quarter];
  for (int i = i0; i < ntq; ++i) {
    const int it = i - i0;
    const int buf = (NB == 2) ? (it & 1) : 0;
    const int row = i * 128 + rit;
    const bool row_ok = row < p.seq_q;
    float m = 0.f, inv = 0.f, delta = 0.f;
    if (row_ok) {
      const int64_t sidx = ((int64_t)b * p.heads + h) * p.seq_q + row;
      const float2 st = __ldg(reinterpret_cast<const float2*>(stats + sidx * 2));
      m = st.x; inv = st.y;
      delta = __ldg(delta_ws + sidx);
    }
    if (tid == 0) {
      if (it == 0) mbar_wait(&bars[0], kphase & 1);
      mbar_wait(&bars[1 + buf], quse[buf] & 1); quse[buf]++;
      tc_fence_after();
      const uint64_t off = (uint64_t)((buf * TB) >> 4);
      mma_qk_desc<D>(tmem_base, desc_q + off, desc_k);             // S  = Q_i K_j^T
      mma_qk_desc<D>(tmem_base + 128, desc_do + off, desc_v);      // dP = dO_i V_j^T
      umma_commit(&bars[3]);
      if (NB == 2 && i + 1 < ntq) {   // the next query tile, after this tile's MMAs are on their way
        mbar_arrive_expect_tx(&bars[1 + (buf ^ 1)], 2 * TB);
        tma_tile<D>(sQ + (buf ^ 1) * TB, &map_q, &bars[1 + (buf ^ 1)], colq, rowq + (i + 1) * 128);
        tma_tile<D>(sdO + (buf ^ 1) * TB, &map_do, &bars[1 + (buf ^ 1)], colq, rowq + (i + 1) * 128);
      }
    }
    const bool none = !(m > -FLT_MAX);
    const uint32_t mw = row_word(p, kword, kb, i, rit, quarter, none);
    const float* bias_row = kBias ? p.rel_bias + (int64_t)h * (p.seq_q + p.seq_k - 1) + (p.seq_q - 1 - min(row, p.seq_q - 1)) : nullptr;
    const int64_t drow = ((int64_t)b * p.heads + h) * p.seq_q + row;
    __syncwarp();
    mbar_wait(&bars[3], aphase & 1); aphase++;
    tc_fence_after();
    softmax_grad_chunk<true, kBias, kDrop>(p, lane_addr, quarter, mw, bias_row, kb * 128 + quarter * 32, row_ok, m, inv,
                                           delta, drow, smem_u32(sP), smem_u32(sdS), rit);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint64_t off = (uint64_t)((buf * TB) >> 4);
      mma_tn_desc<D>(tmem_base + cdV, desc_p_mn, desc_do_mn + off, it != 0);    // dV_j += P~^T dO_i
      mma_tn_desc<D>(tmem_base + cdK, desc_ds_mn, desc_q_mn + off, it != 0);    // dK_j += dS^T Q_i
      umma_commit(&bars[4]);
      if (NB == 1 && i + 1 < ntq) {
        mbar_wait(&bars[4], bphase & 1);
        mbar_arrive_expect_tx(&bars[1], 2 * TB);
        tma_tile<D>(sQ, &map_q, &bars[1], colq, rowq + (i + 1) * 128);
        tma_tile<D>(sdO, &map_do, &bars[1], colq, rowq + (i + 1) * 128);
      }
    }
    __syncwarp();
    mbar_wait(&bars[4], bphase & 1); bphase++;
    tc_fence_after();
  }
  // quarters 0,1 store dV, quarters 2,3 store dK; each stores half of the D columns of its key row
  {
    const int key = kb * 128 + rit;
    const bool is_k = quarter >= 2;
    const int half = quarter & 1;
    const uint32_t src = lane_addr + (is_k ? cdK : cdV) + half * (D / 2);
    __nv_bfloat16* dst = (is_k ? dk + ((int64_t)rowk + key) * lddk : dv + ((int64_t)rowk + key) * lddv) + colq + half * (D / 2);
    const float mul = is_k ? p.scale : 1.f;
#pragma unroll
    for (int c = 0; c < D / 64; ++c) {
      uint32_t r[32];
      tmem_ld_32x32(src + c * 32, r);
      tmem_ld_wait();
      if (key < p.seq_k) store_row_bf16(dst + c * 32, r, mul);
    }
  }
  kphase++;
  tc_fence_before();   // the next item's MMAs overwrite TMEM this item's threads have just read
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

struct Maps { CUtensorMap q, k, v, d_o; };

int build_maps(Maps& mp, const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, const void* d_o,
               int64_t lddo, int64_t batch, int64_t seq_q, int64_t seq_k, int64_t heads, int d, int64_t total_tokens = 0) {
  int rc;
  const uint64_t cols = (uint64_t)(heads * d);
  const uint64_t rows_q = (uint64_t)(total_tokens > 0 ? total_tokens : batch * seq_q);
  const uint64_t rows_k = (uint64_t)(total_tokens > 0 ? total_tokens : batch * seq_k);
  if ((rc = make_tensor_map_2d(&mp.q, q, cols, rows_q, (uint64_t)ldq, 64, 128))) return rc;
  if ((rc = make_tensor_map_2d(&mp.k, k, cols, rows_k, (uint64_t)ldk, 64, 128))) return rc;
  if ((rc = make_tensor_map_2d(&mp.v, v, cols, rows_k, (uint64_t)ldv, 64, 128))) return rc;
  if (d_o != nullptr && (rc = make_tensor_map_2d(&mp.d_o, d_o, cols, rows_q, (uint64_t)lddo, 64, 128))) return rc;
  return 0;
}

size_t kbits_bytes(int64_t seq_k) {
  const size_t words = 4 * (size_t)((seq_k + 127) / 128);
  return (words + (words & 1)) * 4;
}

template <int D>
int launch_fwd(const Maps& mp, const AttnParams& p, void* o, int64_t ldo, float* stats, int64_t batch, cudaStream_t stream) {
  constexpr int TB = (D / 64) * 16384;
  constexpr int NS = (D == 64) ? 3 : 1;
  const size_t smem = (2 + 2 * NS) * (size_t)TB + 65536 + kbits_bytes(p.seq_k) + FwdBars<NS>::count * 8 + 16 + 2 * 2 * 128 * 4;
  const bool bias = p.rel_bias != nullptr, drop = p.drop_thresh != 0;
  auto kern = bias ? (drop ? sattn_fwd_kernel<D, true, true> : sattn_fwd_kernel<D, true, false>)
                   : (drop ? sattn_fwd_kernel<D, false, true> : sattn_fwd_kernel<D, false, false>);
  MMGL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int ntq = (p.seq_q + 127) / 128;
  dim3 grid((unsigned)((ntq + 1) / 2), (unsigned)p.heads, (unsigned)batch);
  kern<<<grid, 608, smem, stream>>>(mp.q, mp.k, mp.v, p, (__nv_bfloat16*)o, ldo, stats);
  return check_launch("mmgl_attn_fwd");
}

template <int D, bool kBias, bool kDrop>
int launch_bwd_v(const Maps& mp, const AttnParams& p, const void* o, int64_t ldo, const void* d_o, int64_t lddo,
                 const float* stats, void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv, float* delta_ws,
                 int64_t batch, cudaStream_t stream) {
  constexpr int TB = (D / 64) * 16384;
  constexpr int NB = (D == 64) ? 2 : 1;
  {
    const size_t smem = (2 + 2 * NB) * (size_t)TB + 32768 + 2048 + 1024 + kbits_bytes(p.seq_k) + 5 * 8 + 16;
    auto kern = sattn_bwd_dq_kernel<D, kBias, kDrop>;
    MMGL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t items = (int64_t)((p.seq_q + 127) / 128) * p.heads * batch;
    dim3 grid((unsigned)(items < sm_count() ? items : sm_count()));
    kern<<<grid, 512, smem, stream>>>(mp.q, mp.d_o, mp.k, mp.v, p, (const __nv_bfloat16*)o, ldo,
                                      (const __nv_bfloat16*)d_o, lddo, stats, (__nv_bfloat16*)dq, lddq, delta_ws);
    if (int rc = check_launch("mmgl_attn_bwd(dq)")) return rc;
  }
  {
    const size_t smem = (2 + 2 * NB) * (size_t)TB + 65536 + 16 + 5 * 8 + 16;
    auto kern = sattn_bwd_dkv_kernel<D, kBias, kDrop>;
    MMGL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t items = (int64_t)((p.seq_k + 127) / 128) * p.heads * batch;
    dim3 grid((unsigned)(items < sm_count() ? items : sm_count()));
    kern<<<grid, 512, smem, stream>>>(mp.q, mp.d_o, mp.k, mp.v, p, stats, delta_ws, (__nv_bfloat16*)dk, lddk,
                                      (__nv_bfloat16*)dv, lddv);
    return check_launch("mmgl_attn_bwd(dkv)");
  }
}

template <int D>
int launch_bwd(const Maps& mp, const AttnParams& p, const void* o, int64_t ldo, const void* d_o, int64_t lddo,
               const float* stats, void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv, float* delta_ws,
               int64_t batch, cudaStream_t stream) {
  const bool bias = p.rel_bias != nullptr, drop = p.drop_thresh != 0;
#define MMGL_BWD(B_, R_) launch_bwd_v<D, B_, R_>(mp, p, o, ldo, d_o, lddo, stats, dq, lddq, dk, lddk, dv, lddv, delta_ws, batch, stream)
  if (bias) return drop ? MMGL_BWD(true, true) : MMGL_BWD(true, false);
  return drop ? MMGL_BWD(false, true) : MMGL_BWD(false, false);
#undef MMGL_BWD
}

int fill_params(const char* who, const mmgl_attn_args* a, AttnParams& p) {
  MMGL_REQUIRE(a->batch > 0 && a->seq_q > 0 && a->seq_k > 0 && a->heads > 0, "%s: empty problem", who);
  MMGL_REQUIRE(a->head_dim == 64 || a->head_dim == 128, "%s: head_dim must be 64 or 128 (got %lld)", who, (long long)a->head_dim);
  MMGL_REQUIRE(a->batch < 65536 && a->heads < 65536 && a->batch * a->heads * ((a->seq_q + a->seq_k) / 128 + 2) < (1ll << 31),
               "%s: batch/heads too large for the grid", who);
  MMGL_REQUIRE(a->seq_k <= 8192 && a->seq_q <= (1 << 20), "%s: seq_k must be <= 8192", who);
  MMGL_REQUIRE(!a->causal || a->seq_q <= a->seq_k, "%s: causal needs seq_q <= seq_k (keys = prefix + the queries' own positions)", who);
  MMGL_REQUIRE(a->scale > 0.f, "%s: scale must be positive", who);
  MMGL_REQUIRE(a->dropout_p >= 0.f && a->dropout_p < 1.f, "%s: dropout_p must be in [0,1)", who);
  p.key_mask = a->key_mask; p.rel_bias = a->rel_bias; p.cu_seqlens = a->cu_seqlens; p.d_rel_bias = nullptr;
  p.seq_q = (int)a->seq_q; p.seq_k = (int)a->seq_k; p.heads = (int)a->heads; p.causal = a->causal; p.batch = (int)a->batch;
  p.coff = (int)(a->seq_k - a->seq_q);
  p.scale = a->scale;
  p.drop_thresh = (uint32_t)(a->dropout_p * 65536.f + 0.5f);
  p.drop_scale = p.drop_thresh ? 65536.f / (65536.f - (float)p.drop_thresh) : 1.f;
  p.drop_seed = a->dropout_seed;
  return 0;
}

}  // namespace
}  // namespace mmgl

using namespace mmgl;

#ifdef MMGL_TRACE
extern "C" int mmgl_debug_trace(unsigned long long* host_dst) {
  cudaDeviceSynchronize();
  return (int)cudaMemcpyFromSymbol(host_dst, g_trace, sizeof(g_trace));
}
#endif

extern "C" int mmgl_attn_fwd(const mmgl_attn_args* a, void* stream_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  MMGL_REQUIRE(a != nullptr, "mmgl_attn_fwd: null args");
  MMGL_BIND(a->q, "mmgl_attn_fwd");
  AttnParams p;
  if (int rc = fill_params("mmgl_attn_fwd", a, p)) return rc;
  if (a->cu_seqlens != nullptr)
    MMGL_REQUIRE(a->seq_q == a->seq_k && a->total_tokens > 0 && a->key_mask == nullptr && a->rel_bias == nullptr && a->dropout_p == 0.f,
                 "mmgl_attn_fwd: cu_seqlens needs seq_q == seq_k (the longest sample), total_tokens, and no key mask / bias / dropout");
  MMGL_REQUIRE(a->k && a->v && a->o && (a->stats || a->cu_seqlens) && aligned16(a->q) && aligned16(a->k) && aligned16(a->v) && aligned16(a->o) &&
               a->ldq % 8 == 0 && a->ldk % 8 == 0 && a->ldv % 8 == 0 && a->ldo % 8 == 0,
               "mmgl_attn_fwd: pointers must be 16B aligned, leading dims %% 8 == 0");
  Maps mp;
  if (int rc = build_maps(mp, a->q, a->ldq, a->k, a->ldk, a->v, a->ldv, nullptr, 0, a->batch, a->seq_q, a->seq_k, a->heads, (int)a->head_dim,
                          a->cu_seqlens ? a->total_tokens : 0)) return rc;
  if (a->head_dim == 64) return launch_fwd<64>(mp, p, a->o, a->ldo, a->stats, a->batch, s);
  return launch_fwd<128>(mp, p, a->o, a->ldo, a->stats, a->batch, s);
}

extern "C" size_t mmgl_attn_bwd_workspace_bytes(int64_t batch, int64_t seq_q, int64_t heads) {
  return (size_t)(batch * seq_q * heads) * sizeof(float);
}

extern "C" int mmgl_attn_bwd(const mmgl_attn_args* a, const void* d_o, int64_t lddo, void* dq, int64_t lddq, void* dk,
                             int64_t lddk, void* dv, int64_t lddv, float* d_rel_bias, void* workspace,
                             size_t workspace_bytes, void* stream_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  MMGL_REQUIRE(a != nullptr, "mmgl_attn_bwd: null args");
  MMGL_BIND(a->q, "mmgl_attn_bwd");
  AttnParams p;
  if (int rc = fill_params("mmgl_attn_bwd", a, p)) return rc;
  MMGL_REQUIRE(d_o && a->k && a->v && a->o && a->stats && dq && dk && dv, "mmgl_attn_bwd: null pointer");
  MMGL_REQUIRE(a->cu_seqlens == nullptr, "mmgl_attn_bwd: variable-length batches are forward-only (frozen encoders)");
  MMGL_REQUIRE(d_rel_bias == nullptr || a->rel_bias != nullptr, "mmgl_attn_bwd: d_rel_bias without rel_bias");
  p.d_rel_bias = d_rel_bias;
  MMGL_REQUIRE(workspace != nullptr && workspace_bytes >= mmgl_attn_bwd_workspace_bytes(a->batch, a->seq_q, a->heads),
               "mmgl_attn_bwd: workspace too small (need mmgl_attn_bwd_workspace_bytes)");
  MMGL_REQUIRE(aligned16(d_o) && aligned16(a->q) && aligned16(a->k) && aligned16(a->v) && aligned16(a->o) && aligned16(dq) &&
               aligned16(dk) && aligned16(dv), "mmgl_attn_bwd: pointers must be 16B aligned");
  MMGL_REQUIRE(lddo % 8 == 0 && a->ldq % 8 == 0 && a->ldk % 8 == 0 && a->ldv % 8 == 0 && a->ldo % 8 == 0 && lddq % 8 == 0 &&
               lddk % 8 == 0 && lddv % 8 == 0, "mmgl_attn_bwd: leading dims must be multiples of 8");
  Maps mp;
  if (int rc = build_maps(mp, a->q, a->ldq, a->k, a->ldk, a->v, a->ldv, d_o, lddo, a->batch, a->seq_q, a->seq_k, a->heads, (int)a->head_dim)) return rc;
  float* ws = reinterpret_cast<float*>(workspace);
  if (a->head_dim == 64)
    return launch_bwd<64>(mp, p, a->o, a->ldo, d_o, lddo, a->stats, dq, lddq, dk, lddk, dv, lddv, ws, a->batch, s);
  return launch_bwd<128>(mp, p, a->o, a->ldo, d_o, lddo, a->stats, dq, lddq, dk, lddk, dv, lddv, ws, a->batch, s);
}
