// Helpers shared by the self-attention forward (sattn_sm100.cu) and backward (sattn_bwd_sm100.cu) kernels: launch
// parameters, bit-mask construction, dropout keep words, UMMA issue helpers over precomputed descriptor bases.
#pragma once
#include <cfloat>
#include <cstdlib>
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
  int dq_red;                // backward (experiment switch MMGL_SATTN_RED=1): fold dQ parts with red.global.add instead of ld + add + st
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
// the same with P read from tensor memory (TS form): the 64 keys of the half block are packed bf16 pairs in two groups of
// 16 columns, keys 0..31 at a_tmem + 0 and keys 32..63 at a_tmem + 32 (each softmax thread overwrites the head of the
// score columns it has just read)
template <int D>
__device__ __forceinline__ void mma_pv_half_ts(uint32_t tmem, uint32_t a_tmem, uint64_t db0, bool accumulate) {
  const uint32_t idesc = make_idesc_bf16(128, D, 0, 1);
#pragma unroll
  for (int k = 0; k < 4; ++k)
    umma_f16_ts(tmem, a_tmem + (k >> 1) * 32 + (k & 1) * 8, db0 + (uint64_t)((k * 2048) >> 4), idesc, (accumulate || k != 0) ? 1u : 0u);
}
template <int D>
__device__ __forceinline__ void tma_tile(uint8_t* dst, const CUtensorMap* map, uint64_t* bar, int col0, int row0) {
#pragma unroll
  for (int j = 0; j < D / 64; ++j) tma_load_2d(dst + j * 16384, map, bar, col0 + 64 * j, row0);
}



struct Maps { CUtensorMap q, k, v, d_o; };

inline int build_maps(Maps& mp, const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, const void* d_o,
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

inline size_t kbits_bytes(int64_t seq_k) {
  const size_t words = 4 * (size_t)((seq_k + 127) / 128);
  return (words + (words & 1)) * 4;
}

inline int fill_params(const char* who, const mmgl_attn_args* a, AttnParams& p) {
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
  {
    static const int red = [] { const char* e = getenv("MMGL_SATTN_RED"); return (e && e[0] == '1') ? 1 : 0; }();
    p.dq_red = red;
  }
  return 0;
}

}  // namespace
}  // namespace mmgl
