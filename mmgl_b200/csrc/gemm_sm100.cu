// tcgen05 GEMM with fused epilogue for sm_100a.
//
//   D = epilogue( A0 * B0^T (+ A1 * B1^T) )        bf16 x bf16 -> fp32 (TMEM) -> bf16/fp32
//
// Design (B200-first, not a port: the reference runs these contractions through nn.Linear/cuBLAS):
//   * persistent CTAs (one per SM), static round-robin tile scheduler, tile = 128 x BN (BN in 64/128/256)
//   * warp 0  : TMA producer  (cp.async.bulk.tensor, 128B swizzle, multi-stage mbarrier ring)
//   * warp 1  : TMEM allocator + single-thread tcgen05.mma issuer (UMMA 128 x BN x 16, fp32 accum in TMEM)
//   * warps 2-5: epilogue (tcgen05.ld -> registers -> fused bias/scale/ReLU/gate/residual -> global)
//   * two accumulator stages in TMEM so the epilogue of tile i overlaps the main loop of tile i+1
//   * both operands may be K-major or MN-major (UMMA descriptors do the transposition), so forward,
//     dgrad and wgrad of nn.Linear run through the same kernel with no transposed copies in HBM
//   * an optional second operand pair accumulates into the same TMEM tile (LoRA: x W^T + (x A^T) B^T;
//     GCN: [X, adj X] W^T without materialising the concat)
#include <cuda.h>
#include <cuda_bf16.h>

#include <mutex>
#include <unordered_map>

#include "../../include/mmgl_b200.h"
#include "common.cuh"
#include "ptx.cuh"

namespace mmgl {

constexpr int BM = 128;
constexpr int BK = 64;                 // 64 bf16 = 128 B = one swizzle row
constexpr int kGemmThreads = 320;      // 10 warps: TMA producer, MMA issuer, 8 epilogue warps
constexpr int kATileBytes = BM * BK * 2;
// per-k-block time of one wave of CTA-pair tiles relative to one wave of single-CTA 128 x 256 tiles (B200, tools/microbench.py:
// fc1 1410 vs 1300 TFLOP/s at BN = 256; the 256 x 128 pair tile is L2-ingest bound again)
constexpr size_t kSkFlagBytes = 1024;   // stream-K arrival counters at the head of the workspace
constexpr double kPairCost256 = 0.92;
constexpr double kPairCost128 = 1.50;

template <int BN> struct GemmCfg {
  static constexpr int kBTileBytes = BN * BK * 2;
  static constexpr int kStageBytes = kATileBytes + kBTileBytes;
  static constexpr int kStages = (BN == 256) ? 4 : (BN == 192 ? 5 : (BN == 128 ? 6 : 8));
  static constexpr int kTmemCols = (2 * BN <= 128) ? 128 : (2 * BN <= 256 ? 256 : 512);  // two accumulator stages, pow2
  static constexpr int kBarrierBytes = 256;
  static constexpr int kStoreStageBytes = 8 * 2048;   // one 32-row x 64-byte transpose buffer per epilogue warp
  static constexpr int kSmemBytes = kStages * kStageBytes + kBarrierBytes + kStoreStageBytes + 1024;  // +1024: manual alignment
};

struct GemmParams {
  int64_t m, n;
  int32_t kblocks0, kblocks1;
  int32_t m_blocks, n_blocks;
  int32_t n_fastest;  // tile rasterisation: 1 = consecutive tiles walk N first (each A tile is streamed once)
  int32_t m_blocks_pair;  // CTA-pair kernel: number of 256-row tiles
  // stream-K tail of the CTA-pair kernel: tiles [0, sk_full) run data-parallel (whole K per tile); each of the last
  // sk_tail tiles is cut into sk_split K-slices owned by different CTA pairs; slices > 0 park their fp32 partial tile in
  // sk_ws and bump sk_flags[tile]; slice 0 adds them in and runs the epilogue.  sk_split <= 1: plain data-parallel.
  // sk_park_all (skinny outputs): EVERY slice parks its partial and splitk_reduce_kernel sums them and runs the epilogue,
  // so the fix-up is parallel over the output instead of serial in the owner's 512 threads.
  int32_t sk_full, sk_split, sk_tail, sk_park_all;
  float* sk_ws; int32_t* sk_flags;
  void* d; int64_t ldd; int32_t out_fp32; int32_t accumulate;
  float alpha; int32_t relu;
  const float* bias;
  const float* gate;
  const __nv_bfloat16* residual; int64_t ldres;
  __nv_bfloat16* aux; int64_t ldaux;
  const __nv_bfloat16* relu_mask; int64_t ldmask;
  int32_t vec_ok;
  int32_t staged_store;   // bf16 fast path: transpose each chunk through shared memory for 64-byte row segments per store
  uint32_t drop_thresh; float drop_scale; uint64_t drop_seed; int64_t drop_groups;  // drop_thresh == 0: no dropout
};

// activation codes of mmgl_gemm_args.relu: 1 = ReLU, 2 = GELU (erf form, nn.GELU / HF "gelu"), 3 = quick-GELU (CLIP)
__device__ __forceinline__ float act_gelu(float v) {
  // 0.5 v (1 + erf(v / sqrt2)) with erf from Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7, far below the bf16 rounding
  // of the output): one rcp + one ex2 and no branches, ~3x fewer issue slots than erff() in a short-K epilogue.
  // q = 1 - erf(z), z = |v| / sqrt2, is formed directly so the negative side does not cancel.
  const float z = fabsf(v) * 0.70710678118654752f;
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.f)));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  float ez;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ez) : "f"(-1.4426950408889634f * z * z));
  const float q = poly * t * ez;
  const float hv = 0.5f * v;
  return v >= 0.f ? fmaf(-hv, q, v) : hv * q;
}
__device__ __forceinline__ float act_quick_gelu(float v) {
  // v * sigmoid(1.702 v) = v / (1 + 2^(-1.702 log2e v))
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-2.4554669595930157f * v));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + e));
  return v * r;
}
// the activation switch sits OUTSIDE the element loop: one uniform branch per chunk, straight-line code per element
template <int N>
__device__ __forceinline__ void apply_act_vec(float (&v)[N], int act) {
  if (act == 1) {
#pragma unroll
    for (int j = 0; j < N; ++j) v[j] = fmaxf(v[j], 0.f);
  } else if (act == 2) {
#pragma unroll
    for (int j = 0; j < N; ++j) v[j] = act_gelu(v[j]);
  } else {
#pragma unroll
    for (int j = 0; j < N; ++j) v[j] = act_quick_gelu(v[j]);
  }
}
__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == 1) return fmaxf(v, 0.f);
  if (act == 2) return act_gelu(v);
  return act_quick_gelu(v);
}

// 8 consecutive output columns of one row: fused epilogue + store.
__device__ __forceinline__ void epilogue_store8(const GemmParams& p, float (&v)[8], int64_t row, int64_t col,
                                                float gate_t) {
  if (p.bias != nullptr) {
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + col + 4));
    v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
    v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] *= p.alpha;
  if (p.relu) apply_act_vec<8>(v, p.relu);
  if (p.relu_mask != nullptr) {
    const uint4 mk = __ldg(reinterpret_cast<const uint4*>(p.relu_mask + row * p.ldmask + col));
    const uint32_t w[4] = {mk.x, mk.y, mk.z, mk.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (!(bf16lo(w[j]) > 0.f)) v[2 * j] = 0.f;
      if (!(bf16hi(w[j]) > 0.f)) v[2 * j + 1] = 0.f;
    }
  }
  if (p.drop_thresh != 0) {
    const DropBits bits = dropout_bits(p.drop_seed, row, col >> 3, p.drop_groups);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = dropout_keep(bits, j, p.drop_thresh) ? v[j] * p.drop_scale : 0.f;
  }
  if (p.aux != nullptr) {
    uint4 o;
    o.x = pack_bf16(v[0], v[1]); o.y = pack_bf16(v[2], v[3]); o.z = pack_bf16(v[4], v[5]); o.w = pack_bf16(v[6], v[7]);
    *reinterpret_cast<uint4*>(p.aux + row * p.ldaux + col) = o;
  }
  if (p.gate != nullptr) {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] *= gate_t;
  }
  if (p.residual != nullptr) {
    const uint4 r = __ldg(reinterpret_cast<const uint4*>(p.residual + row * p.ldres + col));
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) { v[2 * j] += bf16lo(w[j]); v[2 * j + 1] += bf16hi(w[j]); }
  }
  if (p.out_fp32) {
    float* dp = reinterpret_cast<float*>(p.d) + row * p.ldd + col;
    if (p.accumulate) {
      const float4 a0 = *reinterpret_cast<const float4*>(dp);
      const float4 a1 = *reinterpret_cast<const float4*>(dp + 4);
      v[0] += a0.x; v[1] += a0.y; v[2] += a0.z; v[3] += a0.w;
      v[4] += a1.x; v[5] += a1.y; v[6] += a1.z; v[7] += a1.w;
    }
    *reinterpret_cast<float4*>(dp) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(dp + 4) = make_float4(v[4], v[5], v[6], v[7]);
  } else {
    __nv_bfloat16* dp = reinterpret_cast<__nv_bfloat16*>(p.d) + row * p.ldd + col;
    if (p.accumulate) {
      const uint4 a = *reinterpret_cast<const uint4*>(dp);
      const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) { v[2 * j] += bf16lo(w[j]); v[2 * j + 1] += bf16hi(w[j]); }
    }
    uint4 o;
    o.x = pack_bf16(v[0], v[1]); o.y = pack_bf16(v[2], v[3]); o.z = pack_bf16(v[4], v[5]); o.w = pack_bf16(v[6], v[7]);
    *reinterpret_cast<uint4*>(dp) = o;
  }
}

// Scalar tail (N not a multiple of 8, or unaligned epilogue tensors).
__device__ __forceinline__ void epilogue_store1(const GemmParams& p, float v, int64_t row, int64_t col, float gate_t) {
  if (p.bias != nullptr) v += __ldg(p.bias + col);
  v *= p.alpha;
  if (p.relu) v = apply_act(v, p.relu);
  if (p.relu_mask != nullptr && !(__bfloat162float(p.relu_mask[row * p.ldmask + col]) > 0.f)) v = 0.f;
  if (p.drop_thresh != 0) {
    const DropBits bits = dropout_bits(p.drop_seed, row, col >> 3, p.drop_groups);
    v = dropout_keep(bits, (int)(col & 7), p.drop_thresh) ? v * p.drop_scale : 0.f;
  }
  if (p.aux != nullptr) p.aux[row * p.ldaux + col] = __float2bfloat16_rn(v);
  if (p.gate != nullptr) v *= gate_t;
  if (p.residual != nullptr) v += __bfloat162float(p.residual[row * p.ldres + col]);
  if (p.out_fp32) {
    float* dp = reinterpret_cast<float*>(p.d) + row * p.ldd + col;
    if (p.accumulate) v += *dp;
    *dp = v;
  } else {
    __nv_bfloat16* dp = reinterpret_cast<__nv_bfloat16*>(p.d) + row * p.ldd + col;
    if (p.accumulate) v += __bfloat162float(*dp);
    *dp = __float2bfloat16_rn(v);
  }
}

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
  uint4 o;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(o.x), "=r"(o.y), "=r"(o.z), "=r"(o.w) : "r"(addr) : "memory");
  return o;
}

// Epilogue of one warp: its 32 accumulator rows x the 32-column chunks [c_begin, c_end) of the tile at TMEM `taddr`.
// For short-K tiles the epilogue, not the MMA, is the critical path, so the common case (bf16 output, 16-byte aligned
// operands) issues the residual / ReLU-mask loads of chunk c+1 before it touches chunk c: their global latency hides
// behind the TMEM load and the arithmetic of the current chunk.
//
// Stores: a thread owns one output ROW, so a direct store instruction of the warp touches 32 different 128-byte lines
// with 16 bytes each.  When the warp's 32 rows are all inside the matrix the bf16 chunk (32 rows x 64 bytes) is instead
// transposed through `stg`, the warp's private 2 KB of shared memory (16-byte units XOR-swizzled so that neither side
// has bank conflicts), and written back as 8 rows x 64 contiguous bytes per instruction: a quarter of the lines per
// store.  stg == 0 keeps the direct stores.
__device__ __forceinline__ void epilogue_chunks(const GemmParams& p, uint32_t taddr, int64_t row, int n0, int c_begin,
                                                int c_end, float gate_t, uint32_t stg = 0) {
  const bool fast = p.vec_ok && !p.out_fp32 && !p.accumulate;
  const bool row_ok = row < p.m;
  const int lane = threadIdx.x & 31;
  const bool staged = stg != 0 && (row - lane + 31) < p.m;     // warp-uniform: every row of this warp is valid
  uint4 res_n[4], msk_n[4];
#pragma unroll
  for (int g = 0; g < 4; ++g) { res_n[g] = make_uint4(0, 0, 0, 0); msk_n[g] = make_uint4(0, 0, 0, 0); }
  auto prefetch = [&](int c) {
    const int64_t col0 = n0 + c * 32;
    if (staged && col0 + 32 <= p.n) {
      // coalesced pattern: instruction i fetches rows 8i .. 8i+7 of the warp, 64 contiguous bytes each; transpose_in()
      // hands every thread its own row afterwards
      const int64_t r0 = row - lane + (lane >> 2);
      const int cg = (lane & 3) * 8;
      if (p.residual != nullptr) {
#pragma unroll
        for (int i = 0; i < 4; ++i) res_n[i] = __ldg(reinterpret_cast<const uint4*>(p.residual + (r0 + 8 * i) * p.ldres + col0 + cg));
      }
      if (p.relu_mask != nullptr) {
#pragma unroll
        for (int i = 0; i < 4; ++i) msk_n[i] = __ldg(reinterpret_cast<const uint4*>(p.relu_mask + (r0 + 8 * i) * p.ldmask + col0 + cg));
      }
    } else if (row_ok && col0 + 32 <= p.n) {
      if (p.residual != nullptr) {
#pragma unroll
        for (int g = 0; g < 4; ++g) res_n[g] = __ldg(reinterpret_cast<const uint4*>(p.residual + row * p.ldres + col0) + g);
      }
      if (p.relu_mask != nullptr) {
#pragma unroll
        for (int g = 0; g < 4; ++g) msk_n[g] = __ldg(reinterpret_cast<const uint4*>(p.relu_mask + row * p.ldmask + col0) + g);
      }
    }
  };
  auto transpose_in = [&](uint4 (&t)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r_in = 8 * i + (lane >> 2), cg = lane & 3;
      st_shared_v4(stg + r_in * 64 + ((cg ^ ((r_in >> 1) & 3)) << 4), t[i].x, t[i].y, t[i].z, t[i].w);
    }
    __syncwarp();
    const uint32_t sw = (lane >> 1) & 3;
#pragma unroll
    for (int g = 0; g < 4; ++g) t[g] = ld_shared_v4(stg + lane * 64 + ((g ^ sw) << 4));
    __syncwarp();
  };
  auto store_chunk = [&](const float (&v)[32], __nv_bfloat16* base, int64_t ld, int64_t col0) {
    if (staged) {
      const uint32_t sw = (lane >> 1) & 3;
#pragma unroll
      for (int g = 0; g < 4; ++g)
        st_shared_v4(stg + lane * 64 + ((g ^ sw) << 4), pack_bf16(v[8 * g], v[8 * g + 1]), pack_bf16(v[8 * g + 2], v[8 * g + 3]),
                     pack_bf16(v[8 * g + 4], v[8 * g + 5]), pack_bf16(v[8 * g + 6], v[8 * g + 7]));
      __syncwarp();
      __nv_bfloat16* dbase = base + (row - lane) * ld + col0;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r_in = 8 * i + (lane >> 2), cg = lane & 3;
        const uint4 o = ld_shared_v4(stg + r_in * 64 + ((cg ^ ((r_in >> 1) & 3)) << 4));
        *reinterpret_cast<uint4*>(dbase + (int64_t)r_in * ld + cg * 8) = o;
      }
      __syncwarp();
    } else {
      __nv_bfloat16* dp = base + row * ld + col0;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint4 o;
        o.x = pack_bf16(v[8 * g], v[8 * g + 1]); o.y = pack_bf16(v[8 * g + 2], v[8 * g + 3]);
        o.z = pack_bf16(v[8 * g + 4], v[8 * g + 5]); o.w = pack_bf16(v[8 * g + 6], v[8 * g + 7]);
        reinterpret_cast<uint4*>(dp)[g] = o;
      }
    }
  };
  if (fast) prefetch(c_begin);
#pragma unroll 1
  for (int c = c_begin; c < c_end; ++c) {
    uint32_t r[32];
    tmem_ld_32x32(taddr + c * 32, r);
    uint4 res[4], msk[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) { res[g] = res_n[g]; msk[g] = msk_n[g]; }
    if (fast && c + 1 < c_end) prefetch(c + 1);
    tmem_ld_wait();
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
    const int64_t col0 = n0 + c * 32;
    if (!row_ok || col0 >= p.n) continue;
    if (fast && col0 + 32 <= p.n) {
      if (staged) {
        if (p.residual != nullptr) transpose_in(res);
        if (p.relu_mask != nullptr) transpose_in(msk);
      }
      if (p.bias != nullptr) {
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col0) + g);
          v[4 * g] += b4.x; v[4 * g + 1] += b4.y; v[4 * g + 2] += b4.z; v[4 * g + 3] += b4.w;
        }
      }
      if (p.alpha != 1.f) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] *= p.alpha;
      }
      if (p.relu) apply_act_vec<32>(v, p.relu);
      if (p.relu_mask != nullptr) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const uint32_t w[4] = {msk[g].x, msk[g].y, msk[g].z, msk[g].w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (!(bf16lo(w[j]) > 0.f)) v[8 * g + 2 * j] = 0.f;
            if (!(bf16hi(w[j]) > 0.f)) v[8 * g + 2 * j + 1] = 0.f;
          }
        }
      }
      if (p.drop_thresh != 0) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const DropBits bits = dropout_bits(p.drop_seed, row, (col0 >> 3) + g, p.drop_groups);
#pragma unroll
          for (int j = 0; j < 8; ++j) v[8 * g + j] = dropout_keep(bits, j, p.drop_thresh) ? v[8 * g + j] * p.drop_scale : 0.f;
        }
      }
      if (p.aux != nullptr) store_chunk(v, p.aux, p.ldaux, col0);
      if (p.gate != nullptr) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] *= gate_t;
      }
      if (p.residual != nullptr) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const uint32_t w[4] = {res[g].x, res[g].y, res[g].z, res[g].w};
#pragma unroll
          for (int j = 0; j < 4; ++j) { v[8 * g + 2 * j] += bf16lo(w[j]); v[8 * g + 2 * j + 1] += bf16hi(w[j]); }
        }
      }
      store_chunk(v, reinterpret_cast<__nv_bfloat16*>(p.d), p.ldd, col0);
    } else if (p.vec_ok && col0 + 32 <= p.n) {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        float v8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v8[j] = v[g * 8 + j];
        epilogue_store8(p, v8, row, col0 + g * 8, gate_t);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (col0 + j < p.n) epilogue_store1(p, v[j], row, col0 + j, gate_t);
    }
  }
}

template <int BN, int A_MN, int B_MN>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_b0,
                    const __grid_constant__ CUtensorMap map_a1, const __grid_constant__ CUtensorMap map_b1,
                    const GemmParams p) {
  using Cfg = GemmCfg<BN>;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full = empty_bar + kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // lane-0 broadcast: provably warp-uniform for ptxas
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.m_blocks * p.n_blocks;
  const int kblocks = p.kblocks0 + p.kblocks1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a0);
    tma_prefetch_desc(&map_b0);
    if (p.kblocks1 > 0) { tma_prefetch_desc(&map_a1); tma_prefetch_desc(&map_b1); }
    for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], 8); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<Cfg::kTmemCols>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);
  pdl_launch();
  pdl_wait();   // everything above overlapped the previous kernel's tail; global memory is touched only below

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (p.n_fastest ? tile / p.n_blocks : tile % p.m_blocks) * BM;
        const int n0 = (p.n_fastest ? tile % p.n_blocks : tile / p.m_blocks) * BN;
        for (int kb = 0; kb < kblocks; ++kb) {
          const bool second = kb >= p.kblocks0;
          const CUtensorMap* ma = second ? &map_a1 : &map_a0;
          const CUtensorMap* mb = second ? &map_b1 : &map_b0;
          const int k0 = (second ? kb - p.kblocks0 : kb) * BK;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          uint8_t* sa = smem + stage * Cfg::kStageBytes;
          uint8_t* sb = sa + kATileBytes;
          if (A_MN) {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j) tma_load_2d(sa + j * 8192, ma, &full_bar[stage], m0 + 64 * j, k0);
          } else {
            tma_load_2d(sa, ma, &full_bar[stage], k0, m0);
          }
          if (B_MN) {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j) tma_load_2d(sb + j * 8192, mb, &full_bar[stage], n0 + 64 * j, k0);
          } else {
            tma_load_2d(sb, mb, &full_bar[stage], k0, n0);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer: the whole warp walks the schedule, one elected lane issues =====================
    // (operands derived from warp-uniform values stay in uniform registers: the four UTCHMMAs of a k-block go out back to
    // back instead of one ELECT / R2UR waterfall loop each, which `if (lane == 0)` compiles to)
    {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN, A_MN, B_MN);
      int stage = 0; uint32_t phase = 0;
      int as = 0; uint32_t aphase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint32_t sb = sa + kATileBytes;
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              const uint64_t da = A_MN ? make_smem_desc(sa + k * 2048, 8192, 1024) : make_smem_desc(sa + k * 32, 16, 1024);
              const uint64_t db = B_MN ? make_smem_desc(sb + k * 2048, 8192, 1024) : make_smem_desc(sb + k * 32, 16, 1024);
              umma_f16_ss(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
            }
            umma_commit(&empty_bar[stage]);  // smem slot reusable once these MMAs have read it
            if (kb == kblocks - 1) umma_commit(&tmem_full[as]);   // accumulator complete -> epilogue
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue warps (2..9): lane quarter = warp % 4, column half = (warp - 2) / 4 =====================
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;
    constexpr int kChunksPerHalf = BN / 64;
    const float gate_t = (p.gate != nullptr) ? tanh_precise(__ldg(p.gate)) : 1.f;
    const uint32_t stg = p.staged_store ? smem_u32(smem + kStages * Cfg::kStageBytes + 256 + (warp - 2) * 2048) : 0u;
    int as = 0; uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (p.n_fastest ? tile / p.n_blocks : tile % p.m_blocks) * BM;
      const int n0 = (p.n_fastest ? tile % p.n_blocks : tile / p.m_blocks) * BN;
      mbar_wait(&tmem_full[as], aphase);
      tc_fence_after();
      const int64_t row = m0 + quarter * 32 + lane;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + as * BN;
      epilogue_chunks(p, taddr, row, n0, half * kChunksPerHalf, (half + 1) * kChunksPerHalf, gate_t, stg);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[as]);
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<Cfg::kTmemCols>(tmem_base);
}


// ------------------------------------------------------------------------------------ CTA-pair (cta_group::2) variant
// Two CTAs on the two SMs of a TPC cooperate on a 256 x BN tile: each loads its own 128 rows of A and HALF of the
// B tile, one thread of the leader CTA issues tcgen05.mma.cta_group::2 (M = 256), each CTA's TMEM receives its own 128
// accumulator rows.  Per SM and k-block the operand ingest drops from 16 KB + BN*128 B to 16 KB + BN*64 B, which is what
// bounds the single-CTA kernel (L2 -> SM bandwidth), so the pair kernel is tensor-pipe bound at BN = 256.
struct PairWork { int tile, kb0, kb1, slice, tail; };
// it-th work item of a CTA pair; false when the pair is done.  Identical in all warp roles.
__device__ __forceinline__ bool pair_next_work(const GemmParams& p, int cluster, int num_clusters, int num_tiles, int kblocks,
                                               int it, PairWork& w) {
  const int tile = cluster + it * num_clusters;
  w.kb0 = 0; w.kb1 = kblocks; w.slice = 0; w.tail = -1;
  if (tile < p.sk_full || p.sk_split <= 1) {
    w.tile = tile;
    return tile < num_tiles;
  }
  if (tile >= p.sk_full + num_clusters || cluster >= p.sk_tail * p.sk_split) return false;   // one tail item per pair
  w.tail = cluster / p.sk_split;
  w.slice = cluster % p.sk_split;
  w.tile = p.sk_full + w.tail;
  w.kb0 = (int)(((int64_t)w.slice * kblocks) / p.sk_split);
  w.kb1 = (int)(((int64_t)(w.slice + 1) * kblocks) / p.sk_split);
  return true;
}
__device__ __forceinline__ void spin_until_at_least(const int32_t* flag, int32_t want) {
  const long long t0 = clock64();
  while (*reinterpret_cast<const volatile int32_t*>(flag) < want) {
    __nanosleep(64);
    if (clock64() - t0 > 4000000000LL) {
      printf("mmgl: stream-K fix-up wait timed out (block %d)\n", (int)blockIdx.x);
      __trap();
    }
  }
}

template <int BN> struct PairCfg {
  static constexpr int kBHalfBytes = (BN / 2) * BK * 2;
  static constexpr int kStageBytes = kATileBytes + kBHalfBytes;   // per CTA
  static constexpr int kStages = (BN == 256) ? 6 : 8;
  static constexpr int kTmemCols = (2 * BN <= 256) ? 256 : 512;
  static constexpr int kStoreStageBytes = 8 * 2048;
  static constexpr int kSmemBytes = kStages * kStageBytes + 256 + kStoreStageBytes + 1024;
};

template <int BN, int A_MN, int B_MN>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tcgen05_pair_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_b0,
                         const __grid_constant__ CUtensorMap map_a1, const __grid_constant__ CUtensorMap map_b1,
                         const GemmParams p) {
  using Cfg = PairCfg<BN>;
  constexpr int kStages = Cfg::kStages;
  constexpr int BNH = BN / 2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full = empty_bar + kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // lane-0 broadcast: provably warp-uniform for ptxas
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const int num_tiles = p.m_blocks_pair * p.n_blocks;
  const int kblocks = p.kblocks0 + p.kblocks1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a0);
    tma_prefetch_desc(&map_b0);
    if (p.kblocks1 > 0) { tma_prefetch_desc(&map_a1); tma_prefetch_desc(&map_b1); }
    for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], 16); }  // 8 warps x 2 CTAs
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_pair<Cfg::kTmemCols>(tmem_ptr);
  tc_fence_before();
  cluster_sync_all();   // barrier inits and the TMEM allocation of BOTH CTAs are visible before any remote signal
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);
  pdl_launch();
  pdl_wait();   // the prologue above overlapped the previous kernel's tail; global memory is touched only below

  if (warp == 0) {
    // ===================== TMA producer (both CTAs: own A rows, own half of B) =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      PairWork w;
      for (int it = 0; pair_next_work(p, cluster, num_clusters, num_tiles, kblocks, it, w); ++it) {
        const int tile = w.tile;
        const int mt = p.n_fastest ? tile / p.n_blocks : tile % p.m_blocks_pair;
        const int nt = p.n_fastest ? tile % p.n_blocks : tile / p.m_blocks_pair;
        const int m0 = mt * 256 + (int)rank * BM;
        const int n0 = nt * BN + (int)rank * BNH;
        for (int kb = w.kb0; kb < w.kb1; ++kb) {
          const bool second = kb >= p.kblocks0;
          const CUtensorMap* ma = second ? &map_a1 : &map_a0;
          const CUtensorMap* mb = second ? &map_b1 : &map_b0;
          const int k0 = (second ? kb - p.kblocks0 : kb) * BK;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::kStageBytes);
          uint8_t* sa = smem + stage * Cfg::kStageBytes;
          uint8_t* sb = sa + kATileBytes;
          if (A_MN) {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j) tma_load_2d_pair(sa + j * 8192, ma, &full_bar[stage], m0 + 64 * j, k0);
          } else {
            tma_load_2d_pair(sa, ma, &full_bar[stage], k0, m0);
          }
          if (B_MN) {
#pragma unroll
            for (int j = 0; j < BNH / 64; ++j) tma_load_2d_pair(sb + j * 8192, mb, &full_bar[stage], n0 + 64 * j, k0);
          } else {
            tma_load_2d_pair(sb, mb, &full_bar[stage], k0, n0);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer (warp 1 of the LEADER CTA; one elected lane issues, see the single-CTA kernel) =====
    if (rank == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(256, BN, A_MN, B_MN);
      int stage = 0; uint32_t phase = 0;
      int as = 0; uint32_t aphase = 0;
      PairWork w;
      for (int it = 0; pair_next_work(p, cluster, num_clusters, num_tiles, kblocks, it, w); ++it) {
        mbar_wait(&tmem_empty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = w.kb0; kb < w.kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint32_t sb = sa + kATileBytes;
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              const uint64_t da = A_MN ? make_smem_desc(sa + k * 2048, 8192, 1024) : make_smem_desc(sa + k * 32, 16, 1024);
              const uint64_t db = B_MN ? make_smem_desc(sb + k * 2048, 8192, 1024) : make_smem_desc(sb + k * 32, 16, 1024);
              umma_f16_ss_pair(d_tmem, da, db, idesc, (kb != w.kb0 || k != 0) ? 1u : 0u);
            }
            umma_commit_pair(&empty_bar[stage]);   // frees this stage in BOTH CTAs
            if (kb == w.kb1 - 1) umma_commit_pair(&tmem_full[as]);   // accumulator complete -> epilogue warps of BOTH CTAs
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue warps (2..9) of both CTAs: own 128 accumulator rows, two column halves =====================
    const int quarter = warp & 3;
    const int c_begin = ((warp - 2) >> 2) * (BN / 64), c_end = c_begin + BN / 64;
    const float gate_t = (p.gate != nullptr) ? tanh_precise(__ldg(p.gate)) : 1.f;
    const uint32_t stg = p.staged_store ? smem_u32(smem + kStages * Cfg::kStageBytes + 256 + (warp - 2) * 2048) : 0u;
    int as = 0; uint32_t aphase = 0;
    PairWork w;
    for (int it = 0; pair_next_work(p, cluster, num_clusters, num_tiles, kblocks, it, w); ++it) {
      const int tile = w.tile;
      const int mt = p.n_fastest ? tile / p.n_blocks : tile % p.m_blocks_pair;
      const int nt = p.n_fastest ? tile % p.n_blocks : tile / p.m_blocks_pair;
      const int n0 = nt * BN;
      mbar_wait(&tmem_full[as], aphase);
      tc_fence_after();
      const int row_in_tile = (int)rank * BM + quarter * 32 + lane;
      const int64_t row = (int64_t)mt * 256 + row_in_tile;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + as * BN;
      const bool park_all = p.sk_park_all != 0 && w.tail >= 0;
      const int others = (w.tail >= 0) ? (park_all ? p.sk_split : p.sk_split - 1) : 0;
      if (w.slice > 0 || park_all) {
        // ---- stream-K contributor: park the raw fp32 partial tile, then signal the owner (park_all: the reduce kernel
        // that follows on the stream reads it; rows beyond M are never read)
        float* dst = p.sk_ws + ((size_t)(w.tail * others + (park_all ? w.slice : w.slice - 1)) * 256 + row_in_tile) * BN;
#pragma unroll 1
        for (int c = c_begin; c < c_end; ++c) {
          uint32_t r[32];
          tmem_ld_32x32(taddr + c * 32, r);
          tmem_ld_wait();
          if (park_all && row >= p.m) continue;
#pragma unroll
          for (int g = 0; g < 8; ++g)
            __stcg(reinterpret_cast<float4*>(dst + c * 32) + g,
                   make_float4(__uint_as_float(r[4 * g]), __uint_as_float(r[4 * g + 1]), __uint_as_float(r[4 * g + 2]),
                               __uint_as_float(r[4 * g + 3])));
        }
        if (!park_all) {
          __threadfence();
          __syncwarp();
          if (lane == 0) atomicAdd(p.sk_flags + w.tail, 1);
        }
      } else if (others > 0) {
        // ---- stream-K owner: wait for the 16 epilogue warps of every contributor pair, add their partials
        if (lane == 0) spin_until_at_least(p.sk_flags + w.tail, 16 * others);
        __syncwarp();
        __threadfence();
#pragma unroll 1
        for (int c = c_begin; c < c_end; ++c) {
          uint32_t r[32];
          tmem_ld_32x32(taddr + c * 32, r);
          tmem_ld_wait();
          for (int j = 0; j < others; ++j) {
            const float4* src = reinterpret_cast<const float4*>(
                p.sk_ws + ((size_t)(w.tail * others + j) * 256 + row_in_tile) * BN + c * 32);
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              const float4 v = __ldcg(src + g);
              r[4 * g] = __float_as_uint(__uint_as_float(r[4 * g]) + v.x);
              r[4 * g + 1] = __float_as_uint(__uint_as_float(r[4 * g + 1]) + v.y);
              r[4 * g + 2] = __float_as_uint(__uint_as_float(r[4 * g + 2]) + v.z);
              r[4 * g + 3] = __float_as_uint(__uint_as_float(r[4 * g + 3]) + v.w);
            }
          }
          const int64_t col0 = n0 + c * 32;
          if (row < p.m && col0 < p.n) {
            if (p.vec_ok && col0 + 32 <= p.n) {
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                float v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[g * 8 + j]);
                epilogue_store8(p, v, row, col0 + g * 8, gate_t);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j < p.n) epilogue_store1(p, __uint_as_float(r[j]), row, col0 + j, gate_t);
            }
          }
        }
      } else {
        epilogue_chunks(p, taddr, row, n0, c_begin, c_end, gate_t, stg);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(&tmem_empty[as]);
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  }

  tc_fence_before();
  cluster_sync_all();   // nobody leaves while the peer may still signal its barriers or read its TMEM / smem
  if (warp == 1) tmem_dealloc_pair<Cfg::kTmemCols>(tmem_base);
}

// Second half of a K-sliced skinny GEMM: sums the sk_split parked partial tiles of every output element in slice order
// (deterministic) and applies the full epilogue.  One thread per output element; the outputs are small by construction.
__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const GemmParams p, int bn) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= p.m * p.n) return;
  const int64_t row = e / p.n, col = e % p.n;
  const int mt = (int)(row >> 8), rit = (int)(row & 255);
  const int nt = (int)(col / bn), cit = (int)(col % bn);
  const int tile = p.n_fastest ? mt * p.n_blocks + nt : nt * p.m_blocks_pair + mt;
  const float* src = p.sk_ws + (((size_t)(tile - p.sk_full) * p.sk_split) * 256 + rit) * bn + cit;
  float acc = 0.f;
#pragma unroll 4
  for (int s = 0; s < p.sk_split; ++s) acc += __ldcg(src + (size_t)s * 256 * bn);   // loads overlap, adds stay in slice order
  const float gate_t = (p.gate != nullptr) ? tanh_precise(__ldg(p.gate)) : 1.f;
  epilogue_store1(p, acc, row, col, gate_t);
}

// --------------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

struct MapKey {
  const void* ptr; uint64_t d0, d1, ld; uint32_t b0, b1;
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && d0 == o.d0 && d1 == o.d1 && ld == o.ld && b0 == o.b0 && b1 == o.b1;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr);
    auto mix = [&](uint64_t v) { h ^= v + 0x9e3779b97f4a7c15ULL + (h << 6) + (h >> 2); };
    mix(k.d0); mix(k.d1); mix(k.ld); mix(k.b0); mix(k.b1);
    return h;
  }
};

// 2-D bf16 tensor map: dims (inner d0, outer d1), row pitch ld elements, box (b0 = 64 inner, b1 rows), 128B swizzle.
// Descriptors only encode address + geometry, so caching by (ptr, geometry) is safe across reuse of the address.
int make_tensor_map_2d(CUtensorMap* out, const void* ptr, uint64_t d0, uint64_t d1, uint64_t ld, uint32_t b0,
                       uint32_t b1) {
  static std::mutex mu;
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  const MapKey key{ptr, d0, d1, ld, b0, b1};
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return 0; }
  }
  EncodeTiledFn enc = get_encode_fn();
  if (enc == nullptr) { set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)"); return 4; }
  const cuuint64_t dims[2] = {d0, d1};
  const cuuint64_t strides[1] = {ld * 2};
  const cuuint32_t box[2] = {b0, b1};
  const cuuint32_t estr[2] = {1, 1};
  CUtensorMap m;
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): ptr=%p dims=(%llu,%llu) ld=%llu box=(%u,%u)", (int)r, ptr,
              (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)ld, b0, b1);
    return 4;
  }
  {
    std::lock_guard<std::mutex> lk(mu);
    if (cache.size() > 4096) cache.clear();
    cache.emplace(key, m);
  }
  *out = m;
  return 0;
}

// operand with `rows` along M or N and `k` along K
static int operand_map(CUtensorMap* out, const void* ptr, int mn_major, int64_t rows, int64_t k, int64_t ld,
                       int box_rows) {
  if (mn_major) return make_tensor_map_2d(out, ptr, (uint64_t)rows, (uint64_t)k, (uint64_t)ld, 64, 64);
  return make_tensor_map_2d(out, ptr, (uint64_t)k, (uint64_t)rows, (uint64_t)ld, 64, (uint32_t)box_rows);
}

static int pick_block_n(int64_t m, int64_t n, int sms) {
  const int64_t mb = (m + BM - 1) / BM;
  const int cand[4] = {256, 192, 128, 64};
  // measured time of one 64-deep k-block of a 128 x BN tile, relative to BN = 256 (B200, profiles/r01 microbench):
  // small tiles are far from proportionally cheaper (operand traffic per SM does not shrink with BN)
  const double tile_cost[4] = {1.00, 0.88, 0.82, 0.76};
  int best = 256; double best_cost = 1e30;
  for (int i = 0; i < 4; ++i) {
    const int64_t tiles = mb * ((n + cand[i] - 1) / cand[i]);
    const int64_t waves = (tiles + sms - 1) / sms;
    const double cost = waves * tile_cost[i];
    if (cost < best_cost * 0.98) { best_cost = cost; best = cand[i]; }
  }
  return best;
}

// CTA-pair selection: compare estimated time (waves x per-tile cost); costs relative to one k-block of a 128 x 256
// single-CTA tile, measured on B200 (tools/microbench.py).  Returns true and sets *bn when the pair kernel wins.
static bool pick_pair(int64_t m, int64_t n, int sms, int* bn, bool stream_k) {
  if (m < 256) return false;
  const int single_cand[4] = {256, 192, 128, 64};
  const double single_cost[4] = {1.00, 0.88, 0.82, 0.76};
  const int64_t mb = (m + BM - 1) / BM;
  double best_single = 1e30;
  for (int i = 0; i < 4; ++i) {
    const int64_t tiles = mb * ((n + single_cand[i] - 1) / single_cand[i]);
    const double c = (double)((tiles + sms - 1) / sms) * single_cost[i];
    if (c < best_single) best_single = c;
  }
  const int pair_cand[2] = {256, 128};
  const double pair_cost[2] = {kPairCost256, kPairCost128};   // one k-block of a 256 x BN pair tile
  const int pairs = sms / 2;
  const int64_t mb2 = (m + 255) / 256;
  double best_pair = 1e30; int best_bn = 256;
  for (int i = 0; i < 2; ++i) {
    const int64_t tiles = mb2 * ((n + pair_cand[i] - 1) / pair_cand[i]);
    double waves = (double)((tiles + pairs - 1) / pairs);
    if (stream_k && i == 0) {   // the tail wave is cut into K-slices: costs 1/split of a wave plus the fix-up
      const int64_t r = tiles % pairs;
      int64_t split = r > 0 ? pairs / r : 1;
      if (split > 8) split = 8;
      if (split >= 2) waves = (double)(tiles / pairs) + 1.0 / (double)split + 0.10;
    }
    const double c = waves * pair_cost[i];
    if (c < best_pair) { best_pair = c; best_bn = pair_cand[i]; }
  }
  if (best_pair < best_single * 0.97) { *bn = best_bn; return true; }
  return false;
}

template <int BN, int A_MN, int B_MN>
static int launch_gemm(const CUtensorMap& a0, const CUtensorMap& b0, const CUtensorMap& a1, const CUtensorMap& b1,
                       const GemmParams& p, cudaStream_t stream) {
  auto kern = gemm_tcgen05_kernel<BN, A_MN, B_MN>;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [&] {
    attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<BN>::kSmemBytes);
  });
  if (attr_err != cudaSuccess) {
    set_error("cudaFuncSetAttribute(smem=%d) failed: %s", GemmCfg<BN>::kSmemBytes, cudaGetErrorString(attr_err));
    return 3;
  }
  const int tiles = p.m_blocks * p.n_blocks;
  const int grid = tiles < sm_count() ? tiles : sm_count();
  cudaError_t le = launch_pdl(kern, dim3(grid), dim3(kGemmThreads), GemmCfg<BN>::kSmemBytes, stream, a0, b0, a1, b1, p);
  if (le != cudaSuccess) { set_error("mmgl_gemm_bf16: launch failed: %s", cudaGetErrorString(le)); return 1; }
  return check_launch("mmgl_gemm_bf16");
}

template <int BN, int A_MN, int B_MN>
static int launch_gemm_pair(const CUtensorMap& a0, const CUtensorMap& b0, const CUtensorMap& a1, const CUtensorMap& b1,
                            const GemmParams& p, cudaStream_t stream) {
  auto kern = gemm_tcgen05_pair_kernel<BN, A_MN, B_MN>;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [&] {
    attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, PairCfg<BN>::kSmemBytes);
  });
  if (attr_err != cudaSuccess) {
    set_error("cudaFuncSetAttribute(pair, smem=%d) failed: %s", PairCfg<BN>::kSmemBytes, cudaGetErrorString(attr_err));
    return 3;
  }
  const int tiles = p.m_blocks_pair * p.n_blocks;
  const int pairs = sm_count() / 2;
  int clusters = tiles < pairs ? tiles : pairs;
  if (p.sk_split > 1) {
    clusters = p.sk_full > 0 ? pairs : p.sk_tail * p.sk_split;
    if (!p.sk_park_all &&
        cudaMemsetAsync(p.sk_flags, 0, (size_t)p.sk_tail * sizeof(int32_t), stream) != cudaSuccess) {
      set_error("mmgl_gemm_bf16(pair): cudaMemsetAsync of the stream-K flags failed");
      return 3;
    }
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * clusters);
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = PairCfg<BN>::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr; cfg.numAttrs = 2;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, a0, b0, a1, b1, p);
  if (e != cudaSuccess) {
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    set_error("mmgl_gemm_bf16(pair): launch failed: %s", cudaGetErrorString(e));
    return 1;
  }
  const int rc = check_launch("mmgl_gemm_bf16(pair)");
  if (rc != 0 || !p.sk_park_all || p.sk_split <= 1) return rc;
  const int64_t elems = p.m * p.n;
  splitk_reduce_kernel<<<(unsigned)((elems + 255) / 256), 256, 0, stream>>>(p, BN);
  return check_launch("mmgl_gemm_bf16(split-K reduce)");
}

template <int BN>
static int dispatch_major_pair(int a_mn, int b_mn, const CUtensorMap& a0, const CUtensorMap& b0, const CUtensorMap& a1,
                               const CUtensorMap& b1, const GemmParams& p, cudaStream_t s) {
  if (!a_mn && !b_mn) return launch_gemm_pair<BN, 0, 0>(a0, b0, a1, b1, p, s);
  if (!a_mn && b_mn) return launch_gemm_pair<BN, 0, 1>(a0, b0, a1, b1, p, s);
  if (a_mn && !b_mn) return launch_gemm_pair<BN, 1, 0>(a0, b0, a1, b1, p, s);
  return launch_gemm_pair<BN, 1, 1>(a0, b0, a1, b1, p, s);
}

template <int BN>
static int dispatch_major(int a_mn, int b_mn, const CUtensorMap& a0, const CUtensorMap& b0, const CUtensorMap& a1,
                          const CUtensorMap& b1, const GemmParams& p, cudaStream_t s) {
  if (!a_mn && !b_mn) return launch_gemm<BN, 0, 0>(a0, b0, a1, b1, p, s);
  if (!a_mn && b_mn) return launch_gemm<BN, 0, 1>(a0, b0, a1, b1, p, s);
  if (a_mn && !b_mn) return launch_gemm<BN, 1, 0>(a0, b0, a1, b1, p, s);
  return launch_gemm<BN, 1, 1>(a0, b0, a1, b1, p, s);
}

}  // namespace mmgl

using namespace mmgl;

extern "C" size_t mmgl_gemm_workspace_bytes(void) {
  return kSkFlagBytes + (size_t)74 * 256 * 256 * sizeof(float);
}

extern "C" int mmgl_gemm_bf16(const mmgl_gemm_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MMGL_REQUIRE(a != nullptr, "mmgl_gemm_bf16: null args");
  MMGL_BIND(a->d, "mmgl_gemm_bf16");
  MMGL_REQUIRE(a->m > 0 && a->n > 0 && a->k0 > 0 && a->k1 >= 0, "mmgl_gemm_bf16: bad sizes m=%lld n=%lld k0=%lld k1=%lld",
               (long long)a->m, (long long)a->n, (long long)a->k0, (long long)a->k1);
  MMGL_REQUIRE(a->a0 && a->b0 && a->d, "mmgl_gemm_bf16: null operand");
  MMGL_REQUIRE(a->lda0 % 8 == 0 && a->ldb0 % 8 == 0 && aligned16(a->a0) && aligned16(a->b0),
               "mmgl_gemm_bf16: operands need 16-byte aligned base and leading dims %% 8 == 0");
  if (a->k1 > 0)
    MMGL_REQUIRE(a->a1 && a->b1 && a->lda1 % 8 == 0 && a->ldb1 % 8 == 0 && aligned16(a->a1) && aligned16(a->b1),
                 "mmgl_gemm_bf16: second operand pair needs aligned pointers / leading dims");
  MMGL_REQUIRE(a->m < (1ll << 31) && a->n < (1ll << 31), "mmgl_gemm_bf16: m, n must fit in int32");

  const int sms = sm_count();
  int bn = a->force_block_n ? a->force_block_n : pick_block_n(a->m, a->n, sms);
  MMGL_REQUIRE(bn == 64 || bn == 128 || bn == 192 || bn == 256, "mmgl_gemm_bf16: force_block_n must be 64, 128, 192 or 256");

  // CTA-pair kernel: 256 x {128,256} tiles.  pair = 0: heuristic, 1: never, 2: always (tests, tuning)
  bool use_pair = false;
  // Skinny output over a long K (the rank-r LoRA weight gradients: 64 x 768 over 4608 tokens): the data-parallel
  // schedule occupies 6 of 148 SMs; with stream_k == 2 every tile is cut into K-slices and a reduce kernel sums them.
  // Opt-in: alone and L2-cold 30 -> 19.5 us per GEMM, but inside the cfg3 step (operands warm in L2, two launches instead
  // of one) the gain was within run-to-run noise on B200, so the default schedule stays data-parallel.
  const int64_t kblocks_all = (a->k0 + BK - 1) / BK + (a->k1 + BK - 1) / BK;
  const bool skinny = a->pair == 0 && a->force_block_n == 0 && a->stream_k == 2 && a->workspace != nullptr &&
                      ((a->m + 127) / 128) * ((a->n + 127) / 128) * 8 <= sms && kblocks_all >= 32;
  if (a->pair == 2) {
    use_pair = true;
    if (bn != 128) bn = 256;
  } else if (skinny) {
    use_pair = true;
    bn = 128;
  } else if (a->pair == 0 && a->force_block_n == 0) {
    use_pair = pick_pair(a->m, a->n, sms, &bn, a->stream_k == 2 && a->workspace != nullptr);
  }

  GemmParams p;
  p.m = a->m; p.n = a->n;
  p.kblocks0 = (int32_t)((a->k0 + BK - 1) / BK);
  p.kblocks1 = (int32_t)((a->k1 + BK - 1) / BK);
  p.m_blocks = (int32_t)((a->m + BM - 1) / BM);
  p.n_blocks = (int32_t)((a->n + bn - 1) / bn);
  p.m_blocks_pair = (int32_t)((a->m + 255) / 256);
  p.sk_full = 0; p.sk_split = 1; p.sk_tail = 0; p.sk_park_all = 0; p.sk_ws = nullptr; p.sk_flags = nullptr;
  if (use_pair && (a->stream_k == 2 || skinny) && a->workspace != nullptr) {
    const int pairs = sms / 2;
    const int tiles = p.m_blocks_pair * p.n_blocks;
    const int full = (tiles / pairs) * pairs, r = tiles - full;
    const int kblocks = p.kblocks0 + p.kblocks1;
    int split = (r > 0) ? pairs / r : 1;
    if (split > kblocks) split = kblocks;
    if (skinny) {
      if (split > kblocks / 4) split = kblocks / 4;     // at least 4 k-blocks per slice
      if (split > 12) split = 12;
    } else if (split > 8) {
      split = 8;
    }
    const bool park_all = skinny && full == 0;
    const size_t need = kSkFlagBytes + (size_t)r * (split > 1 ? split - (park_all ? 0 : 1) : 0) * 256 * bn * sizeof(float);
    if (split >= 2 && (size_t)a->workspace_bytes >= need && aligned16(a->workspace)) {
      p.sk_full = full; p.sk_split = split; p.sk_tail = r; p.sk_park_all = park_all ? 1 : 0;
      p.sk_flags = reinterpret_cast<int32_t*>(a->workspace);
      p.sk_ws = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(a->workspace) + kSkFlagBytes);
    }
  }
  p.n_fastest = (a->raster == 1) ? 0 : ((a->raster == 2) ? 1 : (p.m_blocks >= p.n_blocks ? 1 : 0));
  p.d = a->d; p.ldd = a->ldd; p.out_fp32 = a->out_fp32; p.accumulate = a->accumulate;
  p.alpha = a->alpha; p.relu = a->relu; p.bias = a->bias; p.gate = a->gate;
  p.residual = reinterpret_cast<const __nv_bfloat16*>(a->residual); p.ldres = a->ldres;
  p.aux = reinterpret_cast<__nv_bfloat16*>(a->aux); p.ldaux = a->ldaux;
  p.relu_mask = reinterpret_cast<const __nv_bfloat16*>(a->relu_mask); p.ldmask = a->ldmask;
  const int esz = a->out_fp32 ? 4 : 2;
  bool vec = aligned16(a->d) && (a->ldd * esz) % 16 == 0;
  if (a->bias) vec = vec && aligned16(a->bias);
  if (a->residual) vec = vec && aligned16(a->residual) && a->ldres % 8 == 0;
  if (a->aux) vec = vec && aligned16(a->aux) && a->ldaux % 8 == 0;
  if (a->relu_mask) vec = vec && aligned16(a->relu_mask) && a->ldmask % 8 == 0;
  p.vec_ok = vec ? 1 : 0;
  static const int staged_env = [] { const char* e = getenv("MMGL_GEMM_STAGED"); return e ? atoi(e) : 1; }();
  p.staged_store = staged_env;
  MMGL_REQUIRE(a->dropout_p >= 0.f && a->dropout_p < 1.f, "mmgl_gemm_bf16: dropout_p must be in [0,1)");
  p.drop_thresh = (uint32_t)(a->dropout_p * 65536.f + 0.5f);
  p.drop_scale = p.drop_thresh ? 65536.f / (65536.f - (float)p.drop_thresh) : 1.f;
  p.drop_seed = a->dropout_seed;
  p.drop_groups = (a->n + 7) / 8;

  CUtensorMap ma0, mb0, ma1, mb1;
  int rc;
  if ((rc = operand_map(&ma0, a->a0, a->a_mn_major, a->m, a->k0, a->lda0, BM))) return rc;
  const int b_box = use_pair ? bn / 2 : bn;
  if ((rc = operand_map(&mb0, a->b0, a->b_mn_major, a->n, a->k0, a->ldb0, b_box))) return rc;
  if (a->k1 > 0) {
    if ((rc = operand_map(&ma1, a->a1, a->a_mn_major, a->m, a->k1, a->lda1, BM))) return rc;
    if ((rc = operand_map(&mb1, a->b1, a->b_mn_major, a->n, a->k1, a->ldb1, b_box))) return rc;
  } else {
    ma1 = ma0; mb1 = mb0;
  }
  if (use_pair) {
    if (bn == 256) return dispatch_major_pair<256>(a->a_mn_major, a->b_mn_major, ma0, mb0, ma1, mb1, p, stream);
    return dispatch_major_pair<128>(a->a_mn_major, a->b_mn_major, ma0, mb0, ma1, mb1, p, stream);
  }
  switch (bn) {
    case 256: return dispatch_major<256>(a->a_mn_major, a->b_mn_major, ma0, mb0, ma1, mb1, p, stream);
    case 192: return dispatch_major<192>(a->a_mn_major, a->b_mn_major, ma0, mb0, ma1, mb1, p, stream);
    case 128: return dispatch_major<128>(a->a_mn_major, a->b_mn_major, ma0, mb0, ma1, mb1, p, stream);
    default:  return dispatch_major<64>(a->a_mn_major, a->b_mn_major, ma0, mb0, ma1, mb1, p, stream);
  }
}
