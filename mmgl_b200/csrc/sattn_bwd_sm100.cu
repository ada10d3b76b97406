// Self-attention backward, ONE pass: dQ, dK and dV of a (sample, head) from a single
// recomputation of S = Q K^T and dP = dO V^T per 128 x 128 block -- five tcgen05 MMAs and one softmax-gradient pass per
// block (the round-1 pair of kernels formed S, dP and the softmax gradient twice: seven MMAs, two passes).
//
// Arithmetic restated: the autograd backward of MPTAttention's self branch (model/modelling_cross_attention.py:201-275
// with the mask of :455-476) and of the HF T5 / OPT attention the concat path runs (model/modelling_self_attention.py:332):
//     P = softmax(max(scale * Q K^T + bias + mask, finfo.min)),  P~ = dropout(P)
//     dV = P~^T dO,   dP = dO V^T,   dS = P . (keep/(1-p) . dP - rowsum(dO . O)),   dQ = scale * dS K,   dK = scale * dS^T Q
//
// Work split: a persistent CTA OWNS a whole (sample, head): it walks the key blocks j (outer) and the query tiles i that see
// them (inner), so
//   * dK_j / dV_j accumulate in TMEM over the inner loop and are written once per key block;
//   * the dQ_i contributions of successive key blocks come from the SAME thread in program order -> they are summed in a
//     per-CTA fp32 scratch (L2-resident: grid x seq_q x head_dim floats) with plain st / ld + add + st in a fixed order:
//     deterministic, no atomics, no zero-fill pass, no fp32 -> bf16 conversion kernel; the last key block of a tile adds
//     its part, scales and stores bf16 dQ directly;
//   * delta = rowsum(dO . O) comes from one coalesced pre-pass (sattn_delta_kernel); the softmax threads fetch it together
//     with the row statistics one step ahead.
// Roles: 8 (or 16) softmax-gradient warps (thread = one score row x 32 (16) of the 64 keys of a half block), one
// MMA-issuing warp, one TMA warp that streams K_j / V_j and Q_i / dO_i tiles through double-buffered rings (head_dim 64)
// across step and item boundaries.  S / dP live in TMEM as two 64-key halves, so the tensor core refills a half as soon as
// the threads have pulled it into registers and runs S / dP of the NEXT step while this step's exponentials execute.
// TMEM columns: head_dim 64: 2 x (S 64 + dP 64) + dK 64 + dV 64 + 2 x dQ 64 = 512; head_dim 128: (64 + 64) + 128 + 128 + 128 = 512.
#include "sattn_common.cuh"

namespace mmgl {
namespace {

template <int D>
struct BwdCfg {
  static constexpr int TB = (D / 64) * 16384;        // bytes of one [128][D] bf16 tile
  static constexpr int NKV = (D == 64) ? 2 : 1;      // K / V buffers
  static constexpr int NQ = (D == 64) ? 2 : 1;       // Q / dO buffers
  static constexpr int NSB = (D == 64) ? 2 : 1;      // S / dP half buffers in TMEM
  static constexpr int NDQ = (D == 64) ? 2 : 1;      // dQ-part buffers in TMEM (2: the fold into the scratch overlaps the next MMA group)
  static constexpr uint32_t cS = 0;                  // + buf * 128: S half [128 x 64], then dP half at + 64
  static constexpr uint32_t cdK = NSB * 128;
  static constexpr uint32_t cdV = cdK + D;
  static constexpr uint32_t cdQ = cdV + D;
  // barriers
  static constexpr int kvfull = 0, kvfree = NKV, qfull = 2 * NKV, qfree = 2 * NKV + NQ, sfull = 2 * NKV + 2 * NQ,
                       sfree = sfull + NSB, pfull = sfree + NSB, mma2done = pfull + 1, nbars = mma2done + 1;
};

// (item, key block, query tile) enumeration shared by the three roles: every role walks the same sequence of steps.
// Only the moving parts live in registers; the problem constants are re-read from the kernel parameters (constant bank).
struct StepIter {
  int item, j, i;
  int sc, jc;          // running step / key-block counters of this CTA (buffer indices and barrier phases derive from them)
  static __device__ __forceinline__ int nbk(const AttnParams& p) { return (p.seq_k + 127) >> 7; }
  static __device__ __forceinline__ int ntq(const AttnParams& p) { return (p.seq_q + 127) >> 7; }
  static __device__ __forceinline__ int first_tile(const AttnParams& p, int jj) { return p.causal ? max(0, jj * 128 - p.coff) >> 7 : 0; }
  static __device__ __forceinline__ int blocks_of(const AttnParams& p, int ii) {
    return p.causal ? min(nbk(p), ((ii * 128 + 127 + p.coff) >> 7) + 1) : nbk(p);
  }
  __device__ __forceinline__ void init(int first_item) { item = first_item; sc = 0; jc = 0; j = 0; i = 0; }
  __device__ __forceinline__ bool done(const AttnParams& p) const { return item >= p.batch * p.heads; }
  __device__ __forceinline__ int b(const AttnParams& p) const { return item / p.heads; }
  __device__ __forceinline__ int h(const AttnParams& p) const { return item % p.heads; }
  __device__ __forceinline__ bool first_of_j(const AttnParams& p) const { return i == first_tile(p, j); }
  __device__ __forceinline__ bool last_of_j(const AttnParams& p) const { return i == ntq(p) - 1; }
  __device__ __forceinline__ bool first_of_item() const { return j == 0 && i == 0; }
  __device__ __forceinline__ void next(const AttnParams& p, int stride) {
    ++sc;
    if (++i < ntq(p)) return;
    ++jc;
    if (++j < nbk(p)) { i = first_tile(p, j); return; }
    item += stride; j = 0; i = 0;
  }
};

// keep bits of 16 consecutive keys starting at key0 (multiple of 16) of dropout row `drow`
__device__ __forceinline__ uint32_t keep_half_word(const AttnParams& p, int64_t drow, int key0, int64_t groups_per_row) {
  uint32_t w = 0;
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    const DropBits bits = dropout_bits(p.drop_seed, drow, (key0 >> 3) + g, groups_per_row);
#pragma unroll
    for (int e = 0; e < 8; ++e) w |= dropout_keep(bits, e, p.drop_thresh) ? (1u << (8 * g + e)) : 0u;
  }
  return w;
}

__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr) : "memory");
}

// delta[b, h, row] = rowsum(dO . O) over the head's D columns: one coalesced pass over O and dO ahead of the main kernel (a
// warp covers 256 columns of one row, 8 bf16 per lane; D / 8 neighbouring lanes hold one head).  Forming it inside the main
// kernel (global loads by the softmax threads in the j = 0 pass, partial sums meeting in shared memory behind a block
// barrier) cost ~4500 cycles on each of those steps -- a sixth of the kernel -- against ~13 us for this pass at batch 16.
template <int D>
__global__ void __launch_bounds__(256)
sattn_delta_kernel(const __nv_bfloat16* __restrict__ o, int64_t ldo, const __nv_bfloat16* __restrict__ d_o, int64_t lddo,
                   float* __restrict__ delta, int64_t rows, int seq_q, int heads) {
  pdl_launch();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int chunks = (heads * D + 255) / 256;
  const int64_t wid = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (wid >= rows * chunks) return;
  const int64_t row = wid / chunks;
  const int col = (int)(wid % chunks) * 256 + lane * 8;
  float acc = 0.f;
  if (col < heads * D) {
    const uint4 vo = __ldg(reinterpret_cast<const uint4*>(o + row * ldo + col));
    const uint4 vd = __ldg(reinterpret_cast<const uint4*>(d_o + row * lddo + col));
    const uint32_t wo[4] = {vo.x, vo.y, vo.z, vo.w}, wd[4] = {vd.x, vd.y, vd.z, vd.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) acc += bf16lo(wo[e]) * bf16lo(wd[e]) + bf16hi(wo[e]) * bf16hi(wd[e]);
  }
#pragma unroll
  for (int off = 1; off < D / 8; off <<= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if (col < heads * D && (lane % (D / 8)) == 0) {
    const int64_t b = row / seq_q, r = row % seq_q;
    delta[(b * heads + col / D) * seq_q + r] = acc;
  }
}

// what the softmax threads need to finish a step after its second group of MMAs has completed
struct PrevStep {
  int flags;          // kValid | kRowOk | kFirst (key block 0) | kLast (last key block of the query tile) | kLastOfJ
  int row, j, sc, b, h;
};
constexpr int kValid = 1, kRowOk = 2, kFirst = 4, kLast = 8, kLastOfJ = 16;

// TPR = softmax threads per score row (each owns 64 / TPR keys of every 64-key half block).  4 x TPR softmax warps + the MMA
// warp + the TMA warp.  TPR = 2 (320 threads, 168 registers: no spills, half the per-warp bookkeeping) is the default at
// head_dim 64; TPR = 4 (576 threads, 96 registers) at head_dim 128.  MMGL_SATTN_TPR overrides (experiments).
template <int D, bool kBias, bool kDrop, int TPR>
__global__ void __launch_bounds__(128 * TPR + 64, 1)
sattn_bwd_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_do,
                 const __grid_constant__ CUtensorMap map_k, const __grid_constant__ CUtensorMap map_v,
                 const __grid_constant__ AttnParams p, const __nv_bfloat16* __restrict__ o, int64_t ldo,
                 const __nv_bfloat16* __restrict__ d_o, int64_t lddo, const float* __restrict__ stats,
                 __nv_bfloat16* __restrict__ dq, int64_t lddq, __nv_bfloat16* __restrict__ dk, int64_t lddk,
                 __nv_bfloat16* __restrict__ dv, int64_t lddv, float* __restrict__ dq_ws, const float* __restrict__ delta_ws) {
  using C = BwdCfg<D>;
  constexpr int TB = C::TB;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sK = smem;                       // NKV tiles
  uint8_t* sV = sK + C::NKV * TB;
  uint8_t* sQ = sV + C::NKV * TB;           // NQ tiles
  uint8_t* sdO = sQ + C::NQ * TB;
  uint8_t* sP = sdO + C::NQ * TB;           // [128 q][128 keys] bf16 as two 64-key slabs
  uint8_t* sdS = sP + 32768;
  const int nbk = (p.seq_k + 127) / 128, ntq = (p.seq_q + 127) / 128;
  float* sBins = reinterpret_cast<float*>(sdS + 32768);      // [256] diagonal sums of dS (gradient of the relative-position bias)
  uint32_t* kbits = reinterpret_cast<uint32_t*>(sBins + 256); // [2][4 * nbk] attend bits of the item's sample (double-buffered)
  uint64_t* bars = reinterpret_cast<uint64_t*>(kbits + 2 * (4 * nbk + ((4 * nbk) & 1)));
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + C::nbars);
  const int kb_stride = 4 * nbk + ((4 * nbk) & 1);

  // warp index / TMEM base through a lane-0 broadcast: ptxas then treats them as warp-uniform (see the MMA issuer)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31, tid = threadIdx.x;
  if (tid == 0) {
    tma_prefetch_desc(&map_q); tma_prefetch_desc(&map_do); tma_prefetch_desc(&map_k); tma_prefetch_desc(&map_v);
    for (int i = 0; i < C::nbars; ++i) {
      const bool wide = (i >= C::sfree && i < C::sfree + C::NSB) || i == C::pfull;
      mbar_init(&bars[i], wide ? 128 * TPR : 1);
    }
    fence_barrier_init();
  }
  constexpr int kMmaWarp = 4 * TPR, kTmaWarp = 4 * TPR + 1, kSoftmaxThreads = 128 * TPR;
  constexpr int KPT = 64 / TPR;             // keys per thread per half block
  if (warp == kMmaWarp) tmem_alloc<512>(tmem_ptr);
  if (tid < 256) sBins[tid] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);
  pdl_launch();
  pdl_wait();   // the set-up above overlapped the previous kernel's tail (the delta pre-pass, usually)

  if (warp == kTmaWarp) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      StepIter it;
      it.init(blockIdx.x);
      int tri = 0; (void)tri;
      while (!it.done(p)) {
        const int colq = it.h(p) * D, rowq = it.b(p) * p.seq_q, rowk = it.b(p) * p.seq_k;
        if (it.first_of_j(p)) {
          const int kb = it.jc % C::NKV;
          if (it.jc >= C::NKV) mbar_wait(&bars[C::kvfree + kb], ((it.jc / C::NKV) - 1) & 1);
          mbar_arrive_expect_tx(&bars[C::kvfull + kb], 2 * TB);
          tma_tile<D>(sK + kb * TB, &map_k, &bars[C::kvfull + kb], colq, rowk + it.j * 128);
          tma_tile<D>(sV + kb * TB, &map_v, &bars[C::kvfull + kb], colq, rowk + it.j * 128);
        }
        const int qb = it.sc % C::NQ;
        if (it.sc >= C::NQ) mbar_wait(&bars[C::qfree + qb], ((it.sc / C::NQ) - 1) & 1);
        mbar_arrive_expect_tx(&bars[C::qfull + qb], 2 * TB);
        tma_tile<D>(sQ + qb * TB, &map_q, &bars[C::qfull + qb], colq, rowq + it.i * 128);
        tma_tile<D>(sdO + qb * TB, &map_do, &bars[C::qfull + qb], colq, rowq + it.i * 128);
        TR(3, tri, 1000 + it.sc);
        it.next(p, gridDim.x);
      }
    }
    __syncwarp();
  } else if (warp == kMmaWarp) {
    // ------------------------------------------------------------------ MMA issuer
    // The whole warp walks the schedule (barrier waits included) and one elected lane issues: every operand then derives
    // from warp-uniform values, the descriptors stay in uniform registers and the UTCHMMAs of a group go out back to back
    // (under `if (lane == 0)` ptxas wraps each MMA in an ELECT / R2UR waterfall: ~100 cycles of issue latency per MMA).
    {
      const uint64_t desc_q = make_smem_desc(smem_u32(sQ), 16, 1024), desc_do = make_smem_desc(smem_u32(sdO), 16, 1024);
      const uint64_t desc_k = make_smem_desc(smem_u32(sK), 16, 1024), desc_v = make_smem_desc(smem_u32(sV), 16, 1024);
      const uint64_t desc_ds = make_smem_desc(smem_u32(sdS), 16, 1024);
      const uint64_t desc_k_mn = make_smem_desc(smem_u32(sK), 16384, 1024);
      const uint64_t desc_p_mn = make_smem_desc(smem_u32(sP), 16384, 1024), desc_ds_mn = make_smem_desc(smem_u32(sdS), 16384, 1024);
      const uint64_t desc_q_mn = make_smem_desc(smem_u32(sQ), 16384, 1024), desc_do_mn = make_smem_desc(smem_u32(sdO), 16384, 1024);
      StepIter cur, la;                 // cur: the step whose dV / dK / dQ MMAs are next; la: the step whose S / dP halves are next
      cur.init(blockIdx.x);
      la.init(blockIdx.x);
      int la_half = 0;                  // next half (0 / 1) of step `la` to issue
      int tri = 0; (void)tri;
      // S / dP of half x = 2 * la.sc + la_half into TMEM buffer x % NSB
      auto issue_s = [&]() {
        const int x = 2 * la.sc + la_half, sb = x % C::NSB;
        const int qb = la.sc % C::NQ, kb = la.jc % C::NKV;
        if (la_half == 0) {
          if (la.first_of_j(p)) mbar_wait(&bars[C::kvfull + kb], (la.jc / C::NKV) & 1);
          mbar_wait(&bars[C::qfull + qb], (la.sc / C::NQ) & 1);
        }
        if (x >= C::NSB) mbar_wait(&bars[C::sfree + sb], ((x / C::NSB) - 1) & 1);
        tc_fence_after();
        const uint64_t qoff = (uint64_t)((qb * TB) >> 4), koff = (uint64_t)((kb * TB + la_half * 8192) >> 4);
        if (elect_one()) {
          mma_qk_half<D>(tmem_base + C::cS + sb * 128, desc_q + qoff, desc_k + koff);          // S  = Q_i  K_j[half]^T
          mma_qk_half<D>(tmem_base + C::cS + sb * 128 + 64, desc_do + qoff, desc_v + koff);    // dP = dO_i V_j[half]^T
          umma_commit(&bars[C::sfull + sb]);
          TR(2, tri, 2000 + x);
        }
        __syncwarp();
        if (la_half == 1) la.next(p, gridDim.x);
        la_half ^= 1;
      };
      while (!cur.done(p)) {
        // run S / dP ahead: up to the halves of the NEXT step when its Q / dO tiles have their own buffer (NQ == 2),
        // otherwise to the end of this step (the single Q / dO buffer is released by this step's second MMA group)
        const int x_max = (C::NQ >= 2) ? 2 * cur.sc + 3 : 2 * cur.sc + 1;
        while (!la.done(p) && 2 * la.sc + la_half <= x_max) issue_s();
        const int qb = cur.sc % C::NQ, kb = cur.jc % C::NKV;
        mbar_wait(&bars[C::pfull], cur.sc & 1);
        tc_fence_after();
        const uint64_t qoff = (uint64_t)((qb * TB) >> 4), koff = (uint64_t)((kb * TB) >> 4);
        const bool acc = !cur.first_of_j(p);
        if (elect_one()) {
          TR(2, tri, 3000 + cur.sc);
          mma_tn_desc<D>(tmem_base + C::cdV, desc_p_mn, desc_do_mn + qoff, acc);     // dV_j += P~^T dO_i
          mma_tn_desc<D>(tmem_base + C::cdK, desc_ds_mn, desc_q_mn + qoff, acc);     // dK_j += dS^T Q_i
          mma_pv_desc<D>(tmem_base + C::cdQ + (cur.sc % C::NDQ) * D, desc_ds, desc_k_mn + koff, false);   // dQ part = dS K_j
          umma_commit(&bars[C::mma2done]);
          umma_commit(&bars[C::qfree + qb]);
          if (cur.last_of_j(p)) umma_commit(&bars[C::kvfree + kb]);
          TR(2, tri, 4000 + cur.sc);
        }
        __syncwarp();
        cur.next(p, gridDim.x);
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ softmax-gradient warps (128 x TPR threads)
    const int rit = tid & 127, quarter = tid >> 7;           // row in tile; KPT-key column group of every 64-key half
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    float* ws = dq_ws + (size_t)blockIdx.x * ((size_t)ntq * 128 * D);
    const bool want_dbias = kBias && p.d_rel_bias != nullptr;
    const int64_t dgroups = (p.seq_k + 7) >> 3;
    PrevStep prev;
    prev.flags = 0;

    // finish a step once its dV / dK / dQ MMAs are complete: fold the dQ part into the scratch (or emit dQ), and after
    // the last query tile of a key block write dK_j / dV_j
    auto finish_prev = [&]() {
      const bool row_ok = prev.flags & kRowOk, first = prev.flags & kFirst, last = prev.flags & kLast;
      const int colq = prev.h * D;
      {
        constexpr int NC = D / TPR;                                  // this thread's columns of the dQ row
        // scratch layout [query tile][D / 4 float4 groups][128 rows][4]: the 32 lanes of a warp (32 consecutive rows, same
        // group) touch 512 contiguous bytes per instruction (row-major [row][D] made every access 32 scattered 16-byte pieces)
        float* acc = ws + ((size_t)(prev.row >> 7) * (D / 4) + quarter * (NC / 4)) * 512 + (size_t)(prev.row & 127) * 4;
#pragma unroll
        for (int c0 = 0; c0 < NC; c0 += 16) {
          // the running sum of the earlier key blocks was written by THIS thread (program order, no atomics needed):
          // request it first, then pull the new part out of TMEM.  (red.global.add.v4.f32 was measured as well: same speed.)
          float4 a[4];
          if (row_ok && !first) {
#pragma unroll
            for (int g = 0; g < 4; ++g)
              asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(a[g].x), "=f"(a[g].y), "=f"(a[g].z), "=f"(a[g].w)
                           : "l"(acc + (c0 / 4 + g) * 512) : "memory");
          }
          uint32_t r[16];
          tmem_ld_32x16(lane_addr + C::cdQ + (prev.sc % C::NDQ) * D + quarter * NC + c0, r);   // warp-collective: outside the row predicate
          tmem_ld_wait();
          if (!row_ok) continue;
          float f[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) f[e] = __uint_as_float(r[e]);
          if (!first) {
#pragma unroll
            for (int g = 0; g < 4; ++g) { f[4 * g] += a[g].x; f[4 * g + 1] += a[g].y; f[4 * g + 2] += a[g].z; f[4 * g + 3] += a[g].w; }
          }
          if (last) {
            __nv_bfloat16* dst = dq + ((int64_t)prev.b * p.seq_q + prev.row) * lddq + colq + quarter * NC;
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              uint4 v;
              v.x = pack_bf16(f[8 * g] * p.scale, f[8 * g + 1] * p.scale);
              v.y = pack_bf16(f[8 * g + 2] * p.scale, f[8 * g + 3] * p.scale);
              v.z = pack_bf16(f[8 * g + 4] * p.scale, f[8 * g + 5] * p.scale);
              v.w = pack_bf16(f[8 * g + 6] * p.scale, f[8 * g + 7] * p.scale);
              *reinterpret_cast<uint4*>(dst + c0 + 8 * g) = v;
            }
          } else {
#pragma unroll
            for (int g = 0; g < 4; ++g)
              asm volatile("st.global.cg.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(acc + (c0 / 4 + g) * 512), "f"(f[4 * g]), "f"(f[4 * g + 1]),
                           "f"(f[4 * g + 2]), "f"(f[4 * g + 3]) : "memory");
          }
        }
      }
      if (prev.flags & kLastOfJ) {
        // the lower half of the column groups stores dV, the upper half dK; each stores 2 D / TPR columns of its key row
        const int key = prev.j * 128 + rit;
        const bool is_k = quarter >= TPR / 2;
        const int part = quarter % (TPR / 2);
        constexpr int EC = 2 * D / TPR;
        const uint32_t src = lane_addr + (is_k ? C::cdK : C::cdV) + part * EC;
        const int64_t grow = (int64_t)prev.b * p.seq_k + key;
        __nv_bfloat16* dst = (is_k ? dk + grow * lddk : dv + grow * lddv) + colq + part * EC;
        const float mul = is_k ? p.scale : 1.f;
#pragma unroll
        for (int c = 0; c < EC / 32; ++c) {
          uint32_t r[32];
          tmem_ld_32x32(src + c * 32, r);
          tmem_ld_wait();
          if (key < p.seq_k) store_row_bf16(dst + c * 32, r, mul);
        }
      }
      tc_fence_before();
      prev.flags = 0;
    };

    StepIter it;
    it.init(blockIdx.x);
    int tri = 0; (void)tri;
    float nxt_m = 0.f, nxt_inv = 0.f, nxt_delta = 0.f;
    if (!it.done(p) && rit < p.seq_q) {
      const int64_t sidx = (int64_t)it.item * p.seq_q + rit;
      const float2 st = __ldg(reinterpret_cast<const float2*>(stats + sidx * 2));
      nxt_m = st.x; nxt_inv = st.y;
      nxt_delta = __ldg(delta_ws + sidx);
    }
    int item_count = 0;
    const uint32_t* kb_cur = kbits;
    while (!it.done(p)) {
      const int ib = it.b(p), ih = it.h(p);
      if (it.first_of_item()) {
        // attend bits of this item's sample (other items' readers may still use the other buffer)
        uint32_t* kb_w = kbits + (item_count & 1) * kb_stride;
        for (int w = warp; w < 4 * nbk; w += 4 * TPR) {
          const int key = w * 32 + lane;
          const bool a = key < p.seq_k && (p.key_mask == nullptr || p.key_mask[(int64_t)ib * p.seq_k + key] != 0);
          const uint32_t bits = __ballot_sync(0xffffffffu, a);
          if (lane == 0) kb_w[w] = bits;
        }
        kb_cur = kb_w;
        ++item_count;
        asm volatile("bar.sync 1, %0;" ::"n"(kSoftmaxThreads) : "memory");
      }
      const int i = it.i, j = it.j;
      if (tid == 0) TR(0, tri, 100000 + it.sc * 100);
      const int row = i * 128 + rit;
      const bool row_ok = row < p.seq_q;
      // row statistics (m, 1 / l) and delta of this step were requested a step ago; request the next step's now
      const float m = nxt_m, inv = nxt_inv, delta = nxt_delta;
      {
        StepIter nx = it;
        nx.next(p, gridDim.x);
        const int nrow = nx.i * 128 + rit;
        nxt_m = 0.f; nxt_inv = 0.f; nxt_delta = 0.f;
        if (!nx.done(p) && nrow < p.seq_q) {
          const int64_t sidx = (int64_t)nx.item * p.seq_q + nrow;
          const float2 st = __ldg(reinterpret_cast<const float2*>(stats + sidx * 2));
          nxt_m = st.x; nxt_inv = st.y;
          nxt_delta = __ldg(delta_ws + sidx);
        }
      }
      const bool none = !(m > -FLT_MAX);
      const bool flat = none || !row_ok;       // no attended key (uniform row) or a row beyond the sequence (p = 0)
      const float c1 = flat ? 0.f : p.scale * kL2E, mc = flat ? 0.f : m * kL2E, bsc = flat ? 0.f : kL2E;
      const float inv_ok = row_ok ? inv : 0.f;
      const float* bias_row = kBias ? p.rel_bias + (int64_t)ih * (p.seq_q + p.seq_k - 1) + (p.seq_q - 1 - min(row, p.seq_q - 1)) : nullptr;
      const int64_t drow = (int64_t)it.item * p.seq_q + row;
      const uint32_t p_base = smem_u32(sP), ds_base = smem_u32(sdS);

#pragma unroll 1
      for (int hk = 0; hk < 2; ++hk) {
        const int x = 2 * it.sc + hk, sb = x % C::NSB;
        const int koff = hk * 64 + quarter * KPT;          // first key of this thread's group inside the block
        const int key0 = j * 128 + koff;
        // attend bits of the KPT keys for this row
        uint32_t mw = kb_cur[4 * j + (koff >> 5)];
        if (KPT == 16) mw = (mw >> (koff & 16)) & 0xffffu;
        if (p.causal) mw &= low_bits(row + p.coff - key0 + 1);
        if (none) mw = low_bits(p.seq_k - key0) & (KPT == 16 ? 0xffffu : 0xffffffffu);   // uniform over the existing keys of the visited blocks
        const int lim = max(p.seq_k - 1 - key0, 0);
        const float* bk = kBias ? bias_row + min(key0, p.seq_k - 1) : nullptr;
        const float ks = kDrop ? p.drop_scale : 1.f;

        mbar_wait(&bars[C::sfull + sb], (x / C::NSB) & 1);
        if (tid == 0) TR(0, tri, 100000 + it.sc * 100 + 10 + hk);
        tc_fence_after();
        uint32_t pk[KPT / 2], dk_[KPT / 2];
#pragma unroll
        for (int c = 0; c < KPT; c += 16) {     // 16 keys at a time: S and dP of the group, then P and dS
          uint32_t rs[16], rp[16];
          tmem_ld_32x16(lane_addr + C::cS + sb * 128 + quarter * KPT + c, rs);
          tmem_ld_32x16(lane_addr + C::cS + sb * 128 + 64 + quarter * KPT + c, rp);
          tmem_ld_wait();
          if (c + 16 == KPT) {                  // the whole group is in registers: the tensor core may refill this half
            tc_fence_before();
            mbar_arrive(&bars[C::sfree + sb]);
          }
          uint32_t keep = 0xffffu;
          if (kDrop) keep = keep_half_word(p, drow, key0 + c, dgroups);
#pragma unroll
          for (int e = 0; e < 16; e += 2) {
            float pv[2], dv_[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              float off = -mc;
              if (kBias) off = fmaf(__ldg(bk + min(c + e + u, lim)), bsc, -mc);
              float pr = ex2(fmaf(__uint_as_float(rs[e + u]), c1, off)) * inv_ok;
              pr = (mw >> (c + e + u)) & 1u ? pr : 0.f;
              if (kDrop) {
                const float kmul = (keep >> (e + u)) & 1u ? ks : 0.f;
                pv[u] = pr * kmul;
                dv_[u] = pr * fmaf(__uint_as_float(rp[e + u]), kmul, -delta);
              } else {
                pv[u] = pr;
                dv_[u] = pr * (__uint_as_float(rp[e + u]) - delta);
              }
            }
            pk[(c + e) >> 1] = pack_bf16(pv[0], pv[1]);
            dk_[(c + e) >> 1] = pack_bf16(dv_[0], dv_[1]);
            if (kBias && want_dbias) {   // d bias(key - row) += dS: one bin per diagonal of the 128 x 128 block
              atomicAdd(sBins + (koff + c + e - rit + 127), dv_[0]);
              atomicAdd(sBins + (koff + c + e + 1 - rit + 127), dv_[1]);
            }
          }
        }
        if (tid == 0) TR(0, tri, 100000 + it.sc * 100 + 20 + hk);
        if (hk == 0 && prev.flags) {
          // the previous step's dV / dK / dQ MMAs have read P / dS (and written the dQ part): the tiles can be reused.  Its
          // dQ part is folded into the scratch AFTER this step's P / dS are published (second dQ buffer: the fold overlaps
          // this step's MMA group) -- unless there is a single dQ buffer, or the step closed a key block, whose dK / dV
          // must leave TMEM before this step's MMAs restart the accumulation
          mbar_wait(&bars[C::mma2done], (it.sc - 1) & 1);
          if (tid == 0) TR(0, tri, 100000 + it.sc * 100 + 30);
          tc_fence_after();
          if (C::NDQ == 1 || (prev.flags & kLastOfJ)) finish_prev();
        }
        // KPT keys = KPT / 8 16-byte chunks of row `rit` in the half's 64-key slab
#pragma unroll
        for (int g = 0; g < KPT / 8; ++g) {
          sts128(swz(p_base + hk * 16384, rit, (KPT / 8) * quarter + g), pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]);
          sts128(swz(ds_base + hk * 16384, rit, (KPT / 8) * quarter + g), dk_[4 * g], dk_[4 * g + 1], dk_[4 * g + 2], dk_[4 * g + 3]);
        }
      }
      if (want_dbias) {   // flush this block's 255 diagonals: bin t holds key - row = j * 128 - i * 128 + t - 127
        asm volatile("bar.sync 1, %0;" ::"n"(kSoftmaxThreads) : "memory");
        if (tid < 255) {
          const float v = sBins[tid];
          const int idx = j * 128 - i * 128 + tid - 127 + p.seq_q - 1;
          if (v != 0.f && idx >= 0 && idx < p.seq_q + p.seq_k - 1)
            atomicAdd(p.d_rel_bias + (int64_t)ih * (p.seq_q + p.seq_k - 1) + idx, v);
          sBins[tid] = 0.f;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kSoftmaxThreads) : "memory");
      }
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(&bars[C::pfull]);
      if (tid == 0) TR(0, tri, 100000 + it.sc * 100 + 50);
      if (prev.flags) finish_prev();     // deferred fold of the previous step (overlaps this step's MMA group)
      if (tid == 0) TR(0, tri, 100000 + it.sc * 100 + 60);
      prev.flags = kValid | (row_ok ? kRowOk : 0) | (j == 0 ? kFirst : 0) | (j == StepIter::blocks_of(p, i) - 1 ? kLast : 0) |
                   (it.last_of_j(p) ? kLastOfJ : 0);
      prev.row = row; prev.j = j; prev.sc = it.sc; prev.b = ib; prev.h = ih;
      it.next(p, gridDim.x);
    }
    if (prev.flags) {
      mbar_wait(&bars[C::mma2done], (it.sc - 1) & 1);
      tc_fence_after();
      finish_prev();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc<512>(tmem_base);
}

template <int D>
size_t bwd_smem_bytes(const AttnParams& p) {
  using C = BwdCfg<D>;
  const size_t nbk = (size_t)((p.seq_k + 127) / 128), ntq = (size_t)((p.seq_q + 127) / 128);
  const size_t kb_stride = 4 * nbk + ((4 * nbk) & 1);
  (void)ntq;
  return (size_t)(2 * C::NKV + 2 * C::NQ) * C::TB + 65536 + 1024 + 2 * kb_stride * 4 + C::nbars * 8 + 16;
}

template <int D, bool kBias, bool kDrop>
int launch_bwd_v(const Maps& mp, const AttnParams& p, const void* o, int64_t ldo, const void* d_o, int64_t lddo,
                 const float* stats, void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv, float* ws,
                 size_t ws_bytes, cudaStream_t stream) {
  const size_t smem = bwd_smem_bytes<D>(p);
  MMGL_REQUIRE(smem <= 227 * 1024, "mmgl_attn_bwd: seq_q = %d needs %zu bytes of shared memory (row statistics of a whole (sample, head) "
               "are kept on chip); supported up to about %d queries at head_dim %d", p.seq_q, smem, D == 64 ? 8192 : 8192, D);
  // measured on B200 (tools/attn_bench_small.py): head_dim 64, cfg2 layer shape at batch 16: 262 us with TPR 2 vs 326 us with
  // TPR 4; head_dim 128 (single-buffered tiles, serial halves), cfg5 layer shape: 403 us vs 367 us
  static const int tpr_env = [] { const char* e = getenv("MMGL_SATTN_TPR"); return e ? (e[0] == '4' ? 4 : 2) : 0; }();
  const int tpr = tpr_env ? tpr_env : (D == 64 ? 2 : 4);
  auto kern = tpr == 4 ? sattn_bwd_kernel<D, kBias, kDrop, 4> : sattn_bwd_kernel<D, kBias, kDrop, 2>;
  MMGL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t items = (int64_t)p.heads * p.batch;
  const int64_t grid = items < sm_count() ? items : sm_count();
  // workspace = [delta: batch * heads * seq_q floats, padded to 256 B][dQ scratch: one tile set per CTA]
  const size_t delta_bytes = (((size_t)items * p.seq_q * sizeof(float)) + 255) & ~(size_t)255;
  const size_t need = delta_bytes + (size_t)grid * (size_t)((p.seq_q + 127) / 128) * 128 * D * sizeof(float);
  MMGL_REQUIRE(ws_bytes >= need, "mmgl_attn_bwd: workspace too small (%zu < %zu bytes; use mmgl_attn_bwd_workspace_bytes)", ws_bytes, need);
  float* delta = ws;
  float* dq_ws = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(ws) + delta_bytes);
  {
    const int64_t rows = (int64_t)p.batch * p.seq_q;
    const int64_t warps = rows * ((p.heads * D + 255) / 256);
    MMGL_CUDA(launch_pdl(sattn_delta_kernel<D>, dim3((unsigned)((warps + 7) / 8)), dim3(256), 0, stream, (const __nv_bfloat16*)o, ldo,
                         (const __nv_bfloat16*)d_o, lddo, delta, rows, p.seq_q, p.heads));
    if (int rc = check_launch("mmgl_attn_bwd(delta)")) return rc;
  }
  MMGL_CUDA(launch_pdl(kern, dim3((unsigned)grid), dim3(128 * tpr + 64), smem, stream, mp.q, mp.d_o, mp.k, mp.v, p,
                       (const __nv_bfloat16*)o, ldo, (const __nv_bfloat16*)d_o, lddo, stats, (__nv_bfloat16*)dq, lddq,
                       (__nv_bfloat16*)dk, lddk, (__nv_bfloat16*)dv, lddv, dq_ws, delta));
  return check_launch("mmgl_attn_bwd");
}

template <int D>
int launch_bwd(const Maps& mp, const AttnParams& p, const void* o, int64_t ldo, const void* d_o, int64_t lddo,
               const float* stats, void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv, float* ws,
               size_t ws_bytes, cudaStream_t stream) {
  const bool bias = p.rel_bias != nullptr, drop = p.drop_thresh != 0;
#define MMGL_BWD(B_, R_) launch_bwd_v<D, B_, R_>(mp, p, o, ldo, d_o, lddo, stats, dq, lddq, dk, lddk, dv, lddv, ws, ws_bytes, stream)
  if (bias) return drop ? MMGL_BWD(true, true) : MMGL_BWD(true, false);
  return drop ? MMGL_BWD(false, true) : MMGL_BWD(false, false);
#undef MMGL_BWD
}

}  // namespace
}  // namespace mmgl

using namespace mmgl;

#ifdef MMGL_TRACE
extern "C" int mmgl_debug_trace_bwd(unsigned long long* host_dst) {
  cudaDeviceSynchronize();
  return (int)cudaMemcpyFromSymbol(host_dst, g_trace, sizeof(g_trace));
}
#endif

// fp32 scratch: rowsum(dO . O) per (sample, head, query) + the per-CTA dQ accumulation, one [ceil(seq_q / 128) * 128, head_dim]
// tile set per resident CTA (an upper bound that does not depend on the caller knowing head_dim: 128 columns, one CTA per
// SM, at most batch * heads CTAs)
extern "C" size_t mmgl_attn_bwd_workspace_bytes(int64_t batch, int64_t seq_q, int64_t heads) {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t ctas = batch * heads < sms ? batch * heads : sms;
  const size_t delta_bytes = (((size_t)(batch * heads * seq_q) * sizeof(float)) + 255) & ~(size_t)255;   // rowsum(dO . O)
  return delta_bytes + (size_t)ctas * (size_t)((seq_q + 127) / 128) * 128 * 128 * sizeof(float);
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
  MMGL_REQUIRE(workspace != nullptr, "mmgl_attn_bwd: workspace required (mmgl_attn_bwd_workspace_bytes)");
  MMGL_REQUIRE(aligned16(d_o) && aligned16(a->q) && aligned16(a->k) && aligned16(a->v) && aligned16(a->o) && aligned16(dq) &&
               aligned16(dk) && aligned16(dv) && aligned16(workspace), "mmgl_attn_bwd: pointers must be 16B aligned");
  MMGL_REQUIRE(lddo % 8 == 0 && a->ldq % 8 == 0 && a->ldk % 8 == 0 && a->ldv % 8 == 0 && a->ldo % 8 == 0 && lddq % 8 == 0 &&
               lddk % 8 == 0 && lddv % 8 == 0, "mmgl_attn_bwd: leading dims must be multiples of 8");
  Maps mp;
  if (int rc = build_maps(mp, a->q, a->ldq, a->k, a->ldk, a->v, a->ldv, d_o, lddo, a->batch, a->seq_q, a->seq_k, a->heads, (int)a->head_dim)) return rc;
  float* ws = reinterpret_cast<float*>(workspace);
  if (a->head_dim == 64)
    return launch_bwd<64>(mp, p, a->o, a->ldo, d_o, lddo, a->stats, dq, lddq, dk, lddk, dv, lddv, ws, workspace_bytes, s);
  return launch_bwd<128>(mp, p, a->o, a->ldo, d_o, lddo, a->stats, dq, lddq, dk, lddk, dv, lddv, ws, workspace_bytes, s);
}
