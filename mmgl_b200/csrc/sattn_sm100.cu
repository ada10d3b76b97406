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
//             score row, each thread = one TMEM lane and half of the columns), two MMA-issuing warps (one per tile; the whole
//             warp walks the schedule, an elected lane issues from uniform registers) and one TMA warp.  The K / V blocks
//             stream once through a 4-stage ring (2 at head_dim 128) for both tiles.  Pass 1 double-buffers S in TMEM; pass 2
//             works on 64-key half blocks with three S buffers per tile (two at head_dim 128); P is written back (bf16 pairs,
//             tcgen05.st) over the head of the score columns it came from and P V reads it from tensor memory (TS form), so
//             the tensor core runs ahead of the exponentials and P never touches shared memory.  Only blocks at or below the
//             diagonal are visited when causal (tiles are paired heavy-with-next-heavy, heavy pairs first).  A packed
//             variable-length batch (cu_seqlens) is supported.
//   backward: csrc/sattn_bwd_sm100.cu (single pass).
// Masks are bit masks: one 32-bit word per 32 keys (attend = exists AND not padding), built once per CTA; the causal
// limit of the diagonal block is a per-thread shift.  Interior blocks (all 128 keys attended, not diagonal, no bias)
// take a 3-instruction-per-score path.
#include "sattn_common.cuh"

namespace mmgl {
namespace {

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
  static constexpr int pfull = 10 + 4 * NS;  // [2 parities][2 tiles]  pass 2: tile t's P half-block is in tensor memory (256 arrivals)
  static constexpr int pfree = 14 + 4 * NS;  // [2 parities][2 tiles]  P V of that half-block complete: O_t updated
  static constexpr int s2full = 18 + 4 * NS; // [3 buffers][2 tiles]  pass 2: S of a half block complete in TMEM buffer (index 2 * buf + t)
  static constexpr int count = 24 + 4 * NS;
};

template <int D, bool kBias, bool kDrop>
__global__ void __launch_bounds__(608, 1)
sattn_fwd_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                 const __grid_constant__ CUtensorMap map_v, const __grid_constant__ AttnParams p,
                 __nv_bfloat16* __restrict__ o, int64_t ldo, float* __restrict__ stats) {
  constexpr int TB = (D / 64) * 16384;   // bytes of one [128][D] tile
  constexpr int NS = (D == 64) ? 4 : 2;  // K / V ring stages
  using B = FwdBars<NS>;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;                    // 2 tiles
  uint8_t* sK = sQ + 2 * TB;             // NS stages
  uint8_t* sV = sK + NS * TB;            // NS stages
  const int nbk_max = (p.seq_k + 127) / 128;                       // layout: sized for the longest sample
  uint32_t* kbits = reinterpret_cast<uint32_t*>(sV + NS * TB);     // 4 * nbk words (P lives in tensor memory: no P tile here)
  uint64_t* bars = reinterpret_cast<uint64_t*>(kbits + 4 * nbk_max + ((4 * nbk_max) & 1));
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + B::count);
  float* xch = reinterpret_cast<float*>(tmem_ptr + 4);   // [2 tiles][2 column halves][128 rows]: row max, then row sum

  const int h = blockIdx.y, b = blockIdx.z;
  // this sample's lengths and first rows: uniform batch, or a packed variable-length batch (cu_seqlens)
  int sq = p.seq_q, sk = p.seq_k, rowq = b * p.seq_q, rowk = b * p.seq_k;
  if (p.cu_seqlens != nullptr) {
    pdl_wait();   // a global read ahead of the common wait point below
    const int c0 = p.cu_seqlens[b], c1 = p.cu_seqlens[b + 1];
    sq = sk = c1 - c0;
    rowq = rowk = c0;
  }
  const int nbk = (sk + 127) / 128;
  const int ntq = (sq + 127) / 128;
  if (ntq - 1 - 2 * (int)blockIdx.x < 0) return;   // a shorter sample has no such tile pair (whole CTA, before any barrier)
  // warp index through a lane-0 broadcast: ptxas then knows it is warp-uniform and keeps everything derived from it (the
  // MMA issuers' descriptors) on the uniform datapath
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
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

  // The producer lane initialises the barriers and starts the first loads (both Q tiles, the first NS K blocks) BEFORE the
  // block-wide set-up (TMEM allocation, key-bit words): the ~3000-cycle first-touch latency of the CTA's operands overlaps it.
  constexpr int kTmaThread = 18 * 32;
  if (threadIdx.x == kTmaThread) {
    tma_prefetch_desc(&map_q); tma_prefetch_desc(&map_k); tma_prefetch_desc(&map_v);
    for (int i = 0; i < B::count; ++i) {
      const bool wide = (i >= B::sfree && i < B::sfree + 4) || (i >= B::pfull && i < B::pfull + 4);
      const bool two = (i >= B::kfree && i < B::kfree + NS) || (i >= B::vfree && i < B::vfree + NS);
      mbar_init(&bars[i], wide ? 256 : (two ? 2 : 1));
    }
    fence_barrier_init();
    pdl_launch();
    pdl_wait();                                          // (every other thread waits below, before its first global access)
    for (int t = 0; t < 2; ++t)
      if (nblk[t] > 0) {
        mbar_arrive_expect_tx(&bars[B::qfull + t], TB);
        tma_tile<D>(sQ + t * TB, &map_q, &bars[B::qfull + t], colq, rowq + qts[t] * 128);
      }
    for (int n = 0; n < NS && n < nbmax; ++n) {           // pass-1 blocks 0 .. NS-1 (a ring stage each, nothing to wait for)
      mbar_arrive_expect_tx(&bars[B::kfull + n], TB);
      tma_tile<D>(sK + n * TB, &map_k, &bars[B::kfull + n], colq, rowk + n * 128);
    }
  }
  if (warp == 16) tmem_alloc<512>(tmem_ptr);
  if (threadIdx.x != kTmaThread) { pdl_launch(); pdl_wait(); }
  build_key_bits(kbits, p.key_mask, b, sk, 4 * nbk);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);
  if (threadIdx.x == 0) TR(0, tri, 2);

  if (warp == 18) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      for (int j = 0; j < NS && j < nbmax; ++j) {     // the first V blocks have a ring stage each: fetch them during pass 1
        mbar_arrive_expect_tx(&bars[B::vfull + j], TB);
        tma_tile<D>(sV + j * TB, &map_v, &bars[B::vfull + j], colq, rowk + j * 128);
      }
      for (int n = 0; n < 2 * nbmax; ++n) {           // K block stream: pass 1 then pass 2
        const int j = n < nbmax ? n : n - nbmax, s = n % NS;
        if (n < NS && n < nbmax) continue;              // issued ahead of the block-wide set-up, above
        if (n >= NS) mbar_wait(&bars[B::kfree + s], ((n / NS) - 1) & 1);
        mbar_arrive_expect_tx(&bars[B::kfull + s], TB);
        tma_tile<D>(sK + s * TB, &map_k, &bars[B::kfull + s], colq, rowk + j * 128);
        TR(3, tri, 100 + n);
        if (n >= nbmax && j >= NS) {                    // pass 2: V block j rides along
          const int sv = j % NS;
          mbar_wait(&bars[B::vfree + sv], ((j / NS) - 1) & 1);
          mbar_arrive_expect_tx(&bars[B::vfull + sv], TB);
          tma_tile<D>(sV + sv * TB, &map_v, &bars[B::vfull + sv], colq, rowk + j * 128);
        }
      }
    }
    __syncwarp();
  } else if (warp >= 16) {
    // ------------------------------------------------------------------ MMA issuers: warp 16 -> tile 0, warp 17 -> tile 1
    // The WHOLE warp walks the loops (barrier waits included) and one elected lane issues: with every operand derived
    // from warp-uniform values the descriptors live in uniform registers and the UTCHMMAs of a block go out back to
    // back.  (Under `if (lane == 0)` ptxas wrapped every MMA in an ELECT / 5 x R2UR waterfall loop, ~100 cycles of
    // dependent issue latency per 48-cycle MMA: the issuer thread, not the tensor pipe, set the pace of pass 2.)
    const int t = warp - 16;
    const int nb = t ? nblk[1] : nblk[0];
    const int nb1 = nblk[1];
    if (nb > 0) {
      const uint64_t dq = make_smem_desc(smem_u32(sQ + t * TB), 16, 1024);
      const uint64_t dk0 = make_smem_desc(smem_u32(sK), 16, 1024);
      const uint64_t dv0 = make_smem_desc(smem_u32(sV), 16384, 1024);
      const uint32_t cS = tmem_base + t * 128;
      // a K / V stage is released by two arrivals, one per tile; the only tile using a block arrives twice
      auto release = [&](uint64_t* bar, int j) {
        umma_commit(bar);
        if (j >= nb1) umma_commit(bar);
      };
      mbar_wait(&bars[B::qfull + t], 0);
      if (t == 0 && lane == 0) TR(2, tri, 10);
      // pass 1: S_t(j) for the row maxima, double-buffered (block j at column offset (j & 1) * 256; the O columns are
      // not in use yet), so the tensor core runs one block ahead of the max reduction
      for (int j = 0; j < nb; ++j) {
        const int s = j % NS;
        mbar_wait(&bars[B::kfull + s], (j / NS) & 1);
        if (j >= 2) mbar_wait(&bars[B::sfree + 2 * (j & 1) + t], ((j >> 1) - 1) & 1);
        tc_fence_after();
        if (elect_one()) {
          if (t == 0) TR(2, tri, 100 + j);
          mma_qk_desc<D>(cS + (j & 1) * 256, dq, dk0 + (uint64_t)((s * TB) >> 4));
          umma_commit(&bars[B::sfull + 2 * (j & 1) + t]);
          release(&bars[B::kfree + s], j);
          if (t == 0) TR(2, tri, 150 + j);
        }
        __syncwarp();
      }
      // pass 2 prologue: S_t(0) once BOTH tiles' pass-1 reads are done (tile 0's O columns alias tile 1's second S buffer
      // and vice versa); barrier phases are consumed in order
      for (int tt = 0; tt < 2; ++tt)
        for (int jj = max(nblk[tt] - 2, 0); jj < nblk[tt]; ++jj)
          mbar_wait(&bars[B::sfree + 2 * (jj & 1) + tt], (jj >> 1) & 1);
      // Pass 2 works on HALF blocks (64 keys) with NSB S buffers per tile in TMEM: half block x = half (x & 1) of key block
      // x >> 1 lives in S buffer x % NSB, and so does its P once the softmax warps have written it back.  S(jj + NSB) is issued right after
      // P V(jj), so the softmax warps always find the next scores ready and the hand-off latency is off the critical path.
      constexpr int NSB = (D == 64) ? 3 : 2;                 // S buffers per tile in pass 2 (TMEM: 2 * NSB * 64 + 2 * D <= 512)
      const uint32_t cS2 = tmem_base + t * (NSB * 64);       // + v * 64
      const uint32_t cO2 = tmem_base + 2 * NSB * 64 + t * D;
      const int nsub = 2 * nb;
      // S of half block x (half x & 1 of key block x >> 1) into S buffer x % NSB; the first half waits for its K stage, the
      // second releases it
      auto issue_s = [&](int x) {
        const int jn = x >> 1, hk = x & 1, n = nbmax + jn, st = n % NS;
        if (hk == 0) mbar_wait(&bars[B::kfull + st], (n / NS) & 1);
        tc_fence_after();
        if (elect_one()) {
          mma_qk_half<D>(cS2 + (x % NSB) * 64, dq, dk0 + (uint64_t)((st * TB + hk * 8192) >> 4));
          umma_commit(&bars[B::s2full + 2 * (x % NSB) + t]);
          if (hk == 1) release(&bars[B::kfree + st], jn);
        }
        __syncwarp();
      };
      if (t == 0 && lane == 0) TR(2, tri, 198);
      for (int x = 0; x < NSB && x < nsub; ++x) issue_s(x);   // the tensor core starts NSB half blocks ahead of the softmax
      if (t == 0 && lane == 0) TR(2, tri, 199);
      for (int jj = 0; jj < nsub; ++jj) {
        const int pp = jj & 1, j = jj >> 1, sv = j % NS;
        mbar_wait(&bars[B::pfull + 2 * pp + t], (jj >> 1) & 1);
        if (t == 0 && lane == 0) TR(2, tri, 200 + jj);
        if (pp == 0) mbar_wait(&bars[B::vfull + sv], (j / NS) & 1);
        tc_fence_after();
        if (elect_one()) {
          mma_pv_half_ts<D>(cO2, cS2 + (jj % NSB) * 64, dv0 + (uint64_t)((sv * TB + pp * 8192) >> 4), jj != 0);
          umma_commit(&bars[B::pfree + 2 * pp + t]);
          if (pp == 1) release(&bars[B::vfree + sv], j);
        }
        __syncwarp();
        if (jj + NSB < nsub) issue_s(jj + NSB);               // into the S buffer the softmax warps have just finished reading
        if (t == 0 && lane == 0) TR(2, tri, 300 + jj);
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
      // NSB S buffers in TMEM (3 at head_dim 64), P written back into the buffer it came from: while this tile's warps exponentiate half
      // block jj, the tensor core already holds S(jj + 1) (and S(jj + 2)) and is free to run P V(jj - 1) and S(jj + NSB), so the
      // softmax -> MMA -> softmax hand-off latency stays off the critical path.  This thread owns chunk 2u + hf.
      float sum = 0.f;
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
        if (jj >= 2) mbar_wait(&bars[B::pfree + 2 * u + t], ((jj >> 1) - 1) & 1);   // consume every phase: the final wait is by parity
        // P (bf16 pairs) overwrites the head of the score columns this thread has just read: the P V MMA takes it from
        // tensor memory, and S(jj + NSB) is issued behind that MMA, so the buffer is not rewritten before it is consumed
        uint32_t pk[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) pk[e] = pack_bf16(__uint_as_float(rc[2 * e]), __uint_as_float(rc[2 * e + 1]));
        tmem_st_32x16(cS2 + v * 64 + hf * 32, pk);
        tmem_st_wait();
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

template <int D>
int launch_fwd(const Maps& mp, const AttnParams& p, void* o, int64_t ldo, float* stats, int64_t batch, cudaStream_t stream) {
  constexpr int TB = (D / 64) * 16384;
  constexpr int NS = (D == 64) ? 4 : 2;
  const size_t smem = (2 + 2 * NS) * (size_t)TB + kbits_bytes(p.seq_k) + FwdBars<NS>::count * 8 + 16 + 2 * 2 * 128 * 4;
  const bool bias = p.rel_bias != nullptr, drop = p.drop_thresh != 0;
  auto kern = bias ? (drop ? sattn_fwd_kernel<D, true, true> : sattn_fwd_kernel<D, true, false>)
                   : (drop ? sattn_fwd_kernel<D, false, true> : sattn_fwd_kernel<D, false, false>);
  MMGL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int ntq = (p.seq_q + 127) / 128;
  dim3 grid((unsigned)((ntq + 1) / 2), (unsigned)p.heads, (unsigned)batch);
  MMGL_CUDA(launch_pdl(kern, grid, dim3(608), smem, stream, mp.q, mp.k, mp.v, p, (__nv_bfloat16*)o, ldo, stats));
  return check_launch("mmgl_attn_fwd");
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

