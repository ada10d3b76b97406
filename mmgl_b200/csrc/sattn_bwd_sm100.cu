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
//     per-CTA fp32 scratch (L2-resident: grid x seq_q x head_dim floats) with st / red.global.add.v4.f32 / ld in a fixed
//     order: deterministic, no zero-fill pass, no fp32 -> bf16 conversion kernel; the last key block of a tile adds its
//     part, scales and stores bf16 dQ directly;
//   * delta = rowsum(dO . O) is formed in the j = 0 pass (every query tile sees key block 0) and kept in shared memory.
// Roles (576 threads): 16 softmax-gradient warps (thread = one score row x 16 of the 64 keys of a half block), one
// MMA-issuing warp, one TMA warp that streams K_j / V_j and Q_i / dO_i tiles through double-buffered rings (head_dim 64)
// across step and item boundaries.  S / dP live in TMEM as two 64-key halves, so the tensor core refills a half as soon as
// the threads have pulled it into registers and runs S / dP of the NEXT step while this step's exponentials execute.
// TMEM columns: head_dim 64: 2 x (S 64 + dP 64) + dK 64 + dV 64 + dQ 64 = 448; head_dim 128: (64 + 64) + 128 + 128 + 128 = 512.
#include "sattn_common.cuh"

namespace mmgl {
namespace {

template <int D>
struct BwdCfg {
  static constexpr int TB = (D / 64) * 16384;        // bytes of one [128][D] bf16 tile
  static constexpr int NKV = (D == 64) ? 2 : 1;      // K / V buffers
  static constexpr int NQ = (D == 64) ? 2 : 1;       // Q / dO buffers
  static constexpr int NSB = (D == 64) ? 2 : 1;      // S / dP half buffers in TMEM
  static constexpr uint32_t cS = 0;                  // + buf * 128: S half [128 x 64], then dP half at + 64
  static constexpr uint32_t cdK = NSB * 128;
  static constexpr uint32_t cdV = cdK + D;
  static constexpr uint32_t cdQ = cdV + D;
  // barriers
  static constexpr int kvfull = 0, kvfree = NKV, qfull = 2 * NKV, qfree = 2 * NKV + NQ, sfull = 2 * NKV + 2 * NQ,
                       sfree = sfull + NSB, pfull = sfree + NSB, mma2done = pfull + 1, nbars = mma2done + 1;
};

// (item, key block, query tile) enumeration shared by the three roles: every role walks the same sequence of steps
struct StepIter {
  int item, n_items, stride, heads, nbk, ntq, causal, coff;
  int b, h, j, i;
  int sc, jc;          // running step / key-block counters of this CTA (buffer indices and barrier phases derive from them)
  bool done;
  __device__ __forceinline__ int first_tile(int jj) const { return causal ? max(0, jj * 128 - coff) / 128 : 0; }
  __device__ __forceinline__ int blocks_of(int ii) const { return causal ? min(nbk, (ii * 128 + 127 + coff) / 128 + 1) : nbk; }
  __device__ __forceinline__ void init(const AttnParams& p, int first_item, int stride_) {
    n_items = p.batch * p.heads; stride = stride_; heads = p.heads; causal = p.causal; coff = p.coff;
    nbk = (p.seq_k + 127) / 128; ntq = (p.seq_q + 127) / 128;
    item = first_item; sc = 0; jc = 0; done = item >= n_items;
    b = item / heads; h = item % heads; j = 0; i = 0;     // first_tile(0) == 0
  }
  __device__ __forceinline__ bool first_of_j() const { return i == first_tile(j); }
  __device__ __forceinline__ bool last_of_j() const { return i == ntq - 1; }
  __device__ __forceinline__ bool first_of_item() const { return j == 0 && i == 0; }
  __device__ __forceinline__ void next() {
    ++sc;
    if (++i < ntq) return;
    ++jc;
    if (++j < nbk) { i = first_tile(j); return; }
    item += stride;
    if (item >= n_items) { done = true; return; }
    b = item / heads; h = item % heads; j = 0; i = 0;
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

__device__ __forceinline__ float delta_partial_row(const __nv_bfloat16* po_, const __nv_bfloat16* pd_, int n16) {
  const uint4* po = reinterpret_cast<const uint4*>(po_);
  const uint4* pd = reinterpret_cast<const uint4*>(pd_);
  float acc = 0.f;
  for (int i = 0; i < n16; ++i) {
    const uint4 vo = __ldg(po + i), vd = __ldg(pd + i);
    const uint32_t wo[4] = {vo.x, vo.y, vo.z, vo.w}, wd[4] = {vd.x, vd.y, vd.z, vd.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) acc += bf16lo(wo[e]) * bf16lo(wd[e]) + bf16hi(wo[e]) * bf16hi(wd[e]);
  }
  return acc;
}

// what the softmax threads need to finish a step after its second group of MMAs has completed
struct PrevStep {
  int valid, row, row_ok, first, last, last_of_j, j, colq, rowq, rowk;
};

template <int D, bool kBias, bool kDrop>
__global__ void __launch_bounds__(576, 1)
sattn_bwd_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_do,
                 const __grid_constant__ CUtensorMap map_k, const __grid_constant__ CUtensorMap map_v,
                 const __grid_constant__ AttnParams p, const __nv_bfloat16* __restrict__ o, int64_t ldo,
                 const __nv_bfloat16* __restrict__ d_o, int64_t lddo, const float* __restrict__ stats,
                 __nv_bfloat16* __restrict__ dq, int64_t lddq, __nv_bfloat16* __restrict__ dk, int64_t lddk,
                 __nv_bfloat16* __restrict__ dv, int64_t lddv, float* __restrict__ dq_ws) {
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
  float* sDelta = reinterpret_cast<float*>(sdS + 32768);     // [ntq * 128] rowsum(dO . O) of the current item
  float* sPart = sDelta + ntq * 128;                          // [2][4][128] partial row sums (exchange, double-buffered)
  float* sBins = sPart + 1024;                                // [256] diagonal sums of dS (gradient of the relative-position bias)
  uint32_t* kbits = reinterpret_cast<uint32_t*>(sBins + 256); // [2][4 * nbk] attend bits of the item's sample (double-buffered)
  uint64_t* bars = reinterpret_cast<uint64_t*>(kbits + 2 * (4 * nbk + ((4 * nbk) & 1)));
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + C::nbars);
  const int kb_stride = 4 * nbk + ((4 * nbk) & 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
  if (tid == 0) {
    tma_prefetch_desc(&map_q); tma_prefetch_desc(&map_do); tma_prefetch_desc(&map_k); tma_prefetch_desc(&map_v);
    for (int i = 0; i < C::nbars; ++i) {
      const bool wide = (i >= C::sfree && i < C::sfree + C::NSB) || i == C::pfull;
      mbar_init(&bars[i], wide ? 512 : 1);
    }
    fence_barrier_init();
  }
  if (warp == 16) tmem_alloc<512>(tmem_ptr);
  if (tid < 256) sBins[tid] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 17) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      StepIter it;
      it.init(p, blockIdx.x, gridDim.x);
      while (!it.done) {
        const int colq = it.h * D, rowq = it.b * p.seq_q, rowk = it.b * p.seq_k;
        if (it.first_of_j()) {
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
        it.next();
      }
    }
    __syncwarp();
  } else if (warp == 16) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint64_t desc_q = make_smem_desc(smem_u32(sQ), 16, 1024), desc_do = make_smem_desc(smem_u32(sdO), 16, 1024);
      const uint64_t desc_k = make_smem_desc(smem_u32(sK), 16, 1024), desc_v = make_smem_desc(smem_u32(sV), 16, 1024);
      const uint64_t desc_ds = make_smem_desc(smem_u32(sdS), 16, 1024);
      const uint64_t desc_k_mn = make_smem_desc(smem_u32(sK), 16384, 1024);
      const uint64_t desc_p_mn = make_smem_desc(smem_u32(sP), 16384, 1024), desc_ds_mn = make_smem_desc(smem_u32(sdS), 16384, 1024);
      const uint64_t desc_q_mn = make_smem_desc(smem_u32(sQ), 16384, 1024), desc_do_mn = make_smem_desc(smem_u32(sdO), 16384, 1024);
      StepIter cur, la;                 // cur: the step whose dV / dK / dQ MMAs are next; la: the step whose S / dP halves are next
      cur.init(p, blockIdx.x, gridDim.x);
      la.init(p, blockIdx.x, gridDim.x);
      int la_half = 0;                  // next half (0 / 1) of step `la` to issue
      // S / dP of half x = 2 * la.sc + la_half into TMEM buffer x % NSB
      auto issue_s = [&]() {
        const int x = 2 * la.sc + la_half, sb = x % C::NSB;
        const int qb = la.sc % C::NQ, kb = la.jc % C::NKV;
        if (la_half == 0) {
          if (la.first_of_j()) mbar_wait(&bars[C::kvfull + kb], (la.jc / C::NKV) & 1);
          mbar_wait(&bars[C::qfull + qb], (la.sc / C::NQ) & 1);
        }
        if (x >= C::NSB) mbar_wait(&bars[C::sfree + sb], ((x / C::NSB) - 1) & 1);
        tc_fence_after();
        const uint64_t qoff = (uint64_t)((qb * TB) >> 4), koff = (uint64_t)((kb * TB + la_half * 8192) >> 4);
        mma_qk_half<D>(tmem_base + C::cS + sb * 128, desc_q + qoff, desc_k + koff);          // S  = Q_i  K_j[half]^T
        mma_qk_half<D>(tmem_base + C::cS + sb * 128 + 64, desc_do + qoff, desc_v + koff);    // dP = dO_i V_j[half]^T
        umma_commit(&bars[C::sfull + sb]);
        if (la_half == 1) la.next();
        la_half ^= 1;
      };
      while (!cur.done) {
        // run S / dP ahead: up to the halves of the NEXT step when its Q / dO tiles have their own buffer (NQ == 2),
        // otherwise to the end of this step (the single Q / dO buffer is released by this step's second MMA group)
        const int x_max = (C::NQ >= 2) ? 2 * cur.sc + 3 : 2 * cur.sc + 1;
        while (!la.done && 2 * la.sc + la_half <= x_max) issue_s();
        const int qb = cur.sc % C::NQ, kb = cur.jc % C::NKV;
        mbar_wait(&bars[C::pfull], cur.sc & 1);
        tc_fence_after();
        const uint64_t qoff = (uint64_t)((qb * TB) >> 4), koff = (uint64_t)((kb * TB) >> 4);
        const bool acc = !cur.first_of_j();
        mma_tn_desc<D>(tmem_base + C::cdV, desc_p_mn, desc_do_mn + qoff, acc);     // dV_j += P~^T dO_i
        mma_tn_desc<D>(tmem_base + C::cdK, desc_ds_mn, desc_q_mn + qoff, acc);     // dK_j += dS^T Q_i
        mma_pv_desc<D>(tmem_base + C::cdQ, desc_ds, desc_k_mn + koff, false);      // dQ part = dS K_j
        umma_commit(&bars[C::mma2done]);
        umma_commit(&bars[C::qfree + qb]);
        if (cur.last_of_j()) umma_commit(&bars[C::kvfree + kb]);
        cur.next();
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ softmax-gradient warps (512 threads)
    const int rit = tid & 127, quarter = tid >> 7;           // row in tile; 16-key column group of every 64-key half
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    float* ws = dq_ws + (size_t)blockIdx.x * ((size_t)ntq * 128 * D);
    const bool want_dbias = kBias && p.d_rel_bias != nullptr;
    const int64_t dgroups = (p.seq_k + 7) >> 3;
    PrevStep prev;
    prev.valid = 0;

    // finish a step once its dV / dK / dQ MMAs are complete: fold the dQ part into the scratch (or emit dQ), and after
    // the last query tile of a key block write dK_j / dV_j
    auto finish_prev = [&]() {
      {
        constexpr int NC = D / 4;                                    // this thread's columns of the dQ row
        float* acc = ws + ((size_t)prev.row * D + quarter * NC);
        __nv_bfloat16* dst = dq + ((int64_t)prev.rowq + prev.row) * lddq + prev.colq + quarter * NC;
#pragma unroll
        for (int c0 = 0; c0 < NC; c0 += 16) {
          uint32_t r[16];
          tmem_ld_32x16(lane_addr + C::cdQ + quarter * NC + c0, r);   // warp-collective: outside the row predicate
          tmem_ld_wait();
          if (!prev.row_ok) continue;
          float f[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) f[e] = __uint_as_float(r[e]);
          if (prev.last) {
            if (!prev.first) {
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                float4 a;
                asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w)
                             : "l"(acc + c0 + 4 * g) : "memory");
                f[4 * g] += a.x; f[4 * g + 1] += a.y; f[4 * g + 2] += a.z; f[4 * g + 3] += a.w;
              }
            }
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              uint4 v;
              v.x = pack_bf16(f[8 * g] * p.scale, f[8 * g + 1] * p.scale);
              v.y = pack_bf16(f[8 * g + 2] * p.scale, f[8 * g + 3] * p.scale);
              v.z = pack_bf16(f[8 * g + 4] * p.scale, f[8 * g + 5] * p.scale);
              v.w = pack_bf16(f[8 * g + 6] * p.scale, f[8 * g + 7] * p.scale);
              *reinterpret_cast<uint4*>(dst + c0 + 8 * g) = v;
            }
          } else if (prev.first) {
#pragma unroll
            for (int g = 0; g < 4; ++g)
              asm volatile("st.global.cg.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(acc + c0 + 4 * g), "f"(f[4 * g]), "f"(f[4 * g + 1]),
                           "f"(f[4 * g + 2]), "f"(f[4 * g + 3]) : "memory");
          } else {
#pragma unroll
            for (int g = 0; g < 4; ++g)
              asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(acc + c0 + 4 * g), "f"(f[4 * g]), "f"(f[4 * g + 1]),
                           "f"(f[4 * g + 2]), "f"(f[4 * g + 3]) : "memory");
          }
        }
      }
      if (prev.last_of_j) {
        // quarters 0,1 store dV, quarters 2,3 store dK; each stores half of the D columns of its key row
        const int key = prev.j * 128 + rit;
        const bool is_k = quarter >= 2;
        const int half = quarter & 1;
        const uint32_t src = lane_addr + (is_k ? C::cdK : C::cdV) + half * (D / 2);
        __nv_bfloat16* dst = (is_k ? dk + ((int64_t)prev.rowk + key) * lddk : dv + ((int64_t)prev.rowk + key) * lddv) + prev.colq + half * (D / 2);
        const float mul = is_k ? p.scale : 1.f;
#pragma unroll
        for (int c = 0; c < D / 64; ++c) {
          uint32_t r[32];
          tmem_ld_32x32(src + c * 32, r);
          tmem_ld_wait();
          if (key < p.seq_k) store_row_bf16(dst + c * 32, r, mul);
        }
      }
      tc_fence_before();
    };

    StepIter it;
    it.init(p, blockIdx.x, gridDim.x);
    int item_count = 0;
    const uint32_t* kb_cur = kbits;
    while (!it.done) {
      const int colq = it.h * D, rowq = it.b * p.seq_q, rowk = it.b * p.seq_k;
      if (it.first_of_item()) {
        // attend bits of this item's sample (other items' readers may still use the other buffer)
        uint32_t* kb_w = kbits + (item_count & 1) * kb_stride;
        for (int w = warp; w < 4 * nbk; w += 16) {
          const int key = w * 32 + lane;
          const bool a = key < p.seq_k && (p.key_mask == nullptr || p.key_mask[(int64_t)it.b * p.seq_k + key] != 0);
          const uint32_t bits = __ballot_sync(0xffffffffu, a);
          if (lane == 0) kb_w[w] = bits;
        }
        kb_cur = kb_w;
        ++item_count;
        asm volatile("bar.sync 1, 512;" ::: "memory");
      }
      const int i = it.i, j = it.j;
      const int row = i * 128 + rit;
      const bool row_ok = row < p.seq_q;
      float m = 0.f, inv = 0.f;
      if (row_ok) {
        const float2 st = __ldg(reinterpret_cast<const float2*>(stats + (((int64_t)it.b * p.heads + it.h) * p.seq_q + row) * 2));
        m = st.x; inv = st.y;
      }
      float delta;
      if (j == 0) {
        // delta = rowsum(dO . O): this thread's quarter of the row, the four quarters meet in shared memory
        float* part = sPart + (i & 1) * 512;
        const int64_t grow = (int64_t)rowq + row;
        part[quarter * 128 + rit] = row_ok ? delta_partial_row(o + grow * ldo + colq + quarter * (D / 4),
                                                               d_o + grow * lddo + colq + quarter * (D / 4), D / 32) : 0.f;
        asm volatile("bar.sync 1, 512;" ::: "memory");
        delta = (part[rit] + part[128 + rit]) + (part[256 + rit] + part[384 + rit]);
        sDelta[row] = delta;      // all four threads of the row write the same value; each later reads what it wrote itself
      } else {
        delta = sDelta[row];
      }
      const bool none = !(m > -FLT_MAX);
      const bool flat = none || !row_ok;       // no attended key (uniform row) or a row beyond the sequence (p = 0)
      const float c1 = flat ? 0.f : p.scale * kL2E, mc = flat ? 0.f : m * kL2E, bsc = flat ? 0.f : kL2E;
      const float inv_ok = row_ok ? inv : 0.f;
      const float* bias_row = kBias ? p.rel_bias + (int64_t)it.h * (p.seq_q + p.seq_k - 1) + (p.seq_q - 1 - min(row, p.seq_q - 1)) : nullptr;
      const int64_t drow = ((int64_t)it.b * p.heads + it.h) * p.seq_q + row;
      const uint32_t p_base = smem_u32(sP), ds_base = smem_u32(sdS);

#pragma unroll 1
      for (int hk = 0; hk < 2; ++hk) {
        const int x = 2 * it.sc + hk, sb = x % C::NSB;
        const int koff = hk * 64 + quarter * 16;           // first key of this thread's group inside the block
        const int key0 = j * 128 + koff;
        // attend bits of the 16 keys for this row
        uint32_t mw = (kb_cur[4 * j + (koff >> 5)] >> (koff & 16)) & 0xffffu;
        if (p.causal) mw &= low_bits(row + p.coff - key0 + 1);
        if (none) mw = low_bits(p.seq_k - key0) & 0xffffu;   // uniform over the existing keys of the visited blocks
        uint32_t keep = 0xffffu;
        if (kDrop) keep = keep_half_word(p, drow, key0, dgroups);
        const float ks = kDrop ? p.drop_scale : 1.f;
        const int lim = max(p.seq_k - 1 - key0, 0);
        const float* bk = kBias ? bias_row + min(key0, p.seq_k - 1) : nullptr;

        mbar_wait(&bars[C::sfull + sb], (x / C::NSB) & 1);
        tc_fence_after();
        uint32_t rs[16], rp[16];
        tmem_ld_32x16(lane_addr + C::cS + sb * 128 + quarter * 16, rs);
        tmem_ld_32x16(lane_addr + C::cS + sb * 128 + 64 + quarter * 16, rp);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&bars[C::sfree + sb]);

        uint32_t pk[8], dk_[8];
#pragma unroll
        for (int e = 0; e < 16; e += 2) {
          float pv[2], dv_[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            float off = -mc;
            if (kBias) off = fmaf(__ldg(bk + min(e + u, lim)), bsc, -mc);
            float pr = ex2(fmaf(__uint_as_float(rs[e + u]), c1, off)) * inv_ok;
            pr = (mw >> (e + u)) & 1u ? pr : 0.f;
            if (kDrop) {
              const float kmul = (keep >> (e + u)) & 1u ? ks : 0.f;
              pv[u] = pr * kmul;
              dv_[u] = pr * fmaf(__uint_as_float(rp[e + u]), kmul, -delta);
            } else {
              pv[u] = pr;
              dv_[u] = pr * (__uint_as_float(rp[e + u]) - delta);
            }
          }
          pk[e >> 1] = pack_bf16(pv[0], pv[1]);
          dk_[e >> 1] = pack_bf16(dv_[0], dv_[1]);
          if (kBias && want_dbias) {   // d bias(key - row) += dS: one bin per diagonal of the 128 x 128 block
            atomicAdd(sBins + (koff + e - rit + 127), dv_[0]);
            atomicAdd(sBins + (koff + e + 1 - rit + 127), dv_[1]);
          }
        }
        if (hk == 0 && prev.valid) {
          // the previous step's dV / dK / dQ MMAs have read P / dS (and written the dQ part): finish it, then reuse the tiles
          mbar_wait(&bars[C::mma2done], (it.sc - 1) & 1);
          tc_fence_after();
          finish_prev();
        }
        // 16 keys = two 16-byte chunks (2 * quarter, 2 * quarter + 1) of row `rit` in the half's 64-key slab
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          sts128(swz(p_base + hk * 16384, rit, 2 * quarter + g), pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]);
          sts128(swz(ds_base + hk * 16384, rit, 2 * quarter + g), dk_[4 * g], dk_[4 * g + 1], dk_[4 * g + 2], dk_[4 * g + 3]);
        }
      }
      if (want_dbias) {   // flush this block's 255 diagonals: bin t holds key - row = j * 128 - i * 128 + t - 127
        asm volatile("bar.sync 1, 512;" ::: "memory");
        if (tid < 255) {
          const float v = sBins[tid];
          const int idx = j * 128 - i * 128 + tid - 127 + p.seq_q - 1;
          if (v != 0.f && idx >= 0 && idx < p.seq_q + p.seq_k - 1)
            atomicAdd(p.d_rel_bias + (int64_t)it.h * (p.seq_q + p.seq_k - 1) + idx, v);
          sBins[tid] = 0.f;
        }
        asm volatile("bar.sync 1, 512;" ::: "memory");
      }
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(&bars[C::pfull]);
      prev.valid = 1; prev.row = row; prev.row_ok = row_ok; prev.first = (j == 0); prev.last = (j == it.blocks_of(i) - 1);
      prev.last_of_j = it.last_of_j(); prev.j = j; prev.colq = colq; prev.rowq = rowq; prev.rowk = rowk;
      it.next();
    }
    if (prev.valid) {
      mbar_wait(&bars[C::mma2done], (it.sc - 1) & 1);
      tc_fence_after();
      finish_prev();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 16) tmem_dealloc<512>(tmem_base);
}

template <int D>
size_t bwd_smem_bytes(const AttnParams& p) {
  using C = BwdCfg<D>;
  const size_t nbk = (size_t)((p.seq_k + 127) / 128), ntq = (size_t)((p.seq_q + 127) / 128);
  const size_t kb_stride = 4 * nbk + ((4 * nbk) & 1);
  return (size_t)(2 * C::NKV + 2 * C::NQ) * C::TB + 65536 + ntq * 512 + 4096 + 1024 + 2 * kb_stride * 4 + C::nbars * 8 + 16;
}

template <int D, bool kBias, bool kDrop>
int launch_bwd_v(const Maps& mp, const AttnParams& p, const void* o, int64_t ldo, const void* d_o, int64_t lddo,
                 const float* stats, void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv, float* ws,
                 size_t ws_bytes, cudaStream_t stream) {
  const size_t smem = bwd_smem_bytes<D>(p);
  MMGL_REQUIRE(smem <= 227 * 1024, "mmgl_attn_bwd: seq_q = %d needs %zu bytes of shared memory (row statistics of a whole (sample, head) "
               "are kept on chip); supported up to about %d queries at head_dim %d", p.seq_q, smem, D == 64 ? 8192 : 8192, D);
  auto kern = sattn_bwd_kernel<D, kBias, kDrop>;
  MMGL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t items = (int64_t)p.heads * p.batch;
  const int64_t grid = items < sm_count() ? items : sm_count();
  const size_t need = (size_t)grid * (size_t)((p.seq_q + 127) / 128) * 128 * D * sizeof(float);
  MMGL_REQUIRE(ws_bytes >= need, "mmgl_attn_bwd: workspace too small (%zu < %zu bytes; use mmgl_attn_bwd_workspace_bytes)", ws_bytes, need);
  kern<<<dim3((unsigned)grid), 576, smem, stream>>>(mp.q, mp.d_o, mp.k, mp.v, p, (const __nv_bfloat16*)o, ldo,
                                                    (const __nv_bfloat16*)d_o, lddo, stats, (__nv_bfloat16*)dq, lddq,
                                                    (__nv_bfloat16*)dk, lddk, (__nv_bfloat16*)dv, lddv, ws);
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

// fp32 scratch for the per-CTA dQ accumulation: one [ceil(seq_q / 128) * 128, head_dim] tile set per resident CTA (an upper
// bound that does not depend on the caller knowing head_dim: 128 columns, one CTA per SM, at most batch * heads CTAs)
extern "C" size_t mmgl_attn_bwd_workspace_bytes(int64_t batch, int64_t seq_q, int64_t heads) {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t ctas = batch * heads < sms ? batch * heads : sms;
  return (size_t)ctas * (size_t)((seq_q + 127) / 128) * 128 * 128 * sizeof(float);
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
