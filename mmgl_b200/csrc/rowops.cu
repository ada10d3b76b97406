// HBM-bound row kernels of the neighbor-fusion step: LayerNorm fwd/bwd, bias / gate gradient reductions,
// ragged neighbor-bank packing (+ Laplacian-PE projection) and the GCN aggregate/concat helpers.
// All are streaming kernels: 16-byte vectorised, coalesced, fp32 math, no atomics (deterministic reductions).
#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <cuda_bf16.h>

#include "../../include/mmgl_b200.h"
#include "common.cuh"
#include "ptx.cuh"

namespace mmgl {

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  f[0] = bf16lo(u.x); f[1] = bf16hi(u.x); f[2] = bf16lo(u.y); f[3] = bf16hi(u.y);
  f[4] = bf16lo(u.z); f[5] = bf16hi(u.z); f[6] = bf16lo(u.w); f[7] = bf16hi(u.w);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 u;
  u.x = pack_bf16(f[0], f[1]); u.y = pack_bf16(f[2], f[3]); u.z = pack_bf16(f[4], f[5]); u.w = pack_bf16(f[6], f[7]);
  return u;
}

// ------------------------------------------------------------------------------------ LayerNorm forward
// one warp per row; the row stays in registers (MAXV 16-byte vectors per lane), two-pass mean / variance.
template <int MAXV>
__global__ void __launch_bounds__(256)
layernorm_fwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ gamma,
                     const float* __restrict__ beta, __nv_bfloat16* __restrict__ y, float* __restrict__ mean,
                     float* __restrict__ rstd, int64_t rows, int hidden, float eps) {
  pdl_launch();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int nvec = hidden >> 3;
  const uint4* xr = reinterpret_cast<const uint4*>(x + row * hidden);
  uint4 buf[MAXV];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int idx = lane + 32 * i;
    if (idx < nvec) {
      buf[i] = __ldg(xr + idx);
      float f[8]; unpack8(buf[i], f);
#pragma unroll
      for (int e = 0; e < 8; ++e) sum += f[e];
    }
  }
  const bool rms = mean == nullptr;   // T5-style RMSNorm: no mean subtraction, no shift
  const float mu = rms ? 0.f : warp_sum(sum) / hidden;
  float var = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int idx = lane + 32 * i;
    if (idx < nvec) {
      float f[8]; unpack8(buf[i], f);
#pragma unroll
      for (int e = 0; e < 8; ++e) { const float d = f[e] - mu; var += d * d; }
    }
  }
  const float rs = rsqrtf(warp_sum(var) / hidden + eps);
  uint4* yr = reinterpret_cast<uint4*>(y + row * hidden);
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int idx = lane + 32 * i;
    if (idx < nvec) {
      float f[8]; unpack8(buf[i], f);
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma) + idx * 2), g1 = __ldg(reinterpret_cast<const float4*>(gamma) + idx * 2 + 1);
      float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
      if (beta != nullptr) {
        b0 = __ldg(reinterpret_cast<const float4*>(beta) + idx * 2);
        b1 = __ldg(reinterpret_cast<const float4*>(beta) + idx * 2 + 1);
      }
      const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] = (f[e] - mu) * rs * gg[e] + bb[e];
      yr[idx] = pack8(f);
    }
  }
  if (lane == 0) { if (!rms) mean[row] = mu; rstd[row] = rs; }
}

// ------------------------------------------------------------------------------------ LayerNorm backward (dx)
template <int MAXV>
__global__ void __launch_bounds__(256)
layernorm_bwd_dx_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                        const float* __restrict__ gamma, const float* __restrict__ mean, const float* __restrict__ rstd,
                        const __nv_bfloat16* __restrict__ d_res, __nv_bfloat16* __restrict__ dx, int64_t rows,
                        int hidden) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int nvec = hidden >> 3;
  const uint4* dyr = reinterpret_cast<const uint4*>(dy + row * hidden);
  const uint4* xr = reinterpret_cast<const uint4*>(x + row * hidden);
  const bool rms = mean == nullptr;
  const float mu = rms ? 0.f : mean[row], rs = rstd[row];
  uint4 bg[MAXV], bx[MAXV];  // g = dy*gamma (kept as bf16-packed? no: keep dy and x packed, recompute)
  float c1 = 0.f, c2 = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int idx = lane + 32 * i;
    if (idx < nvec) {
      bg[i] = __ldg(dyr + idx);
      bx[i] = __ldg(xr + idx);
      float fd[8], fx[8]; unpack8(bg[i], fd); unpack8(bx[i], fx);
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma) + idx * 2), g1 = __ldg(reinterpret_cast<const float4*>(gamma) + idx * 2 + 1);
      const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float gdy = fd[e] * gg[e];
        c1 += gdy;
        c2 += gdy * (fx[e] - mu) * rs;
      }
    }
  }
  c1 = rms ? 0.f : warp_sum(c1) / hidden;   // RMSNorm has no mean term
  c2 = warp_sum(c2) / hidden;
  uint4* dxr = reinterpret_cast<uint4*>(dx + row * hidden);
  const uint4* rr = d_res ? reinterpret_cast<const uint4*>(d_res + row * hidden) : nullptr;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int idx = lane + 32 * i;
    if (idx < nvec) {
      float fd[8], fx[8], out[8]; unpack8(bg[i], fd); unpack8(bx[i], fx);
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma) + idx * 2), g1 = __ldg(reinterpret_cast<const float4*>(gamma) + idx * 2 + 1);
      const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
      for (int e = 0; e < 8; ++e) out[e] = rs * (fd[e] * gg[e] - c1 - (fx[e] - mu) * rs * c2);
      if (rr) {
        float fr[8]; const uint4 r = __ldg(rr + idx); unpack8(r, fr);
#pragma unroll
        for (int e = 0; e < 8; ++e) out[e] += fr[e];
      }
      dxr[idx] = pack8(out);
    }
  }
}

// ------------------------------------------------------------------------------------ staged (TMA bulk) norm backward
// The register-resident backward above keeps a row in flight only while its warp is loading, and fetches d_res after its
// two warp reductions.  Here one producer lane streams tiles of 8 rows of dy, x (and d_res) into a shared-memory ring
// with cp.async.bulk (up to ~190 KB in flight per SM, independent of what the consumer warps are doing) and two
// consumer warps share one row of the tile (partial sums meet in shared memory behind a 64-thread named barrier); persistent CTAs, one per SM.  Same arithmetic.  Measured alone on B200,
// [10240, 2048], L2 flushed: 47.1 -> 41.0 us, with d_res 67.6 -> 47.1 us.  The forward in the same form was SLOWER
// (34.8 us with one warp per row, 32.8 us with two, vs 26.6 us), so it stays register-resident.
constexpr int kNormRows = 8;                 // rows per tile; two consumer warps per row (alternate 512-byte chunks)
constexpr int kNormWarps = 2 * kNormRows;
constexpr int kNormThreads = 32 * (kNormWarps + 1);
constexpr int kNormSmemBudget = 200 * 1024;

template <bool kRes>
__global__ void __launch_bounds__(kNormThreads, 1)
norm_bwd_staged_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                       const float* __restrict__ gamma, const float* __restrict__ mean, const float* __restrict__ rstd,
                       const __nv_bfloat16* __restrict__ d_res, __nv_bfloat16* __restrict__ dx, int64_t rows, int hidden,
                       int stages) {
  extern __shared__ __align__(128) uint8_t norm_smem[];
  constexpr int kTensors = kRes ? 3 : 2;
  const int row_bytes = hidden * 2, tile_bytes = kNormRows * row_bytes, stage_bytes = kTensors * tile_bytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(norm_smem + (size_t)stages * stage_bytes);
  uint64_t* empty = full + stages;
  float* xch = reinterpret_cast<float*>(empty + stages);   // [2 tile parities][rows][2 halves][c1, c2]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], kNormWarps); }
    fence_barrier_init();
  }
  __syncthreads();
  pdl_launch();
  pdl_wait();
  const int64_t num_tiles = (rows + kNormRows - 1) / kNormRows;
  int s = 0; uint32_t ph = 0;
  if (warp == kNormWarps) {
    if (lane == 0) {
      for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&empty[s], ph ^ 1);
        const int64_t r0 = tile * kNormRows;
        const uint32_t bytes = (uint32_t)min((int64_t)kNormRows, rows - r0) * (uint32_t)row_bytes;
        uint8_t* dst = norm_smem + (size_t)s * stage_bytes;
        mbar_arrive_expect_tx(&full[s], kTensors * bytes);
        bulk_load_1d(dst, dy + r0 * hidden, bytes, &full[s]);
        bulk_load_1d(dst + tile_bytes, x + r0 * hidden, bytes, &full[s]);
        if (kRes) bulk_load_1d(dst + 2 * tile_bytes, d_res + r0 * hidden, bytes, &full[s]);
        if (++s == stages) { s = 0; ph ^= 1; }
      }
    }
    return;
  }
  const bool rms = mean == nullptr;
  const int nvec = hidden >> 3;
  const float inv_h = 1.f / (float)hidden;
  const int trow = warp >> 1, half = warp & 1;
  int par = 0;
  for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, par ^= 1) {
    const int64_t row = tile * kNormRows + trow;
    const uint8_t* st = norm_smem + (size_t)s * stage_bytes + (size_t)trow * row_bytes;
    const uint4* sdy = reinterpret_cast<const uint4*>(st);
    const uint4* sx = reinterpret_cast<const uint4*>(st + tile_bytes);
    const uint4* sr = reinterpret_cast<const uint4*>(st + 2 * tile_bytes);
    float mu = 0.f, rs = 0.f;
    if (row < rows) { mu = rms ? 0.f : __ldg(mean + row); rs = __ldg(rstd + row); }
    mbar_wait(&full[s], ph);
    if (row < rows) {
      float c1 = 0.f, c2 = 0.f;
      for (int idx = lane + 32 * half; idx < nvec; idx += 64) {
        float fd[8], fx[8]; unpack8(sdy[idx], fd); unpack8(sx[idx], fx);
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma) + idx * 2), g1 = __ldg(reinterpret_cast<const float4*>(gamma) + idx * 2 + 1);
        const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float gdy = fd[e] * gg[e];
          c1 += gdy;
          c2 += gdy * (fx[e] - mu) * rs;
        }
      }
      c1 = warp_sum(c1);
      c2 = warp_sum(c2);
      float* xr = xch + ((par * kNormRows + trow) * 2) * 2;     // the two halves of the row meet here
      if (lane == 0) { xr[half * 2] = c1; xr[half * 2 + 1] = c2; }
      asm volatile("bar.sync %0, 64;" ::"r"(1 + trow) : "memory");
      c1 = rms ? 0.f : (xr[0] + xr[2]) * inv_h;
      c2 = (xr[1] + xr[3]) * inv_h;
      uint4* dxr = reinterpret_cast<uint4*>(dx + row * hidden);
      for (int idx = lane + 32 * half; idx < nvec; idx += 64) {
        float fd[8], fx[8], out[8]; unpack8(sdy[idx], fd); unpack8(sx[idx], fx);
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma) + idx * 2), g1 = __ldg(reinterpret_cast<const float4*>(gamma) + idx * 2 + 1);
        const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
        for (int e = 0; e < 8; ++e) out[e] = rs * (fd[e] * gg[e] - c1 - (fx[e] - mu) * rs * c2);
        if (kRes) {
          float fr[8]; unpack8(sr[idx], fr);
#pragma unroll
          for (int e = 0; e < 8; ++e) out[e] += fr[e];
        }
        dxr[idx] = pack8(out);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
    if (++s == stages) { s = 0; ph ^= 1; }
  }
}

// ------------------------------------------------------------------------------------ column reductions
// mode 0: partial[c][n] = sum_{rows in chunk c} x[m,n]
// mode 1: partial[c][n] = sum dy[m,n] * (x[m,n]-mean[m])*rstd[m]   (LayerNorm dgamma); dbeta uses mode 0 on dy
// block = 256 threads x 8 columns = 2048 columns; grid (ceil(n/2048), chunks)
template <int MODE>
__global__ void __launch_bounds__(256)
col_partial_kernel(const __nv_bfloat16* __restrict__ a, int64_t lda, const __nv_bfloat16* __restrict__ x, int64_t ldx,
                   const float* __restrict__ mean, const float* __restrict__ rstd, float* __restrict__ partial,
                   int64_t m, int64_t n, int64_t rows_per_chunk) {
  pdl_launch();
  pdl_wait();
  const int64_t col = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 8;
  if (col >= n) return;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_chunk;
  const int64_t r1 = min(m, r0 + rows_per_chunk);
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const bool full = col + 8 <= n;
  int64_t r = r0;
  if (full && MODE == 0) {
    // four rows per trip: independent 16-byte loads in flight (this kernel is pure bandwidth; one load per trip left it
    // latency-bound at ~1 TB/s)
    for (; r + 4 <= r1; r += 4) {
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = __ldg(reinterpret_cast<const uint4*>(a + (r + u) * lda + col));
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float f[8]; unpack8(v[u], f);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] += f[e];
      }
    }
  } else if (full) {
    for (; r + 2 <= r1; r += 2) {
      uint4 va[2], vx[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        va[u] = __ldg(reinterpret_cast<const uint4*>(a + (r + u) * lda + col));
        vx[u] = __ldg(reinterpret_cast<const uint4*>(x + (r + u) * ldx + col));
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        float fa2[8], fx2[8]; unpack8(va[u], fa2); unpack8(vx[u], fx2);
        const float mu = mean != nullptr ? mean[r + u] : 0.f, rs = rstd[r + u];
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] += fa2[e] * (fx2[e] - mu) * rs;
      }
    }
  }
  for (; r < r1; ++r) {
    float fa[8];
    if (full) {
      unpack8(__ldg(reinterpret_cast<const uint4*>(a + r * lda + col)), fa);
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) fa[e] = (col + e < n) ? __bfloat162float(a[r * lda + col + e]) : 0.f;
    }
    if (MODE == 1) {
      float fx[8];
      if (full) {
        unpack8(__ldg(reinterpret_cast<const uint4*>(x + r * ldx + col)), fx);
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) fx[e] = (col + e < n) ? __bfloat162float(x[r * ldx + col + e]) : 0.f;
      }
      const float mu = mean != nullptr ? mean[r] : 0.f, rs = rstd[r];
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] += fa[e] * (fx[e] - mu) * rs;
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] += fa[e];
    }
  }
  float* p = partial + (int64_t)blockIdx.y * n + col;
#pragma unroll
  for (int e = 0; e < 8; ++e) if (col + e < n) p[e] = acc[e];
}

// block = 32 columns x 8 chunk lanes: lane y sums chunks y, y+8, ... (coalesced 128-byte reads), then a fixed-order
// shared-memory reduction over the 8 lanes -> deterministic
__global__ void __launch_bounds__(256)
col_final_kernel(const float* __restrict__ partial, int chunks, int64_t n, float scale, const float* __restrict__ gate,
                 float* __restrict__ out, int accumulate) {
  pdl_launch();
  pdl_wait();
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t col = (int64_t)blockIdx.x * 32 + tx;
  float s = 0.f;
  if (col < n)
    for (int c = ty; c < chunks; c += 8) s += partial[(int64_t)c * n + col];
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && col < n) {
    float t = 0.f;
#pragma unroll
    for (int y = 0; y < 8; ++y) t += red[y][tx];
    if (gate != nullptr) t *= tanh_precise(__ldg(gate));
    t *= scale;
    out[col] = accumulate ? out[col] + t : t;
  }
}

// dot over all elements of two [m,n] bf16 matrices -> per-block partials -> scalar
__global__ void __launch_bounds__(256)
dot_partial_kernel(const __nv_bfloat16* __restrict__ a, int64_t lda, const __nv_bfloat16* __restrict__ b, int64_t ldb,
                   int64_t m, int64_t n, float* __restrict__ partial) {
  pdl_launch();
  pdl_wait();
  const int64_t nvec = n >> 3;  // host guarantees n % 8 == 0
  float acc = 0.f;
  for (int64_t r = blockIdx.x; r < m; r += gridDim.x) {
    const uint4* ar = reinterpret_cast<const uint4*>(a + r * lda);
    const uint4* br = reinterpret_cast<const uint4*>(b + r * ldb);
    for (int64_t i = threadIdx.x; i < nvec; i += 256) {
      float fa[8], fb[8]; unpack8(__ldg(ar + i), fa); unpack8(__ldg(br + i), fb);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc += fa[e] * fb[e];
    }
  }
  __shared__ float red[8];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < 8 ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) partial[blockIdx.x] = v;
  }
}
__global__ void __launch_bounds__(256)
gate_final_kernel(const float* __restrict__ partial, int blocks, const float* __restrict__ gate, float* __restrict__ out,
                  int accumulate) {
  pdl_launch();
  pdl_wait();
  float acc = 0.f;
  for (int i = threadIdx.x; i < blocks; i += 256) acc += partial[i];
  __shared__ float red[8];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = 0.f;
    for (int i = 0; i < 8; ++i) v += red[i];
    const float t = tanh_precise(__ldg(gate));
    v *= (1.f - t * t);
    out[0] = accumulate ? out[0] + v : v;
  }
}

// ------------------------------------------------------------------------------------ softmax cross-entropy
// One CTA per row: single pass online (max, sum) over the bf16 logits in fp32, 16-byte loads.
__device__ __forceinline__ void online_merge(float& m, float& s, float m2, float s2) {
  const float mx = fmaxf(m, m2);
  if (mx == -INFINITY) return;  // both empty: exp(-inf - -inf) would be NaN
  s = s * __expf(m - mx) + s2 * __expf(m2 - mx);
  m = mx;
}
__global__ void __launch_bounds__(256)
ce_fwd_kernel(const __nv_bfloat16* __restrict__ logits, int64_t ld, const int64_t* __restrict__ labels, int64_t vocab,
              int64_t ignore_index, float* __restrict__ lse, float* __restrict__ row_loss, int vec) {
  const int64_t row = blockIdx.x;
  const __nv_bfloat16* x = logits + row * ld;
  float m = -INFINITY, s = 0.f;
  if (vec) {
    const int64_t nvec = vocab >> 3;
    for (int64_t i = threadIdx.x; i < nvec; i += 256) {
      float f[8]; unpack8(__ldg(reinterpret_cast<const uint4*>(x) + i), f);
      float mx = f[0];
#pragma unroll
      for (int e = 1; e < 8; ++e) mx = fmaxf(mx, f[e]);
      float part = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) part += __expf(f[e] - mx);
      online_merge(m, s, mx, part);
    }
    for (int64_t c = (nvec << 3) + threadIdx.x; c < vocab; c += 256) online_merge(m, s, __bfloat162float(x[c]), 1.f);
  } else {
    for (int64_t c = threadIdx.x; c < vocab; c += 256) online_merge(m, s, __bfloat162float(x[c]), 1.f);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
    online_merge(m, s, m2, s2);
  }
  __shared__ float sm[8], ss[8];
  if ((threadIdx.x & 31) == 0) { sm[threadIdx.x >> 5] = m; ss[threadIdx.x >> 5] = s; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float M = sm[0], S = ss[0];
    for (int w = 1; w < 8; ++w) online_merge(M, S, sm[w], ss[w]);
    const float l = M + logf(S);
    lse[row] = l;
    const int64_t lab = labels[row];
    const bool valid = lab != ignore_index && lab >= 0 && lab < vocab;
    row_loss[row] = valid ? l - __bfloat162float(x[lab]) : 0.f;
  }
}
// loss = sum(row_loss) / max(1, #valid); deterministic single-CTA tree reduction
__global__ void __launch_bounds__(1024)
ce_final_kernel(const float* __restrict__ row_loss, const int64_t* __restrict__ labels, int64_t rows, int64_t vocab,
                int64_t ignore_index, float* __restrict__ loss, float* __restrict__ count) {
  float s = 0.f, c = 0.f;
  for (int64_t r = threadIdx.x; r < rows; r += 1024) {
    const int64_t lab = labels[r];
    if (lab != ignore_index && lab >= 0 && lab < vocab) { s += row_loss[r]; c += 1.f; }
  }
  __shared__ float rs[32], rc[32];
  s = warp_sum(s); c = warp_sum(c);
  if ((threadIdx.x & 31) == 0) { rs[threadIdx.x >> 5] = s; rc[threadIdx.x >> 5] = c; }
  __syncthreads();
  if (threadIdx.x < 32) {
    s = warp_sum(rs[threadIdx.x]); c = warp_sum(rc[threadIdx.x]);
    if (threadIdx.x == 0) { count[0] = c; loss[0] = s / fmaxf(c, 1.f); }
  }
}
// dlogits = (softmax - onehot) * dloss / count, grid (rows, column chunks of 2048)
__global__ void __launch_bounds__(256)
ce_bwd_kernel(const __nv_bfloat16* __restrict__ logits, int64_t ld, const int64_t* __restrict__ labels,
              const float* __restrict__ lse, const float* __restrict__ dloss, const float* __restrict__ count,
              __nv_bfloat16* __restrict__ dlogits, int64_t ldd, int64_t vocab, int64_t ignore_index, int vec) {
  const int64_t row = blockIdx.x;
  const int64_t lab = labels[row];
  const bool valid = lab != ignore_index && lab >= 0 && lab < vocab;
  const float g = valid ? __ldg(dloss) / fmaxf(__ldg(count), 1.f) : 0.f;
  const float l = lse[row];
  const __nv_bfloat16* x = logits + row * ld;
  __nv_bfloat16* d = dlogits + row * ldd;
  const int64_t col = ((int64_t)blockIdx.y * 256 + threadIdx.x) * 8;
  if (col >= vocab) return;
  if (vec && col + 8 <= vocab) {
    float f[8]; unpack8(__ldg(reinterpret_cast<const uint4*>(x + col)), f);
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] = (__expf(f[e] - l) - ((col + e == lab) ? 1.f : 0.f)) * g;
    *reinterpret_cast<uint4*>(d + col) = pack8(f);
  } else {
    for (int e = 0; e < 8 && col + e < vocab; ++e)
      d[col + e] = __float2bfloat16_rn((__expf(__bfloat162float(x[col + e]) - l) - ((col + e == lab) ? 1.f : 0.f)) * g);
  }
}

// ------------------------------------------------------------------------------------ dropout re-application
// out = keep ? x * scale : 0 with the GEMM epilogue's counter-based mask.  One thread per 8-column group.
__global__ void __launch_bounds__(256)
dropout_apply_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, __nv_bfloat16* __restrict__ out, int64_t ldo,
                     int64_t m, int64_t n, uint32_t thresh, float scale, uint64_t seed, int vec) {
  pdl_launch();
  pdl_wait();
  const int64_t groups = (n + 7) / 8;
  const int64_t total = m * groups;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int64_t row = i / groups, g = i % groups, col = g * 8;
    const DropBits bits = dropout_bits(seed, row, g, groups);
    if (vec && col + 8 <= n) {
      float f[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(x + row * ldx + col)), f);
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] = dropout_keep(bits, e, thresh) ? f[e] * scale : 0.f;
      *reinterpret_cast<uint4*>(out + row * ldo + col) = pack8(f);
    } else {
      for (int e = 0; e < 8 && col + e < n; ++e) {
        const float v = __bfloat162float(x[row * ldx + col + e]);
        out[row * ldo + col + e] = __float2bfloat16_rn(dropout_keep(bits, e, thresh) ? v * scale : 0.f);
      }
    }
  }
}

// ------------------------------------------------------------------------------------ neighbor bank
// grid (batch*(T+I), ceil(row_width/2048)), 256 threads x 8 columns.
__global__ void __launch_bounds__(256)
bank_pack_fwd_kernel(const mmgl_bank_args a) {
  const int n_src = (int)(a.n_text + a.n_image);
  const int b = blockIdx.x / n_src, j = blockIdx.x % n_src;
  const bool is_text = j < a.n_text;
  const int jj = is_text ? j : j - (int)a.n_text;
  const int64_t src_row = is_text ? (int64_t)b * a.n_text + jj : (int64_t)b * a.n_image + jj;
  const int64_t pos = is_text ? a.text_pos_ids[src_row] : a.image_pos_ids[src_row];
  const int64_t loc = is_text ? a.text_locations[src_row] : a.image_locations[src_row];
  if (loc < 0 || loc >= n_src) return;
  const __nv_bfloat16* proj = reinterpret_cast<const __nv_bfloat16*>(is_text ? a.text_proj : a.image_proj) + src_row * a.row_width;
  const __nv_bfloat16* table = reinterpret_cast<const __nv_bfloat16*>(is_text ? a.text_pos_table : a.image_pos_table);
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(a.bank) + ((int64_t)b * n_src + loc) * a.row_width;

  __shared__ float s_lpe[64];
  if (a.lpe != nullptr) {
    if (threadIdx.x < a.lpe_k) s_lpe[threadIdx.x] = a.lpe[((int64_t)b * (n_src + 1) + loc + 1) * a.lpe_k + threadIdx.x];
    __syncthreads();
  }
  const int64_t col = ((int64_t)blockIdx.y * 256 + threadIdx.x) * 8;
  if (col < a.row_width) {
    float f[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(proj + col)), f);
    if (table != nullptr) {
      float tp[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(table + pos * a.row_width + col)), tp);
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] += tp[e];
    }
    if (a.lpe != nullptr) {
      const __nv_bfloat16* w = reinterpret_cast<const __nv_bfloat16*>(a.lpe_weight);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float acc = a.lpe_bias ? a.lpe_bias[col + e] : 0.f;
        const __nv_bfloat16* wr = w + (col + e) * a.lpe_k;
        for (int kk = 0; kk < a.lpe_k; ++kk) acc += s_lpe[kk] * __bfloat162float(wr[kk]);
        f[e] += acc;
      }
    }
    *reinterpret_cast<uint4*>(out + col) = pack8(f);
  }
  if (blockIdx.y == 0 && threadIdx.x < a.n_tok) a.mask[((int64_t)b * n_src + loc) * a.n_tok + threadIdx.x] = pos > 0 ? 1 : 0;
}

// gather rows of d_bank back to the projection outputs
__global__ void __launch_bounds__(256)
bank_pack_bwd_proj_kernel(const mmgl_bank_bwd_args a) {
  const int n_src = (int)(a.n_text + a.n_image);
  const int b = blockIdx.x / n_src, j = blockIdx.x % n_src;
  const bool is_text = j < a.n_text;
  const int jj = is_text ? j : j - (int)a.n_text;
  const int64_t src_row = is_text ? (int64_t)b * a.n_text + jj : (int64_t)b * a.n_image + jj;
  const int64_t loc = is_text ? a.text_locations[src_row] : a.image_locations[src_row];
  __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(is_text ? a.d_text_proj : a.d_image_proj);
  if (dst == nullptr) return;
  dst += src_row * a.row_width;
  const int64_t col = ((int64_t)blockIdx.y * 256 + threadIdx.x) * 8;
  if (col >= a.row_width) return;
  uint4 v = make_uint4(0, 0, 0, 0);
  if (loc >= 0 && loc < n_src)
    v = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(a.d_bank) + ((int64_t)b * n_src + loc) * a.row_width + col));
  *reinterpret_cast<uint4*>(dst + col) = v;
}

// d_pos_table[p,:] += sum over sources with pos_id == p of d_bank rows.  grid (pos_rows, col chunks); deterministic.
__global__ void __launch_bounds__(256)
bank_pack_bwd_pos_kernel(const __nv_bfloat16* __restrict__ d_bank, const int64_t* __restrict__ pos_ids,
                         const int64_t* __restrict__ locations, int64_t batch, int64_t n_mine, int64_t n_src,
                         int64_t row_width, float* __restrict__ d_table) {
  const int64_t p = blockIdx.x;
  const int64_t col = ((int64_t)blockIdx.y * 256 + threadIdx.x) * 8;
  if (col >= row_width) return;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  bool any = false;
  for (int64_t i = 0; i < batch * n_mine; ++i) {
    if (pos_ids[i] != p) continue;
    const int64_t loc = locations[i];
    if (loc < 0 || loc >= n_src) continue;
    const int64_t b = i / n_mine;
    float f[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(d_bank + (b * n_src + loc) * row_width + col)), f);
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] += f[e];
    any = true;
  }
  if (any) {
    float* o = d_table + p * row_width + col;
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] += acc[e];
  }
}

// d_lpe_weight[c,kk] += sum_{b, slot} d_bank[b,slot,c] * lpe[b,slot+1,kk];  d_lpe_bias[c] += sum d_bank[b,slot,c]
// one thread per column c, k <= 32 accumulators.
__global__ void __launch_bounds__(128)
bank_pack_bwd_lpe_kernel(const __nv_bfloat16* __restrict__ d_bank, const float* __restrict__ lpe, int64_t batch,
                         int64_t n_src, int64_t row_width, int lpe_k, float* __restrict__ d_w, float* __restrict__ d_b) {
  const int64_t c = (int64_t)blockIdx.x * 128 + threadIdx.x;
  if (c >= row_width) return;
  float acc[32];
#pragma unroll
  for (int kk = 0; kk < 32; ++kk) acc[kk] = 0.f;
  float accb = 0.f;
  for (int64_t b = 0; b < batch; ++b)
    for (int64_t s = 0; s < n_src; ++s) {
      const float g = __bfloat162float(d_bank[(b * n_src + s) * row_width + c]);
      const float* lr = lpe + (b * (n_src + 1) + s + 1) * lpe_k;
      accb += g;
#pragma unroll
      for (int kk = 0; kk < 32; ++kk) if (kk < lpe_k) acc[kk] += g * __ldg(lr + kk);
    }
  if (d_w)
#pragma unroll
    for (int kk = 0; kk < 32; ++kk) if (kk < lpe_k) d_w[c * lpe_k + kk] += acc[kk];
  if (d_b) d_b[c] += accb;
}

// ------------------------------------------------------------------------------------ GCN helpers
// grid (batch, ceil(dim/256)), 256 threads = 256 columns; smem: adj [nodes*nodes] + tile [nodes][256] fp32
__global__ void __launch_bounds__(256)
gcn_concat_fwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ adj, __nv_bfloat16* __restrict__ out,
                      int nodes, int64_t dim, int prepend_root) {
  extern __shared__ float sm[];
  float* s_adj = sm;
  float* s_x = sm + nodes * nodes;
  const int b = blockIdx.x;
  const int64_t c = (int64_t)blockIdx.y * 256 + threadIdx.x;
  for (int i = threadIdx.x; i < nodes * nodes; i += 256) s_adj[i] = adj[(int64_t)b * nodes * nodes + i];
  const int x_nodes = prepend_root ? nodes - 1 : nodes;
  for (int j = 0; j < nodes; ++j) {
    float v = 0.f;
    const int xj = prepend_root ? j - 1 : j;
    if (c < dim && xj >= 0) v = __bfloat162float(x[((int64_t)b * x_nodes + xj) * dim + c]);
    s_x[j * 256 + threadIdx.x] = v;
  }
  __syncthreads();
  if (c >= dim) return;
  for (int i = 0; i < nodes; ++i) {
    float agg = 0.f;
    for (int j = 0; j < nodes; ++j) agg += s_adj[i * nodes + j] * s_x[j * 256 + threadIdx.x];
    __nv_bfloat16* o = out + ((int64_t)b * nodes + i) * 2 * dim;
    o[c] = __float2bfloat16_rn(s_x[i * 256 + threadIdx.x]);
    o[dim + c] = __float2bfloat16_rn(agg);
  }
}

__global__ void __launch_bounds__(256)
gcn_combine_bwd_kernel(const __nv_bfloat16* __restrict__ dc, const float* __restrict__ adj,
                       const __nv_bfloat16* __restrict__ relu_mask, __nv_bfloat16* __restrict__ dx, int nodes,
                       int64_t dim, int drop_root) {
  extern __shared__ float sm[];
  float* s_adj = sm;
  float* s_g = sm + nodes * nodes;
  const int b = blockIdx.x;
  const int64_t c = (int64_t)blockIdx.y * 256 + threadIdx.x;
  for (int i = threadIdx.x; i < nodes * nodes; i += 256) s_adj[i] = adj[(int64_t)b * nodes * nodes + i];
  for (int i = 0; i < nodes; ++i)
    s_g[i * 256 + threadIdx.x] = (c < dim) ? __bfloat162float(dc[((int64_t)b * nodes + i) * 2 * dim + dim + c]) : 0.f;
  __syncthreads();
  if (c >= dim) return;
  const int out_nodes = drop_root ? nodes - 1 : nodes;
  for (int j = drop_root ? 1 : 0; j < nodes; ++j) {
    float v = __bfloat162float(dc[((int64_t)b * nodes + j) * 2 * dim + c]);
    for (int i = 0; i < nodes; ++i) v += s_adj[i * nodes + j] * s_g[i * 256 + threadIdx.x];
    if (relu_mask != nullptr && !(__bfloat162float(relu_mask[((int64_t)b * nodes + j) * dim + c]) > 0.f)) v = 0.f;
    const int oj = drop_root ? j - 1 : j;
    dx[((int64_t)b * out_nodes + oj) * dim + c] = __float2bfloat16_rn(v);
  }
}

static int reduce_chunks(int64_t m) {
  int64_t c = (m + 15) / 16;   // 16 rows per chunk: >= 2 CTAs per SM at the step's M = 5120
  if (c > 1024) c = 1024;
  if (c < 1) c = 1;
  return (int)c;
}

}  // namespace mmgl

using namespace mmgl;

// Stages of the shared-memory ring for the staged norm backward, 0 = use the register-resident kernel (few rows: one
// CTA per SM would not be busy; MMGL_NORM_STAGED=0 switches the staged kernel off for A/B measurements).
static int norm_stages(int64_t rows, int64_t hidden, int tensors) {
  static const bool enabled = [] { const char* e = getenv("MMGL_NORM_STAGED"); return !(e && e[0] == '0'); }();
  if (!enabled || rows < (int64_t)kNormRows * 4 * sm_count()) return 0;
  const int64_t stage_bytes = (int64_t)tensors * kNormRows * hidden * 2;
  int64_t stages = kNormSmemBudget / stage_bytes;
  if (stages > 6) stages = 6;
  return stages >= 2 ? (int)stages : 0;
}

static int norm_fwd(const char* who, const void* x, const float* gamma, const float* beta, void* y, float* mean,
                    float* rstd, int64_t rows, int64_t hidden, float eps, cudaStream_t s) {
  MMGL_REQUIRE(rows > 0 && hidden > 0 && hidden % 8 == 0 && hidden <= 8192,
               "%s: hidden must be a multiple of 8 and <= 8192 (got %lld)", who, (long long)hidden);
  MMGL_REQUIRE(aligned16(x) && aligned16(y) && aligned16(gamma) && (!beta || aligned16(beta)), "%s: unaligned", who);
  const unsigned grid = (unsigned)((rows + 7) / 8);
  const auto X = (const __nv_bfloat16*)x; auto Y = (__nv_bfloat16*)y;
  auto kern = hidden <= 1024 ? layernorm_fwd_kernel<4> : hidden <= 2048 ? layernorm_fwd_kernel<8>
              : hidden <= 4096 ? layernorm_fwd_kernel<16> : layernorm_fwd_kernel<32>;
  MMGL_CUDA(launch_pdl(kern, dim3(grid), dim3(256), 0, s, X, gamma, beta, Y, mean, rstd, rows, (int)hidden, eps));
  return check_launch(who);
}

extern "C" int mmgl_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean,
                                  float* rstd, int64_t rows, int64_t hidden, float eps, void* stream_) {
  MMGL_BIND(x, "mmgl_layernorm_fwd");
  MMGL_REQUIRE(x && gamma && beta && y && mean && rstd, "mmgl_layernorm_fwd: null pointer");
  return norm_fwd("mmgl_layernorm_fwd", x, gamma, beta, y, mean, rstd, rows, hidden, eps, reinterpret_cast<cudaStream_t>(stream_));
}

extern "C" int mmgl_rmsnorm_fwd(const void* x, const float* gamma, void* y, float* rstd, int64_t rows, int64_t hidden,
                                float eps, void* stream_) {
  MMGL_BIND(x, "mmgl_rmsnorm_fwd");
  MMGL_REQUIRE(x && gamma && y && rstd, "mmgl_rmsnorm_fwd: null pointer");
  return norm_fwd("mmgl_rmsnorm_fwd", x, gamma, nullptr, y, nullptr, rstd, rows, hidden, eps, reinterpret_cast<cudaStream_t>(stream_));
}

extern "C" size_t mmgl_reduce_workspace_bytes(int64_t m, int64_t n) {
  const size_t a = (size_t)reduce_chunks(m) * (size_t)n * sizeof(float);
  return a > 4096 ? a : 4096;
}
extern "C" size_t mmgl_layernorm_bwd_workspace_bytes(int64_t rows, int64_t hidden) {
  return mmgl_reduce_workspace_bytes(rows, hidden);
}

static int norm_bwd(const char* who, const void* dy, const void* x, const float* gamma, const float* mean,
                    const float* rstd, const void* d_res, void* dx, float* dgamma, float* dbeta, int32_t accumulate,
                    void* workspace, size_t workspace_bytes, int64_t rows, int64_t hidden, cudaStream_t s) {
  MMGL_REQUIRE(rows > 0 && hidden > 0 && hidden % 8 == 0 && hidden <= 4096,
               "%s: hidden must be a multiple of 8 and <= 4096 (got %lld)", who, (long long)hidden);
  MMGL_REQUIRE(aligned16(dy) && aligned16(x) && aligned16(dx) && aligned16(gamma) && (!d_res || aligned16(d_res)),
               "%s: unaligned", who);
  const unsigned grid = (unsigned)((rows + 7) / 8);
  const auto DY = (const __nv_bfloat16*)dy; const auto X = (const __nv_bfloat16*)x;
  const auto R = (const __nv_bfloat16*)d_res; auto DX = (__nv_bfloat16*)dx;
  if (int stages = norm_stages(rows, hidden, R ? 3 : 2)) {
    const size_t smem = (size_t)stages * (R ? 3 : 2) * kNormRows * hidden * 2 + 2 * stages * sizeof(uint64_t) +
                        2 * kNormRows * 4 * sizeof(float);
    static std::once_flag once;
    std::call_once(once, [] {
      cudaFuncSetAttribute(norm_bwd_staged_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kNormSmemBudget + 1024);
      cudaFuncSetAttribute(norm_bwd_staged_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kNormSmemBudget + 1024);
    });
    const unsigned g = (unsigned)std::min<int64_t>((rows + kNormRows - 1) / kNormRows, sm_count());
    MMGL_CUDA(launch_pdl(R ? norm_bwd_staged_kernel<true> : norm_bwd_staged_kernel<false>, dim3(g), dim3(kNormThreads), smem, s, DY, X,
                         gamma, mean, rstd, R, DX, rows, (int)hidden, stages));
  } else
  if (hidden <= 1024) layernorm_bwd_dx_kernel<4><<<grid, 256, 0, s>>>(DY, X, gamma, mean, rstd, R, DX, rows, (int)hidden);
  else if (hidden <= 2048) layernorm_bwd_dx_kernel<8><<<grid, 256, 0, s>>>(DY, X, gamma, mean, rstd, R, DX, rows, (int)hidden);
  else layernorm_bwd_dx_kernel<16><<<grid, 256, 0, s>>>(DY, X, gamma, mean, rstd, R, DX, rows, (int)hidden);
  if (int rc = check_launch(who)) return rc;
  if (dgamma != nullptr || dbeta != nullptr) {
    MMGL_REQUIRE(workspace && workspace_bytes >= mmgl_layernorm_bwd_workspace_bytes(rows, hidden), "%s: workspace too small", who);
    const int chunks = reduce_chunks(rows);
    const int64_t rpc = (rows + chunks - 1) / chunks;
    dim3 g((unsigned)((hidden + 2047) / 2048), (unsigned)chunks);
    float* ws = reinterpret_cast<float*>(workspace);
    if (dgamma) {
      MMGL_CUDA(launch_pdl(col_partial_kernel<1>, g, dim3(256), 0, s, DY, hidden, X, hidden, mean, rstd, ws, rows, hidden, rpc));
      if (int rc = check_launch(who)) return rc;
      MMGL_CUDA(launch_pdl(col_final_kernel, dim3((unsigned)((hidden + 31) / 32)), dim3(256), 0, s, ws, chunks, hidden, 1.f, nullptr, dgamma, accumulate));
      if (int rc = check_launch(who)) return rc;
    }
    if (dbeta) {
      MMGL_CUDA(launch_pdl(col_partial_kernel<0>, g, dim3(256), 0, s, DY, hidden, nullptr, 0, nullptr, nullptr, ws, rows, hidden, rpc));
      if (int rc = check_launch(who)) return rc;
      MMGL_CUDA(launch_pdl(col_final_kernel, dim3((unsigned)((hidden + 31) / 32)), dim3(256), 0, s, ws, chunks, hidden, 1.f, nullptr, dbeta, accumulate));
      if (int rc = check_launch(who)) return rc;
    }
  }
  return 0;
}

extern "C" int mmgl_layernorm_bwd(const void* dy, const void* x, const float* gamma, const float* mean,
                                  const float* rstd, const void* d_res, void* dx, float* dgamma, float* dbeta,
                                  int32_t accumulate, void* workspace, size_t workspace_bytes, int64_t rows,
                                  int64_t hidden, void* stream_) {
  MMGL_BIND(x, "mmgl_layernorm_bwd");
  MMGL_REQUIRE(dy && x && gamma && mean && rstd && dx, "mmgl_layernorm_bwd: null pointer");
  return norm_bwd("mmgl_layernorm_bwd", dy, x, gamma, mean, rstd, d_res, dx, dgamma, dbeta, accumulate, workspace,
                  workspace_bytes, rows, hidden, reinterpret_cast<cudaStream_t>(stream_));
}

extern "C" int mmgl_rmsnorm_bwd(const void* dy, const void* x, const float* gamma, const float* rstd, const void* d_res,
                                void* dx, float* dgamma, int32_t accumulate, void* workspace, size_t workspace_bytes,
                                int64_t rows, int64_t hidden, void* stream_) {
  MMGL_BIND(x, "mmgl_rmsnorm_bwd");
  MMGL_REQUIRE(dy && x && gamma && rstd && dx, "mmgl_rmsnorm_bwd: null pointer");
  return norm_bwd("mmgl_rmsnorm_bwd", dy, x, gamma, nullptr, rstd, d_res, dx, dgamma, nullptr, accumulate, workspace,
                  workspace_bytes, rows, hidden, reinterpret_cast<cudaStream_t>(stream_));
}

extern "C" int mmgl_colsum(const void* x, int64_t ldx, int64_t m, int64_t n, float scale, const float* gate,
                           float* out, int32_t accumulate, void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  MMGL_BIND(x, "mmgl_colsum");
  MMGL_REQUIRE(x && out && m > 0 && n > 0, "mmgl_colsum: bad arguments");
  MMGL_REQUIRE(aligned16(x) && ldx % 8 == 0, "mmgl_colsum: x must be 16B aligned with ld %% 8 == 0");
  MMGL_REQUIRE(workspace && workspace_bytes >= mmgl_reduce_workspace_bytes(m, n), "mmgl_colsum: workspace too small");
  const int chunks = reduce_chunks(m);
  const int64_t rpc = (m + chunks - 1) / chunks;
  dim3 g((unsigned)((n + 2047) / 2048), (unsigned)chunks);
  float* ws = reinterpret_cast<float*>(workspace);
  MMGL_CUDA(launch_pdl(col_partial_kernel<0>, g, dim3(256), 0, s, (const __nv_bfloat16*)x, ldx, nullptr, 0, nullptr, nullptr, ws, m, n, rpc));
  if (int rc = check_launch("mmgl_colsum(partial)")) return rc;
  MMGL_CUDA(launch_pdl(col_final_kernel, dim3((unsigned)((n + 31) / 32)), dim3(256), 0, s, ws, chunks, n, scale, gate, out, accumulate));
  return check_launch("mmgl_colsum(final)");
}

extern "C" int mmgl_gate_grad(const void* dy, int64_t lddy, const void* a, int64_t lda, int64_t m, int64_t n,
                              const float* gate, float* out, int32_t accumulate, void* workspace,
                              size_t workspace_bytes, void* stream_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  MMGL_BIND(dy, "mmgl_gate_grad");
  MMGL_REQUIRE(dy && a && gate && out && m > 0 && n > 0, "mmgl_gate_grad: bad arguments");
  MMGL_REQUIRE(n % 8 == 0 && lddy % 8 == 0 && lda % 8 == 0 && aligned16(dy) && aligned16(a),
               "mmgl_gate_grad: needs n, ld %% 8 == 0 and 16B alignment");
  const int blocks = (int)(m < 592 ? m : 592);
  MMGL_REQUIRE(workspace && workspace_bytes >= (size_t)blocks * sizeof(float), "mmgl_gate_grad: workspace too small");
  float* ws = reinterpret_cast<float*>(workspace);
  MMGL_CUDA(launch_pdl(dot_partial_kernel, dim3(blocks), dim3(256), 0, s, (const __nv_bfloat16*)dy, lddy, (const __nv_bfloat16*)a, lda, m, n, ws));
  if (int rc = check_launch("mmgl_gate_grad(partial)")) return rc;
  MMGL_CUDA(launch_pdl(gate_final_kernel, dim3(1), dim3(256), 0, s, ws, blocks, gate, out, accumulate));
  return check_launch("mmgl_gate_grad(final)");
}

extern "C" int mmgl_ce_fwd(const void* logits, int64_t ld, const int64_t* labels, int64_t rows, int64_t vocab,
                           int64_t ignore_index, float* lse, float* row_loss, float* loss, float* count, void* stream_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  MMGL_BIND(logits, "mmgl_ce_fwd");
  MMGL_REQUIRE(labels && lse && row_loss && loss && count && rows > 0 && vocab > 0 && ld >= vocab, "mmgl_ce_fwd: bad arguments");
  MMGL_REQUIRE(rows < (1ll << 31), "mmgl_ce_fwd: too many rows");
  const int vec = aligned16(logits) && ld % 8 == 0;
  ce_fwd_kernel<<<(unsigned)rows, 256, 0, s>>>((const __nv_bfloat16*)logits, ld, labels, vocab, ignore_index, lse, row_loss, vec);
  if (int rc = check_launch("mmgl_ce_fwd")) return rc;
  ce_final_kernel<<<1, 1024, 0, s>>>(row_loss, labels, rows, vocab, ignore_index, loss, count);
  return check_launch("mmgl_ce_fwd(final)");
}

extern "C" int mmgl_ce_bwd(const void* logits, int64_t ld, const int64_t* labels, const float* lse, const float* dloss,
                           const float* count, void* dlogits, int64_t ldd, int64_t rows, int64_t vocab,
                           int64_t ignore_index, void* stream_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  MMGL_BIND(logits, "mmgl_ce_bwd");
  MMGL_REQUIRE(labels && lse && dloss && count && dlogits && rows > 0 && vocab > 0, "mmgl_ce_bwd: bad arguments");
  MMGL_REQUIRE(rows < (1ll << 31) && (vocab + 2047) / 2048 < 65536, "mmgl_ce_bwd: problem too large for the grid");
  const int vec = aligned16(logits) && aligned16(dlogits) && ld % 8 == 0 && ldd % 8 == 0;
  dim3 g((unsigned)rows, (unsigned)((vocab + 2047) / 2048));
  ce_bwd_kernel<<<g, 256, 0, s>>>((const __nv_bfloat16*)logits, ld, labels, lse, dloss, count, (__nv_bfloat16*)dlogits, ldd,
                                  vocab, ignore_index, vec);
  return check_launch("mmgl_ce_bwd");
}

extern "C" int mmgl_dropout_apply(const void* x, int64_t ldx, void* out, int64_t ldo, int64_t m, int64_t n, float p,
                                  uint64_t seed, void* stream_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  MMGL_BIND(x, "mmgl_dropout_apply");
  MMGL_REQUIRE(x && out && m > 0 && n > 0, "mmgl_dropout_apply: bad arguments");
  MMGL_REQUIRE(p >= 0.f && p < 1.f, "mmgl_dropout_apply: p must be in [0,1)");
  const uint32_t thresh = (uint32_t)(p * 65536.f + 0.5f);
  const float scale = thresh ? 65536.f / (65536.f - (float)thresh) : 1.f;
  const int vec = aligned16(x) && aligned16(out) && ldx % 8 == 0 && ldo % 8 == 0;
  const int64_t total = m * ((n + 7) / 8);
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  MMGL_CUDA(launch_pdl(dropout_apply_kernel, dim3((unsigned)blocks), dim3(256), 0, s, (const __nv_bfloat16*)x, ldx, (__nv_bfloat16*)out,
                       ldo, m, n, thresh, scale, seed, vec));
  return check_launch("mmgl_dropout_apply");
}

extern "C" int mmgl_bank_pack_fwd(const mmgl_bank_args* a, void* stream_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  MMGL_REQUIRE(a && a->bank && a->mask && a->batch > 0, "mmgl_bank_pack_fwd: bad arguments");
  MMGL_BIND(a->bank, "mmgl_bank_pack_fwd");
  MMGL_REQUIRE(a->n_text >= 0 && a->n_image >= 0 && a->n_text + a->n_image > 0, "mmgl_bank_pack_fwd: no neighbors");
  MMGL_REQUIRE(a->row_width % 8 == 0 && a->n_tok > 0 && a->n_tok <= 256 && a->row_width % a->n_tok == 0,
               "mmgl_bank_pack_fwd: row_width must be a multiple of 8 and of n_tok");
  MMGL_REQUIRE(a->n_text == 0 || (a->text_proj && a->text_pos_ids && a->text_locations), "mmgl_bank_pack_fwd: text inputs missing");
  MMGL_REQUIRE(a->n_image == 0 || (a->image_proj && a->image_pos_ids && a->image_locations), "mmgl_bank_pack_fwd: image inputs missing");
  MMGL_REQUIRE(a->lpe == nullptr || (a->lpe_weight && a->lpe_k > 0 && a->lpe_k <= 32), "mmgl_bank_pack_fwd: lpe_k must be in [1,32]");
  const int64_t n_src = a->n_text + a->n_image;
  MMGL_CUDA(cudaMemsetAsync(a->bank, 0, (size_t)a->batch * n_src * a->row_width * 2, s));
  MMGL_CUDA(cudaMemsetAsync(a->mask, 0, (size_t)a->batch * n_src * a->n_tok, s));
  dim3 g((unsigned)(a->batch * n_src), (unsigned)((a->row_width + 2047) / 2048));
  bank_pack_fwd_kernel<<<g, 256, 0, s>>>(*a);
  return check_launch("mmgl_bank_pack_fwd");
}

extern "C" int mmgl_bank_pack_bwd(const mmgl_bank_bwd_args* a, void* stream_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  MMGL_REQUIRE(a && a->d_bank && a->batch > 0 && a->row_width % 8 == 0, "mmgl_bank_pack_bwd: bad arguments");
  MMGL_BIND(a->d_bank, "mmgl_bank_pack_bwd");
  const int64_t n_src = a->n_text + a->n_image;
  const unsigned cchunks = (unsigned)((a->row_width + 2047) / 2048);
  const auto DB = (const __nv_bfloat16*)a->d_bank;
  if (a->d_text_proj || a->d_image_proj) {
    dim3 g((unsigned)(a->batch * n_src), cchunks);
    bank_pack_bwd_proj_kernel<<<g, 256, 0, s>>>(*a);
    if (int rc = check_launch("mmgl_bank_pack_bwd(proj)")) return rc;
  }
  if (a->d_text_pos_table && a->n_text > 0) {
    dim3 g((unsigned)a->text_pos_rows, cchunks);
    bank_pack_bwd_pos_kernel<<<g, 256, 0, s>>>(DB, a->text_pos_ids, a->text_locations, a->batch, a->n_text, n_src, a->row_width, a->d_text_pos_table);
    if (int rc = check_launch("mmgl_bank_pack_bwd(text pos)")) return rc;
  }
  if (a->d_image_pos_table && a->n_image > 0) {
    dim3 g((unsigned)a->image_pos_rows, cchunks);
    bank_pack_bwd_pos_kernel<<<g, 256, 0, s>>>(DB, a->image_pos_ids, a->image_locations, a->batch, a->n_image, n_src, a->row_width, a->d_image_pos_table);
    if (int rc = check_launch("mmgl_bank_pack_bwd(image pos)")) return rc;
  }
  if (a->lpe && (a->d_lpe_weight || a->d_lpe_bias)) {
    MMGL_REQUIRE(a->lpe_k > 0 && a->lpe_k <= 32, "mmgl_bank_pack_bwd: lpe_k must be in [1,32]");
    bank_pack_bwd_lpe_kernel<<<(unsigned)((a->row_width + 127) / 128), 128, 0, s>>>(DB, a->lpe, a->batch, n_src, a->row_width, (int)a->lpe_k, a->d_lpe_weight, a->d_lpe_bias);
    if (int rc = check_launch("mmgl_bank_pack_bwd(lpe)")) return rc;
  }
  return 0;
}

extern "C" int mmgl_gcn_concat_fwd(const void* x, const float* adj, void* out, int64_t batch, int64_t nodes,
                                   int64_t dim, int32_t prepend_root, void* stream_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  MMGL_BIND(x, "mmgl_gcn_concat_fwd");
  MMGL_REQUIRE(x && adj && out && batch > 0 && nodes > 1 && nodes <= 96 && dim > 0, "mmgl_gcn_concat_fwd: bad arguments (nodes <= 96)");
  const size_t smem = (size_t)(nodes * nodes + nodes * 256) * sizeof(float);
  MMGL_CUDA(cudaFuncSetAttribute(gcn_concat_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 g((unsigned)batch, (unsigned)((dim + 255) / 256));
  gcn_concat_fwd_kernel<<<g, 256, smem, s>>>((const __nv_bfloat16*)x, adj, (__nv_bfloat16*)out, (int)nodes, dim, prepend_root);
  return check_launch("mmgl_gcn_concat_fwd");
}

extern "C" int mmgl_gcn_combine_bwd(const void* dc, const float* adj, const void* relu_mask, void* dx, int64_t batch,
                                    int64_t nodes, int64_t dim, int32_t drop_root, void* stream_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  MMGL_BIND(dc, "mmgl_gcn_combine_bwd");
  MMGL_REQUIRE(dc && adj && dx && batch > 0 && nodes > 1 && nodes <= 96 && dim > 0, "mmgl_gcn_combine_bwd: bad arguments (nodes <= 96)");
  const size_t smem = (size_t)(nodes * nodes + nodes * 256) * sizeof(float);
  MMGL_CUDA(cudaFuncSetAttribute(gcn_combine_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 g((unsigned)batch, (unsigned)((dim + 255) / 256));
  gcn_combine_bwd_kernel<<<g, 256, smem, s>>>((const __nv_bfloat16*)dc, adj, (const __nv_bfloat16*)relu_mask, (__nv_bfloat16*)dx, (int)nodes, dim, drop_root);
  return check_launch("mmgl_gcn_combine_bwd");
}
