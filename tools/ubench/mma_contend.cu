// Microbenchmark: rate of the attention forward's MMA groups (4 x S = Q K^T half block, SS, N = 64, K = 64; 4 x P V, N = 64,
// SS or TS) issued by one warp, alone and while 16 other warps stream tcgen05.ld / ex2 / tcgen05.st on the same SM the way
// the softmax warps do.  Answers: does TMEM / shared-memory traffic of the softmax warps slow the tensor pipe down?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I mmgl_b200/csrc tools/ubench/mma_contend.cu -o tools/ubench/mma_contend.bin
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace mmgl;

// LOAD: 0 = issuer alone, 1 = + tcgen05.ld loops, 2 = + ld + 32 ex2 per row chunk, 3 = + ld + ex2 + tcgen05.st
template <int TS, int LOAD>
__global__ void __launch_bounds__(576, 1) k(int reps, long long* out, float* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  __shared__ volatile int stop;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); stop = 0; }
  if (threadIdx.x < 32) tmem_alloc<512>(&tptr);
  for (int i = threadIdx.x; i < 98304 / 4; i += 576) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, tptr, 0);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  if (warp == 16) {
    const uint64_t dq = make_smem_desc(smem_u32(smem), 16, 1024);
    const uint64_t dk = make_smem_desc(smem_u32(smem + 16384), 16, 1024);
    const uint64_t dp = make_smem_desc(smem_u32(smem + 32768), 16, 1024);
    const uint64_t dv = make_smem_desc(smem_u32(smem + 65536), 16384, 1024);
    constexpr uint32_t idesc_s = make_idesc_bf16(128, 64, 0, 0), idesc_o = make_idesc_bf16(128, 64, 0, 1);
    uint32_t phase = 0;
    for (int w = 0; w < 2; ++w) {
      const long long t0 = clock64();
      for (int r = 0; r < reps; ++r) {
        if (elect_one()) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (TS) umma_f16_ts(tmem + 448, tmem + (r % 3) * 64 + (u >> 1) * 32 + (u & 1) * 8, dv + (uint64_t)((u * 2048) >> 4), idesc_o, 1u);
            else umma_f16_ss(tmem + 448, dp + (uint64_t)((u * 32) >> 4), dv + (uint64_t)((u * 2048) >> 4), idesc_o, 1u);
          }
#pragma unroll
          for (int u = 0; u < 4; ++u)
            umma_f16_ss(tmem + (r % 3) * 64, dq + (uint64_t)((u * 32) >> 4), dk + (uint64_t)((u * 32) >> 4), idesc_s, u != 0);
        }
        __syncwarp();
      }
      const long long ti = clock64();
      if (elect_one()) umma_commit(&bar);
      mbar_wait(&bar, phase & 1);
      ++phase;
      const long long t1 = clock64();
      if (w == 1 && blockIdx.x == 0 && (threadIdx.x & 31) == 0) { out[0] = t1 - t0; out[1] = ti - t0; }
    }
    stop = 1;
  } else if (warp < 16 && LOAD > 0) {
    const uint32_t taddr = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16) + 192 + (warp >> 2) * 32;   // columns 192..319: scratch
    float acc = 0.f;
    while (!stop) {
      uint32_t r[32];
      tmem_ld_32x32(taddr, r);
      tmem_ld_wait();
      if (LOAD >= 2) {
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          float x;
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(x) : "f"(__uint_as_float(r[e]) * 1e-3f));
          acc += x;
          r[e] = __float_as_uint(x);
        }
      } else {
#pragma unroll
        for (int e = 0; e < 32; ++e) acc += __uint_as_float(r[e]);
      }
      if (LOAD >= 3) {
        uint32_t pk[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) pk[e] = r[2 * e] ^ r[2 * e + 1];
        tmem_st_32x16(taddr, pk);
        tmem_st_wait();
      }
    }
    if (acc == 123.456f) sink[threadIdx.x] = acc;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<512>(tmem);
}

template <int TS, int LOAD>
void run(long long* d, float* sink, const char* what) {
  const int reps = 300;
  cudaFuncSetAttribute(k<TS, LOAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 98304 + 1024);
  k<TS, LOAD><<<148, 576, 98304 + 1024>>>(reps, d, sink);
  long long h[2] = {0, 0};
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  cudaError_t e = cudaDeviceSynchronize();
  printf("PV %s, %-40s: %7.1f cycles per (4 PV + 4 S) group to completion, %7.1f to issue %s\n", TS ? "TS" : "SS", what,
         (double)h[0] / reps, (double)h[1] / reps, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  float* sink; cudaMalloc(&sink, 4096);
  run<0, 0>(d, sink, "issuer alone");
  run<0, 1>(d, sink, "+ 16 warps tcgen05.ld");
  run<0, 2>(d, sink, "+ 16 warps ld + ex2");
  run<0, 3>(d, sink, "+ 16 warps ld + ex2 + tcgen05.st");
  run<1, 0>(d, sink, "issuer alone");
  run<1, 1>(d, sink, "+ 16 warps tcgen05.ld");
  run<1, 2>(d, sink, "+ 16 warps ld + ex2");
  run<1, 3>(d, sink, "+ 16 warps ld + ex2 + tcgen05.st");
  return 0;
}
