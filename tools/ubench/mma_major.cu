// Microbenchmark: tcgen05.mma (kind::f16, bf16, cta_group::1, M = 128) issue-to-completion time per instruction for the four
// operand-major combinations and N in {64, 128, 256}, SS mode (both operands in shared memory, 128B swizzle).
// One CTA per SM (148 CTAs) so that shared-memory and tensor pipes see the same contention as a full-chip kernel.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I mmgl_b200/csrc tools/ubench/mma_major.cu -o /tmp/mma_major && /tmp/mma_major
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace mmgl;

__global__ void __launch_bounds__(128, 1) k(int m, int n, int a_mn, int b_mn, int reps, int a_tmem, int nacc, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc<512>(&tptr);
  for (int i = threadIdx.x; i < 65536 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tptr;
  if (threadIdx.x == 0) {
    // A tile at smem, B tile at smem + 32768; K-major: SBO 1024, LBO 16; MN-major: LBO 16384 (next 64-wide atom), SBO 1024
    const uint64_t da = make_smem_desc(smem_u32(smem), a_mn ? 16384 : 16, 1024);
    const uint64_t db = make_smem_desc(smem_u32(smem + 32768), b_mn ? 8192 : 16, 1024);
    const uint32_t idesc = make_idesc_bf16(m, n, a_mn, b_mn);
    uint32_t phase = 0;
    for (int w = 0; w < 2; ++w) {   // warm-up + timed
      const long long t0 = clock64();
      for (int r = 0; r < reps; r += 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const uint64_t ka = a_mn ? (uint64_t)((u * 2048) >> 4) : (uint64_t)((((u >> 2) * 16384) + (u & 3) * 32) >> 4);
          const uint64_t kb = b_mn ? (uint64_t)(((u & 3) * 2048) >> 4) : (uint64_t)(((u & 3) * 32) >> 4);
          const uint32_t dcol = tmem + (nacc == 1 ? 0 : (nacc == 2 ? (u & 1) * 64 : (u & 3) * 64));
          if (a_tmem) {
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
                         ::"r"(dcol), "r"(tmem + 256 + u * 8), "l"(db + kb), "r"(idesc), "r"(1u) : "memory");
          } else {
            umma_f16_ss(dcol, da + ka, db + kb, idesc, 1u);
          }
        }
      }
      const long long ti = clock64();
      umma_commit(&bar);
      mbar_wait(&bar, phase & 1);
      ++phase;
      const long long t1 = clock64();
      if (w == 1 && blockIdx.x == 0) out[1] = ti - t0;
      if (w == 1 && blockIdx.x == 0) out[0] = t1 - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<512>(tmem);
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 1024);
  const int reps = 256;
  for (int nacc : {1, 2, 4})
    for (int n : {64, 128, 256})
      for (int mode = 0; mode < 4; mode += 2) {     // 0: SS K/K, 2: TS B K-major
        const int m = 128, a_mn = 0, b_mn = 0, ts = mode >= 2;
        if ((nacc - 1) * 64 + n > 256) continue;
        k<<<148, 128, 65536 + 1024>>>(m, n, a_mn, b_mn, reps, ts, nacc, d);
        long long h[2] = {0, 0}; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        cudaError_t e = cudaDeviceSynchronize();
        printf("accumulators %d  M=%3d N=%3d K=16  A %s : %6.1f cycles / MMA to completion, %6.1f to issue (ideal %d)  %s\n", nacc, m, n,
               ts ? "TMEM   " : "K-major", (double)h[0] / reps, (double)h[1] / reps, 128 * n / 256, e == cudaSuccess ? "" : cudaGetErrorString(e));
      }
  return 0;
}
