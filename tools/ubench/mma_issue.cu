// Microbenchmark: cost of ISSUING tcgen05.mma from (a) `if (threadIdx.x == 0)` (what the round-1 kernels do) versus
// (b) a whole warp with `elect.sync` (uniform control flow), for M = 128, N in {64, 128}, SS mode, 8-way unrolled.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I mmgl_b200/csrc tools/ubench/mma_issue.cu -o tools/ubench/mma_issue.bin
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace mmgl;

template <int MODE, int N>
__global__ void __launch_bounds__(128, 1) k(int reps, long long* out) {
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
  const uint64_t da = make_smem_desc(smem_u32(smem), 16, 1024);
  const uint64_t db = make_smem_desc(smem_u32(smem + 32768), 16, 1024);
  constexpr uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
  const int warp = threadIdx.x >> 5;
  bool issuer;
  if (MODE == 0) issuer = threadIdx.x == 0;
  else issuer = (warp == 0);
  if (issuer) {
    uint32_t phase = 0;
    for (int w = 0; w < 2; ++w) {
      const long long t0 = clock64();
      if (MODE == 0) {
        for (int r = 0; r < reps; r += 8) {
#pragma unroll
          for (int u = 0; u < 8; ++u)
            umma_f16_ss(tmem + (u & 1) * 128, da + (uint64_t)(((u & 3) * 32) >> 4), db + (uint64_t)(((u & 3) * 32) >> 4), idesc, 1u);
        }
      } else if (MODE == 1) {
        for (int r = 0; r < reps; r += 8) {
          if (elect_one()) {
#pragma unroll
            for (int u = 0; u < 8; ++u)
              umma_f16_ss(tmem + (u & 1) * 128, da + (uint64_t)(((u & 3) * 32) >> 4), db + (uint64_t)(((u & 3) * 32) >> 4), idesc, 1u);
          }
          __syncwarp();
        }
      } else {
        const bool lead = elect_one();
        for (int r = 0; r < reps; r += 8) {
#pragma unroll
          for (int u = 0; u < 8; ++u)
            if (lead) umma_f16_ss(tmem + (u & 1) * 128, da + (uint64_t)(((u & 3) * 32) >> 4), db + (uint64_t)(((u & 3) * 32) >> 4), idesc, 1u);
        }
      }
      const long long ti = clock64();
      if (MODE == 0 || elect_one()) {
        umma_commit(&bar);
      }
      mbar_wait(&bar, phase & 1);
      ++phase;
      const long long t1 = clock64();
      if (w == 1 && blockIdx.x == 0 && (threadIdx.x & 31) == 0) { out[0] = t1 - t0; out[1] = ti - t0; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<512>(tmem);
}

template <int MODE, int N>
void run(long long* d, const char* what) {
  const int reps = 512;
  cudaFuncSetAttribute(k<MODE, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 1024);
  k<MODE, N><<<148, 128, 65536 + 1024>>>(reps, d);
  long long h[2] = {0, 0};
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  cudaError_t e = cudaDeviceSynchronize();
  printf("%-44s N=%3d: %6.1f cycles / MMA to completion, %6.1f to issue (ideal %d) %s\n", what, N, (double)h[0] / reps,
         (double)h[1] / reps, 128 * N / 256, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  run<0, 64>(d, "if (threadIdx.x == 0)");
  run<1, 64>(d, "warp: if (elect.sync) { 8 MMAs } syncwarp");
  run<2, 64>(d, "warp: lead = elect.sync once; if (lead) MMA");
  run<0, 128>(d, "if (threadIdx.x == 0)");
  run<1, 128>(d, "warp: if (elect.sync) { 8 MMAs } syncwarp");
  run<2, 128>(d, "warp: lead = elect.sync once; if (lead) MMA");
  return 0;
}
