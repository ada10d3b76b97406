// C entry points of the fused cross-attention core.  The kernels live in xattn_sm100.cu: forward and backward both run
// their contractions on tcgen05 tensor cores with TMEM accumulators and TMA-staged Q / K / V / dO tiles.
// Replaces model/modelling_cross_attention.py:176-177, 206-271, _expand_mask (:68-79) and their autograd backward.
#include <cfloat>
#include <initializer_list>
#include <cuda_bf16.h>

#include "../../include/mmgl_b200.h"
#include "common.cuh"

namespace mmgl {

int xattn_fwd_tc(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, const uint8_t* mask,
                 void* o, int64_t ldo, float* stats, int64_t batch, int64_t seq, int64_t nk, int64_t heads, int64_t d,
                 cudaStream_t stream);
int xattn_bwd_tc(const void* d_o, int64_t lddo, const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                 int64_t ldv, const void* o, int64_t ldo, const float* stats, const uint8_t* mask, void* dq, int64_t lddq,
                 void* dk, int64_t lddk, void* dv, int64_t lddv, int64_t batch, int64_t seq, int64_t nk, int64_t heads,
                 int64_t d, cudaStream_t stream);

static int check_xattn_args(const char* who, int64_t batch, int64_t seq, int64_t nk, int64_t heads, int64_t d,
                            std::initializer_list<int64_t> lds, std::initializer_list<const void*> ptrs) {
  MMGL_REQUIRE(batch > 0 && seq > 0 && nk > 0 && heads > 0, "%s: empty problem", who);
  MMGL_REQUIRE(d == 64 || d == 128, "%s: head_dim must be 64 or 128 (got %lld)", who, (long long)d);
  MMGL_REQUIRE(nk <= 256, "%s: Nk must be <= 256 (got %lld)", who, (long long)nk);
  MMGL_REQUIRE(batch < 65536 && heads < 65536, "%s: batch/heads too large for the grid", who);
  for (int64_t ld : lds) MMGL_REQUIRE(ld % 8 == 0, "%s: leading dims must be multiples of 8", who);
  for (const void* p : ptrs) MMGL_REQUIRE(p != nullptr && aligned16(p), "%s: pointers must be non-null, 16B aligned", who);
  return 0;
}

}  // namespace mmgl

using namespace mmgl;

extern "C" int mmgl_xattn_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                              const uint8_t* mask, void* o, int64_t ldo, float* stats, int64_t batch, int64_t seq,
                              int64_t nk, int64_t heads, int64_t d, void* stream_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  MMGL_BIND(q, "mmgl_xattn_fwd");
  if (int rc = check_xattn_args("mmgl_xattn_fwd", batch, seq, nk, heads, d, {ldq, ldk, ldv, ldo}, {q, k, v, o}))
    return rc;
  MMGL_REQUIRE(mask != nullptr && stats != nullptr, "mmgl_xattn_fwd: null mask/stats");
  MMGL_REQUIRE(d == 64 || nk <= 128, "mmgl_xattn_fwd: head_dim 128 supports Nk <= 128");
  return xattn_fwd_tc(q, ldq, k, ldk, v, ldv, mask, o, ldo, stats, batch, seq, nk, heads, d, s);
  return 0;
}

extern "C" int mmgl_xattn_bwd(const void* d_o, int64_t lddo, const void* q, int64_t ldq, const void* k, int64_t ldk,
                              const void* v, int64_t ldv, const void* o, int64_t ldo, const float* stats,
                              const uint8_t* mask, void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv,
                              int64_t lddv, int64_t batch, int64_t seq, int64_t nk, int64_t heads, int64_t d,
                              void* stream_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  MMGL_BIND(q, "mmgl_xattn_bwd");
  if (int rc = check_xattn_args("mmgl_xattn_bwd", batch, seq, nk, heads, d, {lddo, ldq, ldk, ldv, ldo, lddq, lddk, lddv},
                                {d_o, q, k, v, o, dq, dk, dv}))
    return rc;
  MMGL_REQUIRE(mask != nullptr && stats != nullptr, "mmgl_xattn_bwd: null mask/stats");
  return xattn_bwd_tc(d_o, lddo, q, ldq, k, ldk, v, ldv, o, ldo, stats, mask, dq, lddq, dk, lddk, dv, lddv, batch, seq, nk,
                      heads, d, s);
  return 0;
}
