/* mmgl_b200 -- C ABI of the B200 (sm_100a) kernels behind MMGL's neighbor-fusion training step.
 *
 * The reference (minjiyoon/MMGL) is pure Python: it has no FFI of its own.  Each entry point below
 * replaces a run of PyTorch/HF library calls inside the reference modules; the call sites are cited
 * as  model/<file>:<lines>  relative to the reference root.  The Python binding that a maintainer of
 * the reference would add is a ctypes stub; mmgl_b200/_capi.py is that stub.
 *
 * Conventions
 *   - All pointers are DEVICE pointers owned by the caller (PyTorch allocations); the library never
 *     allocates, frees or retains device memory.  Activations/weights are bf16 unless noted, row-major.
 *   - Every call is asynchronous on `stream` (a cudaStream_t passed as void*); no implicit sync.
 *   - Return 0 on success, non-zero on error; mmgl_last_error_string() describes the last error of the
 *     calling thread.  Nothing throws or exits.
 *   - Entry points are re-entrant (forward thread + autograd backward thread).
 */
#ifndef MMGL_B200_H
#define MMGL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMGL_ABI_VERSION 1

int mmgl_version(void);
const char* mmgl_last_error_string(void);
/* Number of kernels this library launched since load (all threads); bench.py reports it as gpu_launches. */
int64_t mmgl_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * tcgen05 GEMM with fused epilogue (TMA-fed, TMEM accumulators, persistent, warp-specialised).
 *
 *   acc[m,n] = sum_k A0[m,k]*B0[n,k]  (+ sum_k A1[m,k]*B1[n,k] when k1 > 0)       fp32 accumulate
 *   v = alpha * (acc + bias[n])                  bias optional (fp32)
 *   v = act(v)                                   relu = 1: max(v,0); 2: GELU (erf); 3: quick-GELU v*sigmoid(1.702v)
 *   v = relu_mask[m,n] > 0 ? v : 0               if relu_mask (bf16, ld = ldmask)      (ReLU backward)
 *   v = keep(m,n) ? v / (1 - p) : 0              if dropout_p > 0 (counter-based mask from dropout_seed, see
 *                                                mmgl_dropout_apply; p is quantised to 1/65536)
 *   aux[m,n] = v                                 if aux (bf16, ld = ldaux)             (pre-gate value)
 *   v = tanh(*gate) * v                          if gate (device fp32 scalar)
 *   v += residual[m,n]                           if residual (bf16, ld = ldres)
 *   v += D[m,n]                                  if accumulate (D read in its own dtype)
 *   D[m,n] = v                                   bf16 (out_fp32 = 0) or fp32 (out_fp32 = 1), ld = ldd
 *
 * Operand storage: a_mn_major = 0 -> A is [M rows][K cols] (K contiguous, the nn.Linear activation);
 *                  a_mn_major = 1 -> A is [K rows][M cols] (M contiguous; transposed use, e.g. wgrad).
 *                  b_mn_major likewise for B ([N][K] = nn.Linear weight layout when 0; [K][N] when 1).
 * Requirements: bf16 operands, 16-byte aligned base pointers, leading dimensions multiples of 8
 * elements (a TMA constraint: the row pitch of a tensor map is a multiple of 16 bytes).  M, N, K themselves are
 * arbitrary (TMA zero-fills operand tails, output tails are predicated) -- but an operand stored CONTIGUOUSLY has
 * leading dimension = its contiguous extent, so in practice: K % 8 == 0 for K-major operands, M % 8 == 0 (N % 8 == 0)
 * for an MN-major A (B).  A caller with such a shape pads the leading dimension (e.g. a [rows][12] matrix stored with
 * pitch 16); the call returns code 2 with a message otherwise.  Inside this package the only such contraction, the
 * Laplacian-PE projection with K = 12 / 28 (model/modelling_self_attention.py:313), does not go through the GEMM: it
 * is fused into mmgl_bank_pack_fwd.
 *
 * Replaces: nn.Linear / torch.bmm call sites  model/modelling_cross_attention.py:194,198-199,273,352,355
 * (and their autograd backward), :997,:1020 (neighbor projections), model/graph.py:24,29,
 * peft LoRA (model/modelling_self_attention.py:80-87) via the second operand pair.
 */
typedef struct mmgl_gemm_args {
  const void* a0; const void* b0; int64_t k0; int64_t lda0; int64_t ldb0;
  const void* a1; const void* b1; int64_t k1; int64_t lda1; int64_t ldb1;
  int32_t a_mn_major; int32_t b_mn_major;
  int64_t m; int64_t n;
  void* d; int64_t ldd; int32_t out_fp32; int32_t accumulate;
  float alpha; int32_t relu;
  const float* bias;
  const float* gate;
  const void* residual; int64_t ldres;
  void* aux; int64_t ldaux;
  const void* relu_mask; int64_t ldmask;
  int32_t force_block_n; /* 0 = heuristic; 64/128/192/256 to force (tests, tuning) */
  float dropout_p;       /* 0 = no dropout */
  uint64_t dropout_seed;
  int32_t raster;        /* tile order: 0 = heuristic, 1 = M-fastest, 2 = N-fastest (tests, tuning) */
  int32_t pair;          /* CTA-pair (cta_group::2, 256-row tiles) kernel: 0 = heuristic, 1 = never, 2 = always */
  /* Optional scratch for the stream-K tail of the CTA-pair kernel (the last partial wave of tiles is cut into K-slices
   * whose fp32 partial tiles are parked here and added by the slice-0 owner).  Caller-owned, >= mmgl_gemm_workspace_bytes(),
   * not shared between streams that run concurrently; NULL disables stream-K.  stream_k: 2 = on, otherwise off.  With 2,
   * skinny outputs (at most SMs/8 128 x 128 tiles over K >= 2048, e.g. the rank-64 LoRA weight gradients) are cut into
   * up to 12 K-slices per tile that ALL park their partials; a small reduce kernel sums them in slice order and runs the
   * epilogue.  Results stay run-to-run deterministic.  Measured on B200 in round 1: tail wave of big problems -- the
   * owner's serial fix-up reads cost about what the slices save at K <= 8192; skinny outputs -- 30 -> 19.5 us alone and
   * L2-cold, no measurable change inside the cfg3 step.  So data-parallel stays the default. */
  void* workspace; int64_t workspace_bytes; int32_t stream_k; int32_t reserved;
} mmgl_gemm_args;

int mmgl_gemm_bf16(const mmgl_gemm_args* args, void* stream);
size_t mmgl_gemm_workspace_bytes(void);   /* enough for any problem: flags + (SMs/2) fp32 256x256 partial tiles */

/* ------------------------------------------------------------------------------------------------
 * Fused cross-attention core:  O = softmax(max(Q K^T + mask, FLT_MIN_FINITE)) V   per (sample, head).
 * Q [B,S,nh*d] already scaled by d^-1/2 (done in the q_proj epilogue); K,V [B,Nk,nh*d] (may be the two
 * halves of one fused K|V projection: pass ldk = ldv = 2*nh*d); mask [B,Nk] bytes (1 = attend).
 * Head split/merge, mask expansion, clamp, fp32 softmax and both contractions (tcgen05.mma, TMEM accumulators,
 * TMA-staged Q / K / V tiles) are fused; no [B,nh,S,Nk] tensor ever reaches HBM.  stats [B,nh,S,2] fp32 = (row max, 1 / row sum) is saved for backward
 * (log-sum-exp = stats[0] - log(stats[1])).  d in {64,128}; Nk <= 256 forward (128 when d = 128).
 * Replaces model/modelling_cross_attention.py:176-177,206-271 and :68-79 (_expand_mask).
 */
int mmgl_xattn_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                   const uint8_t* mask, void* o, int64_t ldo, float* stats,
                   int64_t batch, int64_t seq, int64_t nk, int64_t heads, int64_t head_dim, void* stream);

/* Backward of the above (five tcgen05 contractions per 128-query tile, dK / dV accumulated in TMEM across tiles):
 * dQ [B,S,nh*d], dK/dV [B,Nk,nh*d] (lddk = lddv = 2*nh*d for a fused d(K|V)).  Nk <= 128.
 * Reference artifact reproduced on purpose: the clamp torch.max(S + mask, finfo.min) (:225-228) ties on every masked
 * entry and torch splits a tie's gradient in half, so dS of a masked entry is 0.5 * P (dP - delta).  It is only
 * non-zero for a sample whose neighbors are ALL masked (P uniform); everywhere else P = 0 on masked entries. */
int mmgl_xattn_bwd(const void* d_o, int64_t lddo, const void* q, int64_t ldq, const void* k, int64_t ldk,
                   const void* v, int64_t ldv, const void* o, int64_t ldo, const float* stats, const uint8_t* mask,
                   void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv,
                   int64_t batch, int64_t seq, int64_t nk, int64_t heads, int64_t head_dim, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Self-attention core of the decoder / encoder layers and of the language model on the concat path
 * (SURVEY 8f row f1, row a7):
 *   P = softmax(max(scale * Q K^T + rel_bias[h][key - row + seq_q - 1] + causal + key padding, finfo.min))
 *   O = (keep / (1 - dropout_p) . P) V     per (sample, head)
 * q, o [B,seq_q,nh*d]; k, v [B,seq_k,nh*d] bf16 (views with leading dims: the three thirds of one fused
 * [B*S, 3*nh*d] QKV projection work in place); key_mask [B,seq_k] bytes (1 = real token) or NULL; causal != 0 masks
 * keys > query + (seq_k - seq_q) (needs seq_q <= seq_k: the queries sit at the END of the key range, as with the
 * virtual-token K/V that peft prefix tuning prepends, model/modelling_self_attention.py:88-92); rel_bias fp32 [heads, seq_q + seq_k - 1] or NULL; stats [B,nh,seq_q,2] =
 * (row max of the masked scores, 1 / row sum) saved for backward.  head_dim in {64,128}; seq_k <= 8192.
 * keep = the counter-based mask of mmgl_dropout_apply over the [batch*heads*seq_q, seq_k] probability matrix,
 * row = (b*heads + h)*seq_q + i.
 *
 * 128 x 128 score blocks on tcgen05 with TMEM accumulators, TMA-staged tiles, two-pass fp32 softmax; only blocks at
 * or below the diagonal are visited when causal.  Forward: two query tiles per CTA in ping-pong (16 softmax warps, two
 * MMA-issuing warps, one TMA warp; pass 2 on 64-key half blocks with multi-buffered S).  Backward = a persistent dQ kernel
 * (items = query tiles) + a persistent dK/dV kernel (items = key blocks), both recomputing P from stats, four threads per
 * score row; the gradient of the bias is optional (its table is frozen under LoRA).
 *
 * Replaces MPTAttention's self branch model/modelling_cross_attention.py:201-275 with the mask of :455-476, and the
 * attention of the HF T5 / OPT language model that model/modelling_self_attention.py:332 runs (HF
 * models/t5/modeling_t5.py T5Attention: scores NOT scaled by d^-1/2, position_bias[h,i,j] a function of j - i,
 * nn.functional.dropout on the softmax output; decoder cross-attention with seq_q != seq_k, no bias).
 * A query whose keys are ALL masked (never the case with the reference's right-padded batches) attends uniformly over
 * the existing keys of the visited blocks rather than over all S keys. */
typedef struct mmgl_attn_args {
  const void* q; int64_t ldq; const void* k; int64_t ldk; const void* v; int64_t ldv;
  const uint8_t* key_mask; const float* rel_bias;
  void* o; int64_t ldo; float* stats;
  int64_t batch; int64_t seq_q; int64_t seq_k; int64_t heads; int64_t head_dim;
  float scale; int32_t causal; float dropout_p; int32_t reserved; uint64_t dropout_seed;
  /* Forward only (the frozen neighbor encoders, model/modelling_cross_attention.py:992): a packed variable-length batch.
   * cu_seqlens int32 [batch + 1] (device): sample b is rows [cu[b], cu[b+1]) of q / k / v / o, which hold total_tokens
   * rows; seq_q = seq_k = the longest sample; no key mask / bias / dropout; stats may be NULL.  NULL = uniform batch. */
  const int32_t* cu_seqlens; int64_t total_tokens;
} mmgl_attn_args;
int mmgl_attn_fwd(const mmgl_attn_args* args, void* stream);
/* o and stats in args are the forward outputs (read here).  One kernel computes dQ, dK and dV from a single
 * recomputation of the scores (csrc/sattn_bwd_sm100.cu): a persistent CTA owns a (sample, head), accumulates dK / dV of a
 * key block in TMEM and sums the dQ contributions of successive key blocks in `workspace`, a caller-owned fp32 scratch of
 * mmgl_attn_bwd_workspace_bytes() (16-byte aligned; one [ceil(seq_q / 128) * 128, 128] tile set per resident CTA; its
 * contents on entry and exit are irrelevant).  The accumulation order is fixed, so dQ / dK / dV are run-to-run
 * deterministic.  seq_q is limited by the shared memory that holds one rowsum(dO . O) per query row (about 8000). */
size_t mmgl_attn_bwd_workspace_bytes(int64_t batch, int64_t seq_q, int64_t heads);
/* d_rel_bias: NULL, or fp32 [heads, seq_q + seq_k - 1] that receives += the gradient of rel_bias (sum of dS over every
 * (sample, row, key) with the same key - row; the caller zeroes it).  Accumulated with fp32 atomics, so its low bits are
 * not run-to-run deterministic; everything else in the library is. */
int mmgl_attn_bwd(const mmgl_attn_args* args, const void* d_o, int64_t lddo, void* dq, int64_t lddq, void* dk,
                  int64_t lddk, void* dv, int64_t lddv, float* d_rel_bias, void* workspace, size_t workspace_bytes,
                  void* stream);



/* ------------------------------------------------------------------------------------------------
 * LayerNorm over the last dim (bf16 in/out, fp32 gamma/beta and statistics).
 * Replaces nn.LayerNorm at model/modelling_cross_attention.py:320,341,350,365.
 * y may be bf16; mean/rstd [rows] fp32 saved for backward.
 */
int mmgl_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                       int64_t rows, int64_t hidden, float eps, void* stream);
/* dx = (d_res ? d_res : 0) + LN'(dy);  dgamma/dbeta fp32 [hidden] (accumulate != 0 -> +=).
 * dgamma/dbeta may be NULL (frozen LayerNorm: only dx).  workspace >= mmgl_layernorm_bwd_workspace_bytes. */
size_t mmgl_layernorm_bwd_workspace_bytes(int64_t rows, int64_t hidden);
int mmgl_layernorm_bwd(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd,
                       const void* d_res, void* dx, float* dgamma, float* dbeta, int32_t accumulate,
                       void* workspace, size_t workspace_bytes, int64_t rows, int64_t hidden, void* stream);

/* T5-style RMSNorm (HF models/t5/modeling_t5.py T5LayerNorm, the norm of the T5 language model on the concat path,
 * model/modelling_self_attention.py:68): y = gamma * x * rsqrt(mean(x^2) + eps); no mean subtraction, no shift.
 * Same kernels as LayerNorm in their mean-free mode; rstd [rows] saved for backward; dgamma may be NULL (frozen).
 * workspace >= mmgl_layernorm_bwd_workspace_bytes. */
int mmgl_rmsnorm_fwd(const void* x, const float* gamma, void* y, float* rstd, int64_t rows, int64_t hidden, float eps,
                     void* stream);
int mmgl_rmsnorm_bwd(const void* dy, const void* x, const float* gamma, const float* rstd, const void* d_res, void* dx,
                     float* dgamma, int32_t accumulate, void* workspace, size_t workspace_bytes, int64_t rows,
                     int64_t hidden, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Reductions used by the backward pass.
 * colsum:   out[n] (+)= scale * tanh?(gate) * sum_m x[m,n]      (bias gradients; x bf16 [M,N] ld)
 * gate_grad: out[0] (+)= (1 - tanh(*gate)^2) * sum_{m,n} dy[m,n]*a[m,n]   (d loss / d gating scalar,
 *           model/modelling_cross_attention.py:335,359)
 */
size_t mmgl_reduce_workspace_bytes(int64_t m, int64_t n);
int mmgl_colsum(const void* x, int64_t ldx, int64_t m, int64_t n, float scale, const float* gate,
                float* out, int32_t accumulate, void* workspace, size_t workspace_bytes, void* stream);
int mmgl_gate_grad(const void* dy, int64_t lddy, const void* a, int64_t lda, int64_t m, int64_t n,
                   const float* gate, float* out, int32_t accumulate, void* workspace, size_t workspace_bytes,
                   void* stream);

/* ------------------------------------------------------------------------------------------------
 * Softmax cross-entropy over bf16 logits, fp32 math (nn.CrossEntropyLoss at model/modelling_cross_attention.py:828-836;
 * the caller expresses the "shift" by passing labels[b,s+1] for row (b,s) and ignore_index for the last position, so
 * the [B,S-1,V] slice copy of the reference is never made).
 *   lse[r] = logsumexp(logits[r,:]);  row_loss[r] = lse[r] - logits[r,labels[r]]  (0 if labels[r] == ignore_index or out
 *   of range);  count[0] = number of contributing rows;  loss[0] = sum(row_loss) / max(1, count)   (mean reduction).
 * Backward: dlogits[r,c] = (exp(logits[r,c] - lse[r]) - [c == labels[r]]) * dloss[0] / count[0]; 0 on ignored rows.
 * No fp32 copy of the logits is ever materialised (the reference's CE reads/writes [B,S,V] fp32 several times). */
int mmgl_ce_fwd(const void* logits, int64_t ld, const int64_t* labels, int64_t rows, int64_t vocab, int64_t ignore_index,
                float* lse, float* row_loss, float* loss, float* count, void* stream);
int mmgl_ce_bwd(const void* logits, int64_t ld, const int64_t* labels, const float* lse, const float* dloss,
                const float* count, void* dlogits, int64_t ldd, int64_t rows, int64_t vocab, int64_t ignore_index,
                void* stream);

/* out[m,n] = keep(m,n) ? x[m,n] / (1 - p) : 0 with the SAME counter-based mask the GEMM epilogue applies for
 * (seed, p): keep(m,n) iff 16 bits of splitmix64(seed ^ (2g + (n%8)/4)) >= round(p*65536), g = m*ceil(N/8) + n/8,
 * lane (n%8)%4.  Used by backward to re-apply the forward mask (nn.functional.dropout,
 * model/modelling_cross_attention.py:332, :356).  x, out bf16 [M,N] with leading dims ldx / ldo; in-place allowed. */
int mmgl_dropout_apply(const void* x, int64_t ldx, void* out, int64_t ldo, int64_t m, int64_t n, float p,
                       uint64_t seed, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Neighbor-bank packing (ragged per-sample interleave of text/image neighbor embeddings).
 *   bank[b, loc, :] = proj[b, j, :] + pos_table[pos_id[b,j], :]  (+ lpe[b, loc+1, :] . W_lpe^T + b_lpe)
 *   mask[b, loc*n_tok + t] = pos_id[b,j] > 0
 * with loc = locations[b,j]; text sources j in [0,T), image sources j in [0,I).  Row width = n_tok*H.
 * pos tables / lpe arguments may be NULL.  Slots not named by any location stay zero / masked.
 * Replaces model/modelling_cross_attention.py:999-1004,1022-1027,1080-1104 and
 * model/modelling_self_attention.py:284-315.
 */
typedef struct mmgl_bank_args {
  const void* text_proj;  const void* text_pos_table;  const int64_t* text_pos_ids;  const int64_t* text_locations;
  const void* image_proj; const void* image_pos_table; const int64_t* image_pos_ids; const int64_t* image_locations;
  int64_t batch; int64_t n_text; int64_t n_image; int64_t row_width; int64_t n_tok;
  const float* lpe; int64_t lpe_k; const void* lpe_weight; const float* lpe_bias; /* lpe [B, T+I+1, k] fp32; W bf16 [row_width, k] */
  void* bank; uint8_t* mask;
} mmgl_bank_args;
int mmgl_bank_pack_fwd(const mmgl_bank_args* args, void* stream);

/* Backward: d_text_proj/d_image_proj (bf16, gathered rows), d_pos tables (fp32 [rows, row_width], +=),
 * d_lpe_weight (fp32 [row_width,k], +=), d_lpe_bias (fp32, +=).  Any output may be NULL. */
typedef struct mmgl_bank_bwd_args {
  const void* d_bank;
  const int64_t* text_pos_ids;  const int64_t* text_locations;
  const int64_t* image_pos_ids; const int64_t* image_locations;
  int64_t batch; int64_t n_text; int64_t n_image; int64_t row_width;
  void* d_text_proj; void* d_image_proj;
  float* d_text_pos_table; int64_t text_pos_rows;
  float* d_image_pos_table; int64_t image_pos_rows;
  const float* lpe; int64_t lpe_k; float* d_lpe_weight; float* d_lpe_bias;
} mmgl_bank_bwd_args;
int mmgl_bank_pack_bwd(const mmgl_bank_bwd_args* args, void* stream);

/* ------------------------------------------------------------------------------------------------
 * GCN helpers (model/graph.py:17-31).  Nodes = 1 null root + N neighbors, adj [B,N+1,N+1] fp32.
 * concat_fwd:  out[b,i,:] = [ xr[b,i,:] , sum_j adj[b,i,j] * xr[b,j,:] ]   ([B*(N+1), 2*D] bf16)
 *              where xr = x with a zero root row prepended when prepend_root != 0 (x is then [B,N,D]),
 *              else xr = x ([B,N+1,D]).
 * combine_bwd: dx[b,j,:] = dc[b,j,:D] + sum_i adj[b,i,j] * dc[b,i,D:]   then optionally * (relu_mask>0);
 *              drop_root != 0 writes rows 1..N only (dx is [B,N,D]).
 */
int mmgl_gcn_concat_fwd(const void* x, const float* adj, void* out, int64_t batch, int64_t nodes, int64_t dim,
                        int32_t prepend_root, void* stream);
int mmgl_gcn_combine_bwd(const void* dc, const float* adj, const void* relu_mask, void* dx, int64_t batch,
                         int64_t nodes, int64_t dim, int32_t drop_root, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Elementwise pieces of frozen Llama decoder layers (BASELINE configs[4], SURVEY 8f row f4: Llama-2-7B with the gated
 * cross-attention blocks interleaved; the reference has no Llama wrapper, so these restate HF transformers'
 * models/llama/modeling_llama.py).
 *
 * mmgl_rope_inplace: rotary position embedding applied in place to the first `sections` (q, k) thirds of a fused
 *   Q|K|V buffer x [rows, ld] bf16 (heads interleaved: column = section * heads * head_dim + head * head_dim + i):
 *   (x[i], x[i + d/2]) -> (x[i] cos - x[i+d/2] sin, x[i+d/2] cos + x[i] sin), cos / sin = cos_sin[pos][i] (fp32 pairs,
 *   [seq, head_dim / 2, 2]), pos = row % seq.  inverse != 0 applies the transposed rotation (the backward).
 *   Replaces apply_rotary_pos_emb / rotate_half (HF modeling_llama.py:110-140).
 * mmgl_swiglu_fwd / _bwd: h = silu(g) * u over gu = [g | u] ([M, 2F] bf16, one GEMM over the row-concatenated gate_proj |
 *   up_proj weights); backward writes dgu = [dg | du].  Replaces LlamaMLP.forward's act_fn(gate_proj(x)) * up_proj(x)
 *   (HF modeling_llama.py:155-165) and its autograd backward.
 */
int mmgl_rope_inplace(void* x, int64_t ld, int64_t rows, int64_t seq, int64_t heads, int64_t head_dim, int64_t sections,
                      const float* cos_sin, int32_t inverse, void* stream);
int mmgl_swiglu_fwd(const void* gu, int64_t ldgu, void* h, int64_t ldh, int64_t m, int64_t f, void* stream);
int mmgl_swiglu_bwd(const void* gu, int64_t ldgu, const void* dh, int64_t lddh, void* dgu, int64_t lddgu, int64_t m, int64_t f,
                    void* stream);

/*
 * mmgl_adamw_step: one AdamW update of ONE fp32 tensor of n elements, in place, with the bf16 shadow of the updated values
 *   written in the same pass (shadow_bf16 may be NULL).  Arithmetic of torch.optim.AdamW (amsgrad = False, maximize = False):
 *     p *= 1 - lr * weight_decay;  m = b1 m + (1 - b1) g;  v = b2 v + (1 - b2) g^2;
 *     p -= lr / (1 - b1^step) * m / (sqrt(v) / sqrt(1 - b2^step) + eps),  g = grad * grad_scale.
 *   `step` is the 1-based count INCLUDING this update.  Replaces optimizer.step() of the torch.optim.AdamW the reference
 *   builds (language_modelling/run_generation.py:329-333, called at :486) and the fp32 -> bf16 weight conversion that
 *   model.bfloat16() (:306-307) stands for.  All pointers fp32 device memory except shadow_bf16.
 */
int mmgl_adamw_step(void* param, const void* grad, void* exp_avg, void* exp_avg_sq, void* shadow_bf16, int64_t n, float lr,
                    float beta1, float beta2, float eps, float weight_decay, int64_t step, float grad_scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MMGL_B200_H */
