"""ctypes binding of libmmgl_b200.so (the C ABI in include/mmgl_b200.h).

PyTorch is plumbing here: it owns device memory and streams; every tensor is passed as a raw device
pointer (``data_ptr()``) together with sizes/leading dimensions, on ``torch.cuda.current_stream()``.
There is NO fallback: if the shared library is missing or a kernel fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libmmgl_b200.so")

_lib = None
_lock = threading.Lock()

c_i64, c_i32, c_f32, c_vp, c_sz = C.c_int64, C.c_int32, C.c_float, C.c_void_p, C.c_size_t


class GemmArgs(C.Structure):
    _fields_ = [
        ("a0", c_vp), ("b0", c_vp), ("k0", c_i64), ("lda0", c_i64), ("ldb0", c_i64),
        ("a1", c_vp), ("b1", c_vp), ("k1", c_i64), ("lda1", c_i64), ("ldb1", c_i64),
        ("a_mn_major", c_i32), ("b_mn_major", c_i32),
        ("m", c_i64), ("n", c_i64),
        ("d", c_vp), ("ldd", c_i64), ("out_fp32", c_i32), ("accumulate", c_i32),
        ("alpha", c_f32), ("relu", c_i32),
        ("bias", c_vp), ("gate", c_vp),
        ("residual", c_vp), ("ldres", c_i64),
        ("aux", c_vp), ("ldaux", c_i64),
        ("relu_mask", c_vp), ("ldmask", c_i64),
        ("force_block_n", c_i32), ("dropout_p", c_f32), ("dropout_seed", C.c_uint64),
        ("raster", c_i32), ("pair", c_i32),
        ("workspace", c_vp), ("workspace_bytes", c_i64), ("stream_k", c_i32), ("reserved", c_i32),
    ]


class BankArgs(C.Structure):
    _fields_ = [
        ("text_proj", c_vp), ("text_pos_table", c_vp), ("text_pos_ids", c_vp), ("text_locations", c_vp),
        ("image_proj", c_vp), ("image_pos_table", c_vp), ("image_pos_ids", c_vp), ("image_locations", c_vp),
        ("batch", c_i64), ("n_text", c_i64), ("n_image", c_i64), ("row_width", c_i64), ("n_tok", c_i64),
        ("lpe", c_vp), ("lpe_k", c_i64), ("lpe_weight", c_vp), ("lpe_bias", c_vp),
        ("bank", c_vp), ("mask", c_vp),
    ]


class BankBwdArgs(C.Structure):
    _fields_ = [
        ("d_bank", c_vp),
        ("text_pos_ids", c_vp), ("text_locations", c_vp), ("image_pos_ids", c_vp), ("image_locations", c_vp),
        ("batch", c_i64), ("n_text", c_i64), ("n_image", c_i64), ("row_width", c_i64),
        ("d_text_proj", c_vp), ("d_image_proj", c_vp),
        ("d_text_pos_table", c_vp), ("text_pos_rows", c_i64),
        ("d_image_pos_table", c_vp), ("image_pos_rows", c_i64),
        ("lpe", c_vp), ("lpe_k", c_i64), ("d_lpe_weight", c_vp), ("d_lpe_bias", c_vp),
    ]


class AttnArgs(C.Structure):
    """mirror of mmgl_attn_args (include/mmgl_b200.h)"""
    _fields_ = [
        ("q", c_vp), ("ldq", c_i64), ("k", c_vp), ("ldk", c_i64), ("v", c_vp), ("ldv", c_i64),
        ("key_mask", c_vp), ("rel_bias", c_vp),
        ("o", c_vp), ("ldo", c_i64), ("stats", c_vp),
        ("batch", c_i64), ("seq_q", c_i64), ("seq_k", c_i64), ("heads", c_i64), ("head_dim", c_i64),
        ("scale", c_f32), ("causal", c_i32), ("dropout_p", c_f32), ("reserved", c_i32), ("dropout_seed", C.c_uint64),
        ("cu_seqlens", c_vp), ("total_tokens", c_i64),
    ]


# name -> (restype, argtypes); also the list of symbols the header declares (checked by the CPU test-suite)
SIGNATURES = {
    "mmgl_version": (c_i32, []),
    "mmgl_last_error_string": (C.c_char_p, []),
    "mmgl_launch_count": (c_i64, []),
    "mmgl_gemm_bf16": (c_i32, [C.POINTER(GemmArgs), c_vp]),
    "mmgl_gemm_workspace_bytes": (c_sz, []),
    "mmgl_xattn_fwd": (c_i32, [c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_vp, c_vp, c_i64, c_vp,
                                 c_i64, c_i64, c_i64, c_i64, c_i64, c_vp]),
    "mmgl_xattn_bwd": (c_i32, [c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_vp, c_vp,
                                 c_vp, c_i64, c_vp, c_i64, c_vp, c_i64,
                                 c_i64, c_i64, c_i64, c_i64, c_i64, c_vp]),
    "mmgl_attn_fwd": (c_i32, [C.POINTER(AttnArgs), c_vp]),
    "mmgl_attn_bwd_workspace_bytes": (c_sz, [c_i64, c_i64, c_i64]),
    "mmgl_attn_bwd": (c_i32, [C.POINTER(AttnArgs), c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_vp, c_vp, c_sz, c_vp]),
    "mmgl_layernorm_fwd": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_f32, c_vp]),
    "mmgl_layernorm_bwd_workspace_bytes": (c_sz, [c_i64, c_i64]),
    "mmgl_layernorm_bwd": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_vp, c_sz,
                                     c_i64, c_i64, c_vp]),
    "mmgl_rmsnorm_fwd": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_f32, c_vp]),
    "mmgl_rmsnorm_bwd": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_vp, c_sz, c_i64, c_i64, c_vp]),
    "mmgl_reduce_workspace_bytes": (c_sz, [c_i64, c_i64]),
    "mmgl_colsum": (c_i32, [c_vp, c_i64, c_i64, c_i64, c_f32, c_vp, c_vp, c_i32, c_vp, c_sz, c_vp]),
    "mmgl_gate_grad": (c_i32, [c_vp, c_i64, c_vp, c_i64, c_i64, c_i64, c_vp, c_vp, c_i32, c_vp, c_sz, c_vp]),
    "mmgl_ce_fwd": (c_i32, [c_vp, c_i64, c_vp, c_i64, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "mmgl_ce_bwd": (c_i32, [c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, c_i64, c_vp]),
    "mmgl_dropout_apply": (c_i32, [c_vp, c_i64, c_vp, c_i64, c_i64, c_i64, c_f32, C.c_uint64, c_vp]),
    "mmgl_bank_pack_fwd": (c_i32, [C.POINTER(BankArgs), c_vp]),
    "mmgl_bank_pack_bwd": (c_i32, [C.POINTER(BankBwdArgs), c_vp]),
    "mmgl_gcn_concat_fwd": (c_i32, [c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, c_i32, c_vp]),
    "mmgl_gcn_combine_bwd": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, c_i32, c_vp]),
    "mmgl_rope_inplace": (c_i32, [c_vp, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_vp, c_i32, c_vp]),
    "mmgl_swiglu_fwd": (c_i32, [c_vp, c_i64, c_vp, c_i64, c_i64, c_i64, c_vp]),
    "mmgl_swiglu_bwd": (c_i32, [c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_i64, c_i64, c_vp]),
    "mmgl_adamw_step": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_f32, c_f32, c_f32, c_f32, c_f32, c_i64, c_f32, c_vp]),
}


def lib():
    """Load (once) and return the shared library.  Raises if it was not built -- no fallback."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise RuntimeError(
                        f"{LIB_PATH} not found: build it with `python -m mmgl_b200.build` "
                        "(mmgl_b200 has no CPU/PyTorch fallback path)")
                handle = C.CDLL(LIB_PATH)
                for name, (res, args) in SIGNATURES.items():
                    fn = getattr(handle, name)
                    fn.restype = res
                    fn.argtypes = args
                if handle.mmgl_version() != 1:
                    raise RuntimeError("libmmgl_b200.so ABI version mismatch")
                _lib = handle
    return _lib


def last_error() -> str:
    return lib().mmgl_last_error_string().decode()


def launch_count() -> int:
    return int(lib().mmgl_launch_count())


# ------------------------------------------------------------------------------------------- per-launch timing
# bench.py brackets every launch of the hot kernels with CUDA events on the launching (= torch current) stream to
# report achieved FLOP/s / GB/s against the roofline.  Off by default: zero overhead on the normal path.
_prof = None
_prof_detail = False


def profile_begin(detail: bool = False):
    global _prof, _prof_detail
    _prof = []
    _prof_detail = detail


def profile_end():
    """-> {kernel: {"launches", "ms", "work"}}; work = algorithmic FLOPs (gemm) or bytes (xattn_*), see DESIGN.md."""
    global _prof
    rec, _prof = _prof or [], None
    torch.cuda.synchronize()
    out = {}
    for name, work, e0, e1 in rec:
        d = out.setdefault(name, {"launches": 0, "ms": 0.0, "work": 0.0})
        d["launches"] += 1
        d["ms"] += e0.elapsed_time(e1)
        d["work"] += work
    return out


class _Timed:
    __slots__ = ("name", "work", "e0")

    def __init__(self, name, work):
        self.name, self.work = name, work

    def __enter__(self):
        if _prof is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *exc):
        if _prof is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            _prof.append((self.name, self.work, self.e0, e1))
        return False


def _check(rc: int, what: str):
    if rc != 0:
        raise RuntimeError(f"{what} failed (rc={rc}): {last_error()}")


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _req_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("mmgl_b200 kernels need CUDA tensors (there is no CPU path)")


def _ld(t: torch.Tensor) -> int:
    """leading dimension (elements) of a 2-D row-major (possibly column-sliced) view"""
    assert t.dim() == 2 and (t.stride(1) == 1 or t.shape[1] == 1), f"need row-major 2-D view, got strides {t.stride()}"
    return t.stride(0) if t.shape[0] > 1 else max(t.stride(0), t.shape[1])


_ws_cache = {}


_DEFAULT_STREAM_K = int(os.environ.get("MMGL_GEMM_STREAM_K", "0"))   # tuning knob: 2 = cut K for the tail wave / skinny outputs


def _gemm_workspace(device) -> torch.Tensor:
    """Stream-K scratch (fp32 partial tiles + arrival counters), one buffer per (device, stream): launches on one
    stream are ordered, so consecutive GEMMs can share it; concurrent streams get their own."""
    key = (device.index if device.index is not None else torch.cuda.current_device(), _stream())
    ws = _ws_cache.get(key)
    if ws is None:
        ws = torch.zeros(int(lib().mmgl_gemm_workspace_bytes()), dtype=torch.uint8, device=device)
        _ws_cache[key] = ws
    return ws


# ------------------------------------------------------------------------------------------- GEMM
def gemm(a: torch.Tensor, b: torch.Tensor, out: torch.Tensor, *, a_t: bool = False, b_t: bool = False,
         a1: Optional[torch.Tensor] = None, b1: Optional[torch.Tensor] = None,
         alpha: float = 1.0, bias: Optional[torch.Tensor] = None, relu: bool = False,
         gate: Optional[torch.Tensor] = None, residual: Optional[torch.Tensor] = None,
         aux: Optional[torch.Tensor] = None, relu_mask: Optional[torch.Tensor] = None,
         accumulate: bool = False, block_n: int = 0, dropout_p: float = 0.0, dropout_seed: int = 0,
         raster: int = 0, pair: int = 0, stream_k: int = 0) -> torch.Tensor:
    """out[M,N] = epilogue(A @ B^T (+ A1 @ B1^T)).

    a:  [M,K] (a_t=False) or [K,M] (a_t=True: A is used transposed, i.e. stored M-contiguous)
    b:  [N,K] (b_t=False, the nn.Linear weight layout) or [K,N] (b_t=True)
    All operands bf16 2-D row-major views (column slices allowed); out bf16 or fp32.
    """
    _req_cuda(a, b, out)
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16
    m, k = (a.shape[1], a.shape[0]) if a_t else (a.shape[0], a.shape[1])
    n, kb = (b.shape[1], b.shape[0]) if b_t else (b.shape[0], b.shape[1])
    assert k == kb, f"K mismatch {k} vs {kb}"
    assert out.shape[0] == m and out.shape[1] == n, f"out shape {tuple(out.shape)} != {(m, n)}"
    g = GemmArgs()
    g.a0, g.b0, g.k0, g.lda0, g.ldb0 = _p(a), _p(b), k, _ld(a), _ld(b)
    if a1 is not None:
        assert b1 is not None and a1.dtype == torch.bfloat16 and b1.dtype == torch.bfloat16
        k1 = a1.shape[0] if a_t else a1.shape[1]
        g.a1, g.b1, g.k1, g.lda1, g.ldb1 = _p(a1), _p(b1), k1, _ld(a1), _ld(b1)
    g.a_mn_major, g.b_mn_major = int(a_t), int(b_t)
    g.m, g.n = m, n
    assert out.dtype in (torch.bfloat16, torch.float32)
    g.d, g.ldd, g.out_fp32, g.accumulate = _p(out), _ld(out), int(out.dtype == torch.float32), int(accumulate)
    g.alpha, g.relu = float(alpha), int(relu)   # relu: False/0 none, True/1 ReLU, 2 GELU(erf), 3 quick-GELU
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == n
    if gate is not None:
        assert gate.dtype == torch.float32 and gate.numel() == 1
    g.bias, g.gate = _p(bias), _p(gate)
    for name, t in (("residual", residual), ("aux", aux), ("relu_mask", relu_mask)):
        if t is not None:
            assert t.dtype == torch.bfloat16 and tuple(t.shape) == (m, n), f"{name} must be bf16 [{m},{n}]"
    g.residual, g.ldres = _p(residual), (_ld(residual) if residual is not None else 0)
    g.aux, g.ldaux = _p(aux), (_ld(aux) if aux is not None else 0)
    g.relu_mask, g.ldmask = _p(relu_mask), (_ld(relu_mask) if relu_mask is not None else 0)
    g.force_block_n = block_n
    g.raster = raster
    g.pair = pair
    ws = _gemm_workspace(out.device)
    g.workspace, g.workspace_bytes, g.stream_k = ws.data_ptr(), ws.numel(), stream_k or _DEFAULT_STREAM_K
    g.dropout_p, g.dropout_seed = float(dropout_p), int(dropout_seed) & 0xFFFFFFFFFFFFFFFF
    with _Timed(f"gemm_tcgen05 m={m} n={n} k={k + int(g.k1)} at={int(a_t)} bt={int(b_t)}" if _prof_detail else "gemm_tcgen05",
                2.0 * m * n * (k + int(g.k1))):
        _check(lib().mmgl_gemm_bf16(C.byref(g), _stream()), "mmgl_gemm_bf16")
    return out


# ------------------------------------------------------------------------------------------- attention
def xattn_fwd(q, k, v, mask, o, stats, batch, seq, nk, heads, head_dim):
    """q,o: [B*S, H] views; k,v: [B*Nk, H] views (may be halves of a fused K|V buffer); mask u8 [B,Nk]."""
    _req_cuda(q, k, v, mask, o, stats)
    assert mask.dtype == torch.uint8 and mask.is_contiguous() and stats.dtype == torch.float32
    h = heads * head_dim
    with _Timed("xattn_fwd", float(batch * ((2 * seq * h + 2 * nk * h) * 2 + nk))):
        _check(lib().mmgl_xattn_fwd(_p(q), _ld(q), _p(k), _ld(k), _p(v), _ld(v), _p(mask), _p(o), _ld(o), _p(stats),
                                    batch, seq, nk, heads, head_dim, _stream()), "mmgl_xattn_fwd")


def xattn_bwd(d_o, q, k, v, o, stats, mask, dq, dk, dv, batch, seq, nk, heads, head_dim):
    _req_cuda(d_o, q, k, v, o, stats, mask, dq, dk, dv)
    h = heads * head_dim
    with _Timed("xattn_bwd", float(batch * ((4 * seq * h + 2 * nk * h) * 2 + (seq * h + 2 * nk * h) * 2 + nk))):
        _check(lib().mmgl_xattn_bwd(_p(d_o), _ld(d_o), _p(q), _ld(q), _p(k), _ld(k), _p(v), _ld(v), _p(o), _ld(o),
                                    _p(stats), _p(mask), _p(dq), _ld(dq), _p(dk), _ld(dk), _p(dv), _ld(dv),
                                    batch, seq, nk, heads, head_dim, _stream()), "mmgl_xattn_bwd")


def _attn_args(q, k, v, key_mask, rel_bias, o, stats, batch, seq_q, seq_k, heads, head_dim, scale, causal, dropout_p,
               dropout_seed):
    a = AttnArgs()
    a.q, a.ldq, a.k, a.ldk, a.v, a.ldv = _p(q), _ld(q), _p(k), _ld(k), _p(v), _ld(v)
    a.key_mask, a.rel_bias = _p(key_mask), _p(rel_bias)
    a.o, a.ldo, a.stats = _p(o), _ld(o), _p(stats)
    a.batch, a.seq_q, a.seq_k, a.heads, a.head_dim = batch, seq_q, seq_k, heads, head_dim
    a.scale, a.causal, a.dropout_p, a.dropout_seed = float(scale), int(causal), float(dropout_p), int(dropout_seed)
    return a


def attn_fwd(q, k, v, key_mask, rel_bias, o, stats, batch, seq_q, seq_k, heads, head_dim, scale, causal,
             dropout_p=0.0, dropout_seed=0, cu_seqlens=None, total_tokens=0):
    """q, o: [B*seq_q, H] views; k, v: [B*seq_k, H] views; key_mask u8 [B,seq_k] or None; rel_bias fp32
    [heads, seq_q+seq_k-1] or None (bias of (row, key) = rel_bias[h][key - row + seq_q - 1])."""
    _req_cuda(q, k, v, key_mask, rel_bias, o)
    assert rel_bias is None or (rel_bias.dtype == torch.float32 and rel_bias.is_contiguous()
                                and tuple(rel_bias.shape) == (heads, seq_q + seq_k - 1))
    assert key_mask is None or (key_mask.dtype == torch.uint8 and key_mask.is_contiguous())
    h = heads * head_dim
    a = _attn_args(q, k, v, key_mask, rel_bias, o, stats, batch, seq_q, seq_k, heads, head_dim, scale, causal, dropout_p,
                   dropout_seed)
    if cu_seqlens is not None:      # packed variable-length batch (forward only)
        _req_cuda(cu_seqlens)
        assert cu_seqlens.dtype == torch.int32 and cu_seqlens.numel() == batch + 1 and cu_seqlens.is_contiguous()
        a.cu_seqlens, a.total_tokens = _p(cu_seqlens), int(total_tokens)
    nbytes = float((2 * total_tokens if cu_seqlens is not None else batch * (seq_q + seq_k)) * h * 2 * 2)
    with _Timed("sattn_fwd", nbytes):
        _check(lib().mmgl_attn_fwd(C.byref(a), _stream()), "mmgl_attn_fwd")


def attn_bwd(d_o, q, k, v, key_mask, rel_bias, o, stats, dq, dk, dv, batch, seq_q, seq_k, heads, head_dim, scale, causal,
             dropout_p=0.0, dropout_seed=0, d_rel_bias=None):
    """d_rel_bias: None, or a ZEROED fp32 [heads, seq_q+seq_k-1] tensor that receives the gradient of rel_bias."""
    _req_cuda(d_o, q, k, v, key_mask, rel_bias, o, stats, dq, dk, dv, d_rel_bias)
    assert d_rel_bias is None or (rel_bias is not None and d_rel_bias.dtype == torch.float32 and d_rel_bias.is_contiguous()
                                  and d_rel_bias.shape == rel_bias.shape)
    h = heads * head_dim
    a = _attn_args(q, k, v, key_mask, rel_bias, o, stats, batch, seq_q, seq_k, heads, head_dim, scale, causal, dropout_p,
                   dropout_seed)
    # fp32 scratch of the per-CTA dQ accumulation (L2-resident; never read by the caller)
    ws = _workspace(lib().mmgl_attn_bwd_workspace_bytes(batch, seq_q, heads), q.device)
    with _Timed("sattn_bwd", float(batch * (seq_q + seq_k) * h * 2 * 4)):
        _check(lib().mmgl_attn_bwd(C.byref(a), _p(d_o), _ld(d_o), _p(dq), _ld(dq), _p(dk), _ld(dk), _p(dv), _ld(dv),
                                   _p(d_rel_bias), _p(ws), ws.numel(), _stream()), "mmgl_attn_bwd")


# ------------------------------------------------------------------------------------------- layernorm
def layernorm_fwd(x, gamma, beta, y, mean, rstd, eps):
    _req_cuda(x, gamma, beta, y, mean, rstd)
    rows, hidden = x.shape
    assert x.is_contiguous() and y.is_contiguous() and gamma.dtype == torch.float32 and beta.dtype == torch.float32
    with _Timed("layernorm_fwd", float(rows * hidden * 4)):
        _check(lib().mmgl_layernorm_fwd(_p(x), _p(gamma), _p(beta), _p(y), _p(mean), _p(rstd), rows, hidden, eps,
                                        _stream()), "mmgl_layernorm_fwd")


def _workspace(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


def layernorm_bwd(dy, x, gamma, mean, rstd, d_res, dx, dgamma=None, dbeta=None, accumulate=False):
    _req_cuda(dy, x, gamma, mean, rstd, dx)
    rows, hidden = x.shape
    assert dy.is_contiguous() and x.is_contiguous() and dx.is_contiguous() and (d_res is None or d_res.is_contiguous())
    ws = None
    nbytes = 0
    if dgamma is not None or dbeta is not None:
        nbytes = lib().mmgl_layernorm_bwd_workspace_bytes(rows, hidden)
        ws = _workspace(nbytes, x.device)
    with _Timed("layernorm_bwd", float(rows * hidden * (8 if d_res is not None else 6))):
        _check(lib().mmgl_layernorm_bwd(_p(dy), _p(x), _p(gamma), _p(mean), _p(rstd), _p(d_res), _p(dx), _p(dgamma),
                                        _p(dbeta), int(accumulate), _p(ws), nbytes, rows, hidden, _stream()),
               "mmgl_layernorm_bwd")


def rmsnorm_fwd(x, gamma, y, rstd, eps):
    _req_cuda(x, gamma, y, rstd)
    rows, hidden = x.shape
    assert x.is_contiguous() and y.is_contiguous() and gamma.dtype == torch.float32
    with _Timed("rmsnorm_fwd", float(rows * hidden * 4)):
        _check(lib().mmgl_rmsnorm_fwd(_p(x), _p(gamma), _p(y), _p(rstd), rows, hidden, eps, _stream()), "mmgl_rmsnorm_fwd")


def rmsnorm_bwd(dy, x, gamma, rstd, d_res, dx, dgamma=None, accumulate=False):
    _req_cuda(dy, x, gamma, rstd, dx)
    rows, hidden = x.shape
    assert dy.is_contiguous() and x.is_contiguous() and dx.is_contiguous() and (d_res is None or d_res.is_contiguous())
    ws, nbytes = None, 0
    if dgamma is not None:
        nbytes = lib().mmgl_layernorm_bwd_workspace_bytes(rows, hidden)
        ws = _workspace(nbytes, x.device)
    with _Timed("rmsnorm_bwd", float(rows * hidden * 6)):
        _check(lib().mmgl_rmsnorm_bwd(_p(dy), _p(x), _p(gamma), _p(rstd), _p(d_res), _p(dx), _p(dgamma), int(accumulate),
                                      _p(ws), nbytes, rows, hidden, _stream()), "mmgl_rmsnorm_bwd")


# ------------------------------------------------------------------------------------------- reductions
def colsum(x, out, scale=1.0, gate=None, accumulate=False):
    _req_cuda(x, out)
    m, n = x.shape
    assert out.dtype == torch.float32 and out.numel() == n and x.dtype == torch.bfloat16
    nbytes = lib().mmgl_reduce_workspace_bytes(m, n)
    ws = _workspace(nbytes, x.device)
    with _Timed("colsum", float(m * n * 2)):
        _check(lib().mmgl_colsum(_p(x), _ld(x), m, n, float(scale), _p(gate), _p(out), int(accumulate), _p(ws), nbytes,
                                 _stream()), "mmgl_colsum")
    return out


def gate_grad(dy, a, gate, out, accumulate=False):
    _req_cuda(dy, a, gate, out)
    m, n = dy.shape
    nbytes = lib().mmgl_reduce_workspace_bytes(m, n)
    ws = _workspace(nbytes, dy.device)
    with _Timed("gate_grad", float(m * n * 4)):
        _check(lib().mmgl_gate_grad(_p(dy), _ld(dy), _p(a), _ld(a), m, n, _p(gate), _p(out), int(accumulate), _p(ws),
                                    nbytes, _stream()), "mmgl_gate_grad")
    return out


def ce_fwd(logits, labels, lse, row_loss, loss, count, ignore_index=-100):
    _req_cuda(logits, labels, lse, row_loss, loss, count)
    rows, vocab = logits.shape
    assert logits.dtype == torch.bfloat16 and labels.dtype == torch.int64 and labels.is_contiguous() and labels.numel() == rows
    with _Timed("ce_fwd", float(rows * vocab * 2)):
        _check(lib().mmgl_ce_fwd(_p(logits), _ld(logits), _p(labels), rows, vocab, ignore_index, _p(lse), _p(row_loss),
                                 _p(loss), _p(count), _stream()), "mmgl_ce_fwd")


def ce_bwd(logits, labels, lse, dloss, count, dlogits, ignore_index=-100):
    _req_cuda(logits, labels, lse, dloss, count, dlogits)
    rows, vocab = logits.shape
    assert dloss.dtype == torch.float32 and dloss.numel() == 1
    _check(lib().mmgl_ce_bwd(_p(logits), _ld(logits), _p(labels), _p(lse), _p(dloss), _p(count), _p(dlogits), _ld(dlogits),
                             rows, vocab, ignore_index, _stream()), "mmgl_ce_bwd")


def dropout_apply(x, out, p, seed):
    """out = keep ? x / (1-p) : 0 with the GEMM epilogue's mask for (seed, p); x, out bf16 2-D views."""
    _req_cuda(x, out)
    m, n = x.shape
    with _Timed("dropout_apply", float(m * n * 4)):
        _check(lib().mmgl_dropout_apply(_p(x), _ld(x), _p(out), _ld(out), m, n, float(p), int(seed) & 0xFFFFFFFFFFFFFFFF,
                                        _stream()), "mmgl_dropout_apply")
    return out


# ------------------------------------------------------------------------------------------- bank / gcn
def bank_pack_fwd(args: BankArgs):
    _check(lib().mmgl_bank_pack_fwd(C.byref(args), _stream()), "mmgl_bank_pack_fwd")


def bank_pack_bwd(args: BankBwdArgs):
    _check(lib().mmgl_bank_pack_bwd(C.byref(args), _stream()), "mmgl_bank_pack_bwd")


def gcn_concat_fwd(x, adj, out, batch, nodes, dim, prepend_root):
    _req_cuda(x, adj, out)
    assert adj.dtype == torch.float32 and adj.is_contiguous() and x.is_contiguous() and out.is_contiguous()
    _check(lib().mmgl_gcn_concat_fwd(_p(x), _p(adj), _p(out), batch, nodes, dim, int(prepend_root), _stream()),
           "mmgl_gcn_concat_fwd")


def gcn_combine_bwd(dc, adj, relu_mask, dx, batch, nodes, dim, drop_root):
    _req_cuda(dc, adj, dx)
    _check(lib().mmgl_gcn_combine_bwd(_p(dc), _p(adj), _p(relu_mask), _p(dx), batch, nodes, dim, int(drop_root),
                                      _stream()), "mmgl_gcn_combine_bwd")


# ------------------------------------------------------------------------------------------- llama pieces
def rope_inplace(x, rows, seq, heads, head_dim, sections, cos_sin, inverse=False):
    """x: [rows, >= sections*heads*head_dim] bf16 row-major view, rotated in place; cos_sin fp32 [seq, head_dim/2, 2]."""
    _req_cuda(x, cos_sin)
    assert x.dtype == torch.bfloat16 and cos_sin.dtype == torch.float32 and cos_sin.is_contiguous()
    assert tuple(cos_sin.shape) == (seq, head_dim // 2, 2)
    with _Timed("rope", float(rows * sections * heads * head_dim * 4)):
        _check(lib().mmgl_rope_inplace(_p(x), _ld(x), rows, seq, heads, head_dim, sections, _p(cos_sin), int(inverse),
                                       _stream()), "mmgl_rope_inplace")


def swiglu_fwd(gu, h):
    _req_cuda(gu, h)
    m, f = h.shape
    assert gu.shape[1] == 2 * f and gu.dtype == torch.bfloat16 and h.dtype == torch.bfloat16
    with _Timed("swiglu_fwd", float(m * f * 6)):
        _check(lib().mmgl_swiglu_fwd(_p(gu), _ld(gu), _p(h), _ld(h), m, f, _stream()), "mmgl_swiglu_fwd")


def swiglu_bwd(gu, dh, dgu):
    _req_cuda(gu, dh, dgu)
    m, f = dh.shape
    with _Timed("swiglu_bwd", float(m * f * 10)):
        _check(lib().mmgl_swiglu_bwd(_p(gu), _ld(gu), _p(dh), _ld(dh), _p(dgu), _ld(dgu), m, f, _stream()), "mmgl_swiglu_bwd")


def adamw_step(param, grad, exp_avg, exp_avg_sq, shadow, lr, beta1, beta2, eps, weight_decay, step, grad_scale=1.0):
    """In-place AdamW update of one fp32 tensor; ``shadow`` (bf16, same shape, or None) receives the updated values."""
    _req_cuda(param, grad, exp_avg, exp_avg_sq)
    n = param.numel()
    for t in (param, grad, exp_avg, exp_avg_sq):
        if t.dtype != torch.float32 or not t.is_contiguous() or t.numel() != n:
            raise ValueError("adamw_step: param, grad, exp_avg, exp_avg_sq must be contiguous fp32 tensors of one size")
    if shadow is not None and (shadow.dtype != torch.bfloat16 or not shadow.is_contiguous() or shadow.numel() != n):
        raise ValueError("adamw_step: shadow must be a contiguous bf16 tensor of the parameter's size")
    with _Timed("adamw_step", float(n * 30)):
        _check(lib().mmgl_adamw_step(_p(param), _p(grad), _p(exp_avg), _p(exp_avg_sq), _p(shadow), n, float(lr), float(beta1),
                                     float(beta2), float(eps), float(weight_decay), int(step), float(grad_scale), _stream()),
               "mmgl_adamw_step")
