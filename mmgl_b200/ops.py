"""Autograd bindings of the sm_100a kernels (libmmgl_b200.so, through the ctypes layer in _capi.py).

Each ``torch.autograd.Function`` here replaces a run of PyTorch/HF library calls inside the reference modules
(file:line citations are relative to the reference root).  Forward AND backward call the C ABI; trainable
parameters receive ordinary ``.grad`` tensors, so ``DistributedDataParallel`` hooks fire exactly as in the
reference's loop (language_modelling/run_generation.py:319, :485).

Numerics: bf16 operands, fp32 accumulation / softmax / LayerNorm statistics.  Parameters may be stored in
fp32 ("master" weights -- a bf16 shadow is cached per parameter version) or in bf16 (the reference's
``model.bfloat16()`` mode, run_generation.py:306-307); gradients are returned in the parameter's dtype.

There is no CPU or eager fallback: every function raises if the tensors are not on a CUDA device.
"""
from __future__ import annotations

import contextlib
import os
import weakref
from typing import Optional

import torch

from . import _capi as K

BF16 = torch.bfloat16
F32 = torch.float32

# --------------------------------------------------------------------------------------------- helpers
_shadow = {}   # (id(param), dtype) -> (weakref(param), version, device, converted tensor)


_epoch = [0]   # bumped by invalidate_weight_cache(): part of every cache entry's validity stamp


def invalidate_weight_cache(trainable_only: bool = False) -> None:
    """Drop every cached bf16 / fp32 shadow and fused-row concatenation (``trainable_only``: only those of parameters
    with ``requires_grad`` -- the frozen layers' fused Wq|Wk|Wv rows then survive an optimizer step).

    Shadows are keyed on the parameter object, its ``_version`` counter, its storage pointer and device.  In-place
    updates made through autograd-visible ops (``copy_``, ``load_state_dict``, torch's foreach optimizers) bump ``_version``
    and refresh the shadow on the next use by themselves; ``torch.optim.AdamW(fused=True)`` and writes THROUGH ``param.data``
    (some EMA helpers, HF Adafactor, manual weight surgery) do not.  Every ``optimizer.step()`` is covered by the global
    post-step hook below; code that changes weights some other way must call this afterwards."""
    if not trainable_only:
        _epoch[0] += 1
        _shadow.clear()
        return
    for key, ent in list(_shadow.items()):
        refs = ent[0] if isinstance(ent[0], tuple) else (ent[0],)
        if any(r() is None or r().requires_grad for r in refs):
            _shadow.pop(key, None)


def _after_any_optimizer_step(optimizer, args, kwargs):
    """Global optimizer post-step hook: the cached bf16 shadows of TRAINABLE parameters are dropped after every
    ``optimizer.step()`` of any optimizer that does not maintain them itself (optim.FusedAdamW does).  The ``_version`` stamp
    alone is not enough: ``torch.optim.AdamW(fused=True)`` (``torch._fused_adamw_``) and optimizers that write through
    ``param.data`` update the weights WITHOUT bumping it -- found in round 2 when the loss of the benchmark loop stopped
    moving under torch's fused AdamW while it fell under ours."""
    if not getattr(optimizer, "keeps_shadows_current", False):
        invalidate_weight_cache(trainable_only=True)


import torch.optim.optimizer as _torch_optimizer  # noqa: E402

_torch_optimizer.register_optimizer_step_post_hook(_after_any_optimizer_step)


def _stamp(p: torch.Tensor):
    return (p._version, p.data_ptr(), p.device, _epoch[0])


def _converted(p: torch.Tensor, dtype) -> torch.Tensor:
    """`p` converted to `dtype`, cached per parameter object, `_version` and storage pointer (identity-keyed: tensors
    do not hash/compare by value, and the entry dies with the parameter).  See invalidate_weight_cache()."""
    key = (id(p), dtype)
    ent = _shadow.get(key)
    st = _stamp(p)
    if ent is not None and ent[0]() is p and ent[1] == st:
        return ent[2]
    conv = p.detach().to(dtype).contiguous()
    _shadow[key] = (weakref.ref(p, lambda _r, k=key: _shadow.pop(k, None)), st, conv)
    return conv


def register_shadow(p: torch.Tensor, shadow: torch.Tensor) -> None:
    """Install ``shadow`` as the current bf16 copy of ``p`` (optim.FusedAdamW writes it in the optimizer pass)."""
    key = (id(p), BF16)
    _shadow[key] = (weakref.ref(p, lambda _r, k=key: _shadow.pop(k, None)), _stamp(p), shadow)


def w16(p: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    """bf16 view of a weight: the tensor itself if already bf16, else a shadow cached per ``_version``."""
    if p is None:
        return None
    if p.dtype == BF16 and p.is_contiguous():
        return p.detach()
    return _converted(p, BF16)


def f32(p: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    """fp32 contiguous view of a small parameter (bias, LayerNorm affine, gate)."""
    if p is None:
        return None
    if p.dtype == F32 and p.is_contiguous():
        return p.detach()
    return _converted(p, F32)


_seed_counter = [0]
_MASK64 = (1 << 64) - 1


_NVTX = os.environ.get("MMGL_NVTX", "0") not in ("", "0")


@contextlib.contextmanager
def nvtx_range(name: str):
    """NVTX range around a phase of the step (neighbor encoders, bank, LM layers, loss) when MMGL_NVTX=1, so that ncu
    (--nvtx --nvtx-include) and Nsight timelines can be cut by phase; a no-op otherwise (no push / pop on the hot path)."""
    if not _NVTX:
        yield
        return
    torch.cuda.nvtx.range_push(name)
    try:
        yield
    finally:
        torch.cuda.nvtx.range_pop()


def next_dropout_seed() -> int:
    """A fresh 64-bit seed per dropout site and step, derived from torch's seed (so ``torch.manual_seed`` makes runs
    repeatable) and a process-local call counter.  No device sync, no RNG kernel: the mask itself is counter-based."""
    _seed_counter[0] += 1
    z = (torch.initial_seed() * 0x9E3779B97F4A7C15 + _seed_counter[0] * 0xD1B54A32D192ED03) & _MASK64
    z ^= z >> 32
    return (z * 0xBF58476D1CE4E5B9) & _MASK64


def peek_dropout_seeds(n: int):
    """The next ``n`` seeds next_dropout_seed() will hand out (tests use it to rebuild the masks)."""
    saved = _seed_counter[0]
    try:
        return [next_dropout_seed() for _ in range(n)]
    finally:
        _seed_counter[0] = saved


def _undrop(dy2, p, seed):
    """gradient w.r.t. the pre-dropout value: dy * keep / (1 - p) (same counter-based mask as forward)."""
    if p <= 0.0:
        return dy2
    out = torch.empty_like(dy2)
    K.dropout_apply(dy2, out, p, seed)
    return out


def fused_rows(params, dtype):
    """Row-concatenation of several parameters (e.g. Wq|Wk|Wv -> [3H, H]) converted to ``dtype``, cached until any of
    them changes.  Used for FROZEN projections only: one N = 3H GEMM instead of three N = H GEMMs."""
    key = (tuple(id(p) for p in params), dtype, "rows")
    ent = _shadow.get(key)
    vers = tuple(_stamp(p) for p in params)
    if ent is not None and all(r() is p for r, p in zip(ent[0], params)) and ent[1] == vers:
        return ent[2]
    conv = torch.cat([p.detach().to(dtype) for p in params], dim=0).contiguous()
    refs = tuple(weakref.ref(p, lambda _r, k=key: _shadow.pop(k, None)) for p in params)
    _shadow[key] = (refs, vers, conv)
    return conv


def _gate32(g: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    return None if g is None else f32(g).reshape(1)


def _as2d(t: torch.Tensor) -> torch.Tensor:
    t2 = t.reshape(-1, t.shape[-1])
    if t2.dtype != BF16:
        t2 = t2.to(BF16)
    return t2 if t2.is_contiguous() else t2.contiguous()


def _new(rows, cols, like, dtype=BF16):
    return torch.empty((rows, cols), dtype=dtype, device=like.device)


def _wgrad(dy2, x2, param, *, gate=None, alpha=1.0):
    """dW[N_out, K_in] = alpha * tanh?(gate) * dy^T x, in the parameter's dtype (fp32 master or bf16)."""
    flat = getattr(param, "_mmgl_grad_flat", None)
    if flat is not None and not param._mmgl_grad_slot_used and param.dtype == F32:
        # train.FlatGradSync: the gradient is written straight into the parameter's slot of the flat all-reduce buffer
        # (a fresh view object, so autograd adopts it as param.grad); a second use in the same backward gets its own tensor
        out = flat[0][flat[1]:flat[1] + param.numel()].view(param.shape)
        param._mmgl_grad_slot_used = True
    else:
        out = torch.empty(param.shape, dtype=param.dtype if param.dtype in (BF16, F32) else F32, device=dy2.device)
    K.gemm(dy2, x2, out, a_t=True, b_t=True, gate=gate, alpha=alpha)
    return out


def _bgrad(dy2, param, *, gate=None, scale=1.0):
    out = torch.empty(param.shape, dtype=F32, device=dy2.device)
    K.colsum(dy2, out, scale=scale, gate=gate)
    return out if param.dtype == F32 else out.to(param.dtype)


def _ln_fwd(x2, gamma, beta, eps):
    y = torch.empty_like(x2)
    mean = torch.empty(x2.shape[0], dtype=F32, device=x2.device)
    rstd = torch.empty_like(mean)
    K.layernorm_fwd(x2, gamma, beta, y, mean, rstd, eps)
    return y, mean, rstd


def _ln_bwd(dy2, x2, gamma, mean, rstd, d_res, want_affine, gamma_param, beta_param):
    dx = torch.empty_like(x2)
    dg = db = None
    if want_affine:
        dg = torch.empty(x2.shape[1], dtype=F32, device=x2.device)
        db = torch.empty_like(dg)
    K.layernorm_bwd(dy2, x2, gamma, mean, rstd, d_res, dx, dg, db)
    if want_affine:
        if gamma_param.dtype != F32:
            dg = dg.to(gamma_param.dtype)
        if beta_param.dtype != F32:
            db = db.to(beta_param.dtype)
    return dx, dg, db


def _scalar_grad(dy2, a2, gate32, param):
    out = torch.empty(1, dtype=F32, device=dy2.device)
    K.gate_grad(dy2, a2, gate32, out)
    return out.reshape(param.shape).to(param.dtype)


def mask_u8(mask: torch.Tensor) -> torch.Tensor:
    """[B,Nk] bool / {0,1} mask, or the reference's additive [B,1,S,Nk] mask (0 = attend), -> uint8 [B,Nk]."""
    if mask.dim() == 4:  # additive, model/modelling_cross_attention.py:68-79
        mask = mask[:, 0, 0, :] == 0
    if mask.dtype != torch.uint8:
        mask = (mask != 0).to(torch.uint8)
    return mask.contiguous()


# --------------------------------------------------------------------------------------------- linear
class LinearFn(torch.autograd.Function):
    """y = dropout(alpha * (x W^T + b)) (+ residual).  nn.Linear call sites: model/modelling_cross_attention.py:194,
    198-199, 273 (+ :332 dropout, :337 residual), 826, 997, 1020; model/modelling_self_attention.py:170, 193, 313."""

    @staticmethod
    def forward(ctx, x, weight, bias, residual, alpha, dropout_p):
        x2 = _as2d(x)
        w = w16(weight)
        y = _new(x2.shape[0], w.shape[0], x2)
        r2 = _as2d(residual) if residual is not None else None
        seed = next_dropout_seed() if dropout_p > 0.0 else 0
        K.gemm(x2, w, y, bias=f32(bias), alpha=alpha, residual=r2, dropout_p=dropout_p, dropout_seed=seed)
        ctx.save_for_backward(x2, weight, bias)
        ctx.alpha = alpha
        ctx.drop = (dropout_p, seed)
        ctx.has_res = residual is not None
        ctx.x_shape = x.shape
        return y.reshape(*x.shape[:-1], w.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x2, weight, bias = ctx.saved_tensors
        dy2 = _undrop(_as2d(dy), *ctx.drop)
        dx = dw = db = dres = None
        if ctx.needs_input_grad[0]:
            dx = _new(x2.shape[0], x2.shape[1], x2)
            K.gemm(dy2, w16(weight), dx, b_t=True, alpha=ctx.alpha)
            dx = dx.reshape(ctx.x_shape)
        if ctx.needs_input_grad[1]:
            dw = _wgrad(dy2, x2, weight, alpha=ctx.alpha)
        if bias is not None and ctx.needs_input_grad[2]:
            db = _bgrad(dy2, bias, scale=ctx.alpha)
        if ctx.has_res and ctx.needs_input_grad[3]:
            dres = dy
        return dx, dw, db, dres, None, None


def linear(x, weight, bias=None, residual=None, alpha=1.0, dropout_p=0.0):
    return LinearFn.apply(x, weight, bias, residual, alpha, float(dropout_p))


class LoRALinearFn(torch.autograd.Function):
    """y = x W^T + b + s * (x A^T) B^T with the rank-r update accumulated into the SAME TMEM tile as the base
    product (second operand pair of mmgl_gemm_bf16), s = lora_alpha / r folded into the stored intermediate.

    LoRA as configured at model/modelling_self_attention.py:80-87 (peft; arithmetic restated, parity unpinned)."""

    @staticmethod
    def forward(ctx, x, weight, bias, lora_a, lora_b, scale):
        x2 = _as2d(x)
        w, a, b = w16(weight), w16(lora_a), w16(lora_b)
        r = a.shape[0]
        t = _new(x2.shape[0], r, x2)
        K.gemm(x2, a, t, alpha=scale)                       # t = s * x A^T            [M, r]
        y = _new(x2.shape[0], w.shape[0], x2)
        K.gemm(x2, w, y, a1=t, b1=b, bias=f32(bias))        # y = x W^T + t B^T + b
        ctx.save_for_backward(x2, t, weight, bias, lora_a, lora_b)
        ctx.scale = scale
        ctx.x_shape = x.shape
        return y.reshape(*x.shape[:-1], w.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x2, t, weight, bias, lora_a, lora_b = ctx.saved_tensors
        dy2 = _as2d(dy)
        s = ctx.scale
        a, b = w16(lora_a), w16(lora_b)
        dt = _new(x2.shape[0], a.shape[0], x2)
        K.gemm(dy2, b, dt, b_t=True, alpha=s)               # dt = s * dy B             [M, r]
        dx = dw = dbias = da = db = None
        if ctx.needs_input_grad[0]:
            dx = _new(x2.shape[0], x2.shape[1], x2)
            K.gemm(dy2, w16(weight), dx, b_t=True, a1=dt, b1=a)   # dx = dy W + dt A
            dx = dx.reshape(ctx.x_shape)
        if ctx.needs_input_grad[1]:
            dw = _wgrad(dy2, x2, weight)
        if bias is not None and ctx.needs_input_grad[2]:
            dbias = _bgrad(dy2, bias)
        if ctx.needs_input_grad[3]:
            da = _wgrad(dt, x2, lora_a)                     # dA = dt^T x               [r, K]
        if ctx.needs_input_grad[4]:
            db = _wgrad(dy2, t, lora_b)                     # dB = dy^T (s x A^T)       [N, r]
        return dx, dw, dbias, da, db, None


def lora_linear(x, weight, bias, lora_a, lora_b, scale):
    return LoRALinearFn.apply(x, weight, bias, lora_a, lora_b, scale)


# --------------------------------------------------------------------------------------------- layernorm
class LayerNormFn(torch.autograd.Function):
    """nn.LayerNorm over the last dim: model/modelling_cross_attention.py:320, 341, 350, 365, 635-636."""

    @staticmethod
    def forward(ctx, x, weight, bias, eps):
        x2 = _as2d(x)
        g = f32(weight)
        y, mean, rstd = _ln_fwd(x2, g, f32(bias), eps)
        ctx.save_for_backward(x2, weight, bias, mean, rstd)
        ctx.x_shape = x.shape
        return y.reshape(x.shape)

    @staticmethod
    def backward(ctx, dy):
        x2, weight, bias, mean, rstd = ctx.saved_tensors
        want = ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        dx, dg, db = _ln_bwd(_as2d(dy), x2, f32(weight), mean, rstd, None, want, weight, bias)
        return dx.reshape(ctx.x_shape), dg, db, None


def layer_norm(x, weight, bias, eps=1e-5):
    return LayerNormFn.apply(x, weight, bias, eps)


class LayerNormForkFn(torch.autograd.Function):
    """The fork at the top of a pre-LN residual block: returns (LayerNorm(x), x).  Feeding the second output to the
    residual input of the block's last GEMM makes both gradients of x -- through the norm and through the residual --
    arrive at THIS backward, where one kernel forms dx = d_res + LN'(dy); autograd would otherwise add them in a separate
    pass over [rows, H] (41 such adds per cfg2 step)."""

    @staticmethod
    def forward(ctx, x, weight, bias, eps):
        x2 = _as2d(x)
        y, mean, rstd = _ln_fwd(x2, f32(weight), f32(bias), eps)
        ctx.save_for_backward(x2, weight, bias, mean, rstd)
        ctx.x_shape = x.shape
        ctx.set_materialize_grads(False)
        return y.reshape(x.shape), x.view_as(x)

    @staticmethod
    def backward(ctx, dy, d_res):
        x2, weight, bias, mean, rstd = ctx.saved_tensors
        want = ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        dy2 = torch.zeros_like(x2) if dy is None else _as2d(dy)
        dx, dg, db = _ln_bwd(dy2, x2, f32(weight), mean, rstd, None if d_res is None else _as2d(d_res), want, weight, bias)
        return dx.reshape(ctx.x_shape), dg, db, None


def layer_norm_fork(x, weight, bias, eps=1e-5):
    """(LayerNorm(x), x): use the second value as the block's residual (see LayerNormForkFn)."""
    if x.dtype != BF16:
        return layer_norm(x, weight, bias, eps), x
    return LayerNormForkFn.apply(x, weight, bias, eps)


class RMSNormFn(torch.autograd.Function):
    """T5LayerNorm (HF models/t5/modeling_t5.py:46-70): y = weight * x * rsqrt(mean(x^2) + eps), fp32 statistics."""

    @staticmethod
    def forward(ctx, x, weight, eps):
        x2 = _as2d(x)
        y = torch.empty_like(x2)
        rstd = torch.empty(x2.shape[0], dtype=F32, device=x2.device)
        K.rmsnorm_fwd(x2, f32(weight), y, rstd, eps)
        ctx.save_for_backward(x2, weight, rstd)
        ctx.x_shape = x.shape
        return y.reshape(x.shape)

    @staticmethod
    def backward(ctx, dy):
        x2, weight, rstd = ctx.saved_tensors
        dx = torch.empty_like(x2)
        dg = torch.empty(x2.shape[1], dtype=F32, device=x2.device) if ctx.needs_input_grad[1] else None
        K.rmsnorm_bwd(_as2d(dy), x2, f32(weight), rstd, None, dx, dg)
        if dg is not None and weight.dtype != F32:
            dg = dg.to(weight.dtype)
        return dx.reshape(ctx.x_shape), dg, None


def rms_norm(x, weight, eps=1e-6):
    return RMSNormFn.apply(x, weight, float(eps))


class RMSNormForkFn(torch.autograd.Function):
    """(RMSNorm(x), x) for the pre-norm residual blocks of T5: see LayerNormForkFn."""

    @staticmethod
    def forward(ctx, x, weight, eps):
        x2 = _as2d(x)
        y = torch.empty_like(x2)
        rstd = torch.empty(x2.shape[0], dtype=F32, device=x2.device)
        K.rmsnorm_fwd(x2, f32(weight), y, rstd, eps)
        ctx.save_for_backward(x2, weight, rstd)
        ctx.x_shape = x.shape
        ctx.set_materialize_grads(False)
        return y.reshape(x.shape), x.view_as(x)

    @staticmethod
    def backward(ctx, dy, d_res):
        x2, weight, rstd = ctx.saved_tensors
        dx = torch.empty_like(x2)
        dg = torch.empty(x2.shape[1], dtype=F32, device=x2.device) if ctx.needs_input_grad[1] else None
        dy2 = torch.zeros_like(x2) if dy is None else _as2d(dy)
        K.rmsnorm_bwd(dy2, x2, f32(weight), rstd, None if d_res is None else _as2d(d_res), dx, dg)
        if dg is not None and weight.dtype != F32:
            dg = dg.to(weight.dtype)
        return dx.reshape(ctx.x_shape), dg, None


def rms_norm_fork(x, weight, eps=1e-6):
    if x.dtype != BF16:
        return rms_norm(x, weight, eps), x
    return RMSNormForkFn.apply(x, weight, float(eps))


class DropoutFn(torch.autograd.Function):
    """nn.Dropout on an activation tensor with the library's counter-based mask (no mask tensor is stored)."""

    @staticmethod
    def forward(ctx, x, p):
        x2 = _as2d(x)
        seed = next_dropout_seed()
        y = torch.empty_like(x2)
        K.dropout_apply(x2, y, p, seed)
        ctx.drop = (p, seed)
        ctx.x_shape = x.shape
        return y.reshape(x.shape)

    @staticmethod
    def backward(ctx, dy):
        return _undrop(_as2d(dy), *ctx.drop).reshape(ctx.x_shape), None


def dropout(x, p):
    """x (bf16 [..., H]) -> dropout(x); identity when p == 0."""
    if p <= 0.0:
        return x
    return DropoutFn.apply(x, float(p))


# --------------------------------------------------------------------------------------------- cross-entropy
class CrossEntropyFn(torch.autograd.Function):
    """mean softmax cross-entropy over bf16 logits [rows, V] with int64 labels [rows] (ignore_index rows skipped):
    nn.CrossEntropyLoss at model/modelling_cross_attention.py:835-836.  fp32 math, no fp32 copy of the logits."""

    @staticmethod
    def forward(ctx, logits2, labels, ignore_index):
        assert logits2.dim() == 2 and logits2.dtype == BF16
        rows = logits2.shape[0]
        labels = labels.reshape(-1).to(torch.int64).contiguous()
        dev = logits2.device
        lse = torch.empty(rows, dtype=F32, device=dev)
        row_loss = torch.empty(rows, dtype=F32, device=dev)
        loss = torch.empty(1, dtype=F32, device=dev)
        count = torch.empty(1, dtype=F32, device=dev)
        K.ce_fwd(logits2, labels, lse, row_loss, loss, count, ignore_index)
        ctx.save_for_backward(logits2, labels, lse, count)
        ctx.ignore_index = ignore_index
        return loss.reshape(())

    @staticmethod
    def backward(ctx, dloss):
        logits2, labels, lse, count = ctx.saved_tensors
        dlogits = torch.empty(logits2.shape, dtype=BF16, device=logits2.device)
        K.ce_bwd(logits2, labels, lse, dloss.reshape(1).to(F32).contiguous(), count, dlogits, ctx.ignore_index)
        return dlogits, None, None


def cross_entropy(logits, labels, ignore_index=-100):
    """logits [..., V] bf16, labels [...] -> scalar fp32 mean loss over the non-ignored positions."""
    l2 = logits.reshape(-1, logits.shape[-1])
    return CrossEntropyFn.apply(l2, labels, int(ignore_index))


def shifted_cross_entropy(logits, labels, ignore_index=-100):
    """Causal-LM loss (model/modelling_cross_attention.py:828-836): position s predicts token s+1.  The shift is applied
    to the small label tensor (last position ignored) instead of slicing/copying the [B,S,V] logits."""
    b, s = labels.shape
    shifted = torch.full((b, s), ignore_index, dtype=torch.int64, device=labels.device)
    shifted[:, :-1] = labels[:, 1:]
    return cross_entropy(logits, shifted, ignore_index)


# --------------------------------------------------------------------------------------------- MLP
class MLPFn(torch.autograd.Function):
    """y = residual + dropout(fc2(dropout_h(relu(fc1(x))))): model/modelling_cross_attention.py:352-361 (non-gated form;
    dropout_h = 0) and HF T5DenseActDense + T5LayerFF (dropout on the hidden activation as well).  ReLU, its backward
    mask and both dropouts live in the GEMM epilogues; no [M,F] elementwise pass touches HBM."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, residual, dropout_p, hidden_dropout_p):
        x2 = _as2d(x)
        f = _new(x2.shape[0], w1.shape[0], x2)
        seed_h = next_dropout_seed() if hidden_dropout_p > 0.0 else 0
        K.gemm(x2, w16(w1), f, bias=f32(b1), relu=True, dropout_p=hidden_dropout_p, dropout_seed=seed_h)
        y = _new(x2.shape[0], w2.shape[0], x2)
        r2 = _as2d(residual) if residual is not None else None
        seed = next_dropout_seed() if dropout_p > 0.0 else 0
        K.gemm(f, w16(w2), y, bias=f32(b2), residual=r2, dropout_p=dropout_p, dropout_seed=seed)
        ctx.save_for_backward(x2, f, w1, b1, w2, b2)
        ctx.has_res = residual is not None
        ctx.drop = (dropout_p, seed)
        ctx.drop_h = (hidden_dropout_p, seed_h)
        ctx.x_shape = x.shape
        return y.reshape(*x.shape[:-1], w2.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x2, f, w1, b1, w2, b2 = ctx.saved_tensors
        dy2 = _undrop(_as2d(dy), *ctx.drop)
        df = torch.empty_like(f)
        # f = dropout_h(relu(.)) is zero where either zeroed it; the kept entries carry the 1 / (1 - p) of dropout_h
        K.gemm(dy2, w16(w2), df, b_t=True, relu_mask=f, dropout_p=ctx.drop_h[0], dropout_seed=ctx.drop_h[1])
        dx = dw1 = db1 = dw2 = db2 = dres = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x2)
            K.gemm(df, w16(w1), dx, b_t=True)
            dx = dx.reshape(ctx.x_shape)
        if ctx.needs_input_grad[1]:
            dw1 = _wgrad(df, x2, w1)
        if b1 is not None and ctx.needs_input_grad[2]:
            db1 = _bgrad(df, b1)
        if ctx.needs_input_grad[3]:
            dw2 = _wgrad(dy2, f, w2)
        if b2 is not None and ctx.needs_input_grad[4]:
            db2 = _bgrad(dy2, b2)
        if ctx.has_res and ctx.needs_input_grad[5]:
            dres = dy
        return dx, dw1, db1, dw2, db2, dres, None, None


def mlp(x, w1, b1, w2, b2, residual=None, dropout_p=0.0, hidden_dropout_p=0.0):
    return MLPFn.apply(x, w1, b1, w2, b2, residual, float(dropout_p), float(hidden_dropout_p))


# --------------------------------------------------------------------------------------------- attention core
class XAttnCoreFn(torch.autograd.Function):
    """O = softmax(max(Q K^T + mask, finfo.min)) V per (sample, head); Q pre-scaled.
    model/modelling_cross_attention.py:176-177, 206-271 (+ :68-79)."""

    @staticmethod
    def forward(ctx, q, k, v, mask, heads):
        b, s, h = q.shape
        nk = k.shape[1]
        q2, k2, v2 = _as2d(q), _as2d(k), _as2d(v)
        o = torch.empty_like(q2)
        stats = torch.empty((b, heads, s, 2), dtype=F32, device=q.device)
        m8 = mask_u8(mask)
        K.xattn_fwd(q2, k2, v2, m8, o, stats, b, s, nk, heads, h // heads)
        ctx.save_for_backward(q2, k2, v2, o, stats, m8)
        ctx.dims = (b, s, nk, heads, h)
        return o.reshape(b, s, h)

    @staticmethod
    def backward(ctx, d_o):
        q2, k2, v2, o, stats, m8 = ctx.saved_tensors
        b, s, nk, heads, h = ctx.dims
        dq, dk, dv = torch.empty_like(q2), torch.empty_like(k2), torch.empty_like(v2)
        K.xattn_bwd(_as2d(d_o), q2, k2, v2, o, stats, m8, dq, dk, dv, b, s, nk, heads, h // heads)
        return dq.reshape(b, s, h), dk.reshape(b, nk, h), dv.reshape(b, nk, h), None, None


def xattn_core(q, k, v, mask, heads):
    return XAttnCoreFn.apply(q, k, v, mask, heads)


def xattn_max_keys(head_dim: int) -> int:
    """Longest neighbor bank (rows = neighbors x tokens per neighbor) the cross-attention core holds on chip per head:
    the limits mmgl_xattn_fwd / _bwd enforce (include/mmgl_b200.h).  0 = head_dim not supported."""
    return {64: 256, 128: 128}.get(int(head_dim), 0)


def _key_mask_u8(key_mask):
    if key_mask is None:
        return None
    km = key_mask if key_mask.dtype == torch.uint8 else (key_mask != 0).to(torch.uint8)
    return km.contiguous()


class SelfAttnFn(torch.autograd.Function):
    """Causal / key-padded self-attention over a fused QKV projection [B,S,3H] (thirds = Q | K | V, heads interleaved
    in H): model/modelling_cross_attention.py:201-275 with the mask of :455-476.  Q is NOT pre-scaled: ``scale`` is
    applied to the scores inside the kernel.  Backward writes dQ | dK | dV straight into one [B,S,3H] buffer."""

    @staticmethod
    def forward(ctx, qkv, key_mask, heads, causal, scale, dropout_p):
        b, s, h3 = qkv.shape
        h = h3 // 3
        qkv2 = _as2d(qkv)
        km = _key_mask_u8(key_mask)
        o = _new(b * s, h, qkv2)
        stats = torch.empty((b, heads, s, 2), dtype=F32, device=qkv.device)
        seed = next_dropout_seed() if dropout_p > 0.0 else 0
        K.attn_fwd(qkv2[:, :h], qkv2[:, h:2 * h], qkv2[:, 2 * h:], km, None, o, stats, b, s, s, heads, h // heads, scale,
                   causal, dropout_p, seed)
        ctx.save_for_backward(qkv2, o, stats, km)
        ctx.cfg = (b, s, h, heads, bool(causal), float(scale), dropout_p, seed)
        return o.reshape(b, s, h)

    @staticmethod
    def backward(ctx, d_o):
        qkv2, o, stats, km = ctx.saved_tensors
        b, s, h, heads, causal, scale, dropout_p, seed = ctx.cfg
        dqkv = torch.empty_like(qkv2)
        K.attn_bwd(_as2d(d_o), qkv2[:, :h], qkv2[:, h:2 * h], qkv2[:, 2 * h:], km, None, o, stats, dqkv[:, :h],
                   dqkv[:, h:2 * h], dqkv[:, 2 * h:], b, s, s, heads, h // heads, scale, causal, dropout_p, seed)
        return dqkv.reshape(b, s, 3 * h), None, None, None, None, None


def self_attention(qkv, key_mask, heads, causal=True, scale=None, dropout_p=0.0):
    """qkv [B,S,3H] bf16 -> [B,S,H]; key_mask [B,S] (1 = real token) or None."""
    if scale is None:
        scale = (qkv.shape[-1] // 3 // heads) ** -0.5
    return SelfAttnFn.apply(qkv, key_mask, heads, bool(causal), float(scale), float(dropout_p))


class AttnFn(torch.autograd.Function):
    """General attention core over separate projections: q [B,Sq,H], k / v [B,Sk,H] (heads interleaved in H),
    key_mask [B,Sk], optional additive relative-position bias rel_bias fp32 [heads, Sq+Sk-1] (bias of (row, key) =
    rel_bias[h, key - row + Sq - 1]; differentiable: a trainable T5 bias table gets its gradient) and dropout on the
    probabilities.  This is the attention of
    the HF T5 / OPT language model that the concat path runs (model/modelling_self_attention.py:332): T5 uses
    scale = 1, a bucketed relative-position bias and, in the decoder, cross-attention with Sq != Sk."""

    @staticmethod
    def forward(ctx, q, k, v, key_mask, rel_bias, heads, causal, scale, dropout_p):
        b, sq, h = q.shape
        sk = k.shape[1]
        q2, k2, v2 = _as2d(q), _as2d(k), _as2d(v)
        km = _key_mask_u8(key_mask)
        rb = None if rel_bias is None else rel_bias.detach().to(F32).contiguous()
        o = _new(b * sq, h, q2)
        stats = torch.empty((b, heads, sq, 2), dtype=F32, device=q.device)
        seed = next_dropout_seed() if dropout_p > 0.0 else 0
        K.attn_fwd(q2, k2, v2, km, rb, o, stats, b, sq, sk, heads, h // heads, scale, causal, dropout_p, seed)
        ctx.save_for_backward(q2, k2, v2, o, stats, km, rb)
        ctx.cfg = (b, sq, sk, h, heads, bool(causal), float(scale), dropout_p, seed)
        ctx.bias_dtype = None if rel_bias is None else rel_bias.dtype
        return o.reshape(b, sq, h)

    @staticmethod
    def backward(ctx, d_o):
        q2, k2, v2, o, stats, km, rb = ctx.saved_tensors
        b, sq, sk, h, heads, causal, scale, dropout_p, seed = ctx.cfg
        dq, dk, dv = torch.empty_like(q2), torch.empty_like(k2), torch.empty_like(v2)
        d_rb = torch.zeros_like(rb) if (rb is not None and ctx.needs_input_grad[4]) else None
        K.attn_bwd(_as2d(d_o), q2, k2, v2, km, rb, o, stats, dq, dk, dv, b, sq, sk, heads, h // heads, scale, causal,
                   dropout_p, seed, d_rel_bias=d_rb)
        if d_rb is not None and ctx.bias_dtype != F32:
            d_rb = d_rb.to(ctx.bias_dtype)
        return dq.reshape(b, sq, h), dk.reshape(b, sk, h), dv.reshape(b, sk, h), None, d_rb, None, None, None, None


def attention(q, k, v, key_mask=None, rel_bias=None, heads=1, causal=False, scale=1.0, dropout_p=0.0):
    return AttnFn.apply(q, k, v, key_mask, rel_bias, heads, bool(causal), float(scale), float(dropout_p))


# --------------------------------------------------------------------------------------------- gated cross layer
class GatedCrossLayerFn(torch.autograd.Function):
    """One whole gated cross-attention block, forward and backward, as a fixed schedule of fused kernels.

    model/modelling_cross_attention.py:304-375 (MPTDecoderLayer, cross_attention=True) with MPTAttention :179-275:

        pre-LN :  h1 = x  + tanh(g1) * Attn(LN1(x), bank);   y = h1 + tanh(g2) * fc2(relu(fc1(LN2(h1))))
        post-LN:  h1 = LN1(x + tanh(g1) * Attn(x, bank));    y = LN2(h1 + tanh(g2) * fc2(relu(fc1(h1))))

    g1 = g2 = None gives the plain residual form (peft_type != "flamingo").  Forward is 2 LayerNorms, 6 GEMMs
    (bias, d^-1/2 scale, ReLU, tanh-gate and residual fused into their epilogues) and the attention core; the
    pre-gate branch outputs are kept (aux epilogue output) for the scalar gate gradients.
    """

    @staticmethod
    def forward(ctx, x, bank, mask, ln1_w, ln1_b, wq, bq, wk, bk, wv, bv, wo, bo, g1,
                ln2_w, ln2_b, w1, b1, w2, b2, g2, heads, eps, pre_ln, dropout_p):
        bsz, s, h = x.shape
        nk = bank.shape[1]
        d = h // heads
        scaling = float(d) ** -0.5
        x2, bank2 = _as2d(x), _as2d(bank)
        m8 = mask_u8(mask)
        m_rows = x2.shape[0]
        g1f, g2f = _gate32(g1), _gate32(g2)
        ln1g, ln2g = f32(ln1_w), f32(ln2_w)

        if pre_ln:
            a_in, mean1, rstd1 = _ln_fwd(x2, ln1g, f32(ln1_b), eps)
        else:
            a_in, mean1, rstd1 = x2, None, None
        q = _new(m_rows, h, x2)
        K.gemm(a_in, w16(wq), q, bias=f32(bq), alpha=scaling)                    # :194
        kv = _new(bank2.shape[0], 2 * h, x2)
        k, v = kv[:, :h], kv[:, h:]
        K.gemm(bank2, w16(wk), k, bias=f32(bk))                                   # :198
        K.gemm(bank2, w16(wv), v, bias=f32(bv))                                   # :199
        o = _new(m_rows, h, x2)
        stats = torch.empty((bsz, heads, s, 2), dtype=F32, device=x.device)
        K.xattn_fwd(q, k, v, m8, o, stats, bsz, s, nk, heads, d)                  # :206-271
        seed1 = next_dropout_seed() if dropout_p > 0.0 else 0
        seed2 = next_dropout_seed() if dropout_p > 0.0 else 0
        a_out = _new(m_rows, h, x2) if g1 is not None else None
        u = _new(m_rows, h, x2)
        K.gemm(o, w16(wo), u, bias=f32(bo), aux=a_out, gate=g1f, residual=x2,     # :273, :332-337
               dropout_p=dropout_p, dropout_seed=seed1)
        if pre_ln:
            h1 = u
            f_in, mean2, rstd2 = _ln_fwd(h1, ln2g, f32(ln2_b), eps)              # :350
        else:
            h1, mean1, rstd1 = _ln_fwd(u, ln1g, f32(ln1_b), eps)                 # :341
            f_in, mean2, rstd2 = h1, None, None
        f = _new(m_rows, w1.shape[0], x2)
        K.gemm(f_in, w16(w1), f, bias=f32(b1), relu=True)                         # :352-353
        c_out = _new(m_rows, h, x2) if g2 is not None else None
        wsum = _new(m_rows, h, x2)
        K.gemm(f, w16(w2), wsum, bias=f32(b2), aux=c_out, gate=g2f, residual=h1,  # :355-361
               dropout_p=dropout_p, dropout_seed=seed2)
        if pre_ln:
            y = wsum
        else:
            y, mean2, rstd2 = _ln_fwd(wsum, ln2g, f32(ln2_b), eps)               # :365

        ctx.pre_ln = pre_ln
        ctx.drop1, ctx.drop2 = (dropout_p, seed1), (dropout_p, seed2)
        ctx.dims = (bsz, s, nk, heads, h)
        ctx.eps = eps
        ctx.x_shape, ctx.bank_shape = x.shape, bank.shape
        ctx.x_dtype, ctx.bank_dtype = x.dtype, bank.dtype
        # activations: pre-LN keeps LN outputs a_in / f_in; post-LN keeps the LN inputs u / wsum
        ctx.save_for_backward(x2, bank2, m8, a_in, mean1, rstd1, q, kv, o, stats, a_out, u, h1, f_in, mean2, rstd2,
                              f, c_out, wsum,
                              ln1_w, ln1_b, wq, bq, wk, bk, wv, bv, wo, bo, g1, ln2_w, ln2_b, w1, b1, w2, b2, g2)
        return y.reshape(x.shape)

    @staticmethod
    def backward(ctx, dy):
        (x2, bank2, m8, a_in, mean1, rstd1, q, kv, o, stats, a_out, u, h1, f_in, mean2, rstd2, f, c_out, wsum,
         ln1_w, ln1_b, wq, bq, wk, bk, wv, bv, wo, bo, g1, ln2_w, ln2_b, w1, b1, w2, b2, g2) = ctx.saved_tensors
        bsz, s, nk, heads, h = ctx.dims
        d = h // heads
        scaling = float(d) ** -0.5
        pre_ln = ctx.pre_ln
        need = ctx.needs_input_grad
        g1f, g2f = _gate32(g1), _gate32(g2)
        ln1g, ln2g = f32(ln1_w), f32(ln2_w)
        dy2 = _as2d(dy)
        k, v = kv[:, :h], kv[:, h:]
        (i_x, i_bank, _i_mask, i_ln1w, i_ln1b, i_wq, i_bq, i_wk, i_bk, i_wv, i_bv, i_wo, i_bo, i_g1,
         i_ln2w, i_ln2b, i_w1, i_b1, i_w2, i_b2, i_g2) = range(21)
        grads = [None] * 25

        # ---- FFN branch
        if pre_ln:
            dw_ = dy2                                                          # grad of wsum (= y)
        else:
            dw_, grads[i_ln2w], grads[i_ln2b] = _ln_bwd(dy2, wsum, ln2g, mean2, rstd2, None,
                                                        need[i_ln2w] or need[i_ln2b], ln2_w, ln2_b)
        dres2 = dw_                                                            # residual path of the FFN branch
        if g2 is not None and need[i_g2]:
            grads[i_g2] = _scalar_grad(dw_, c_out, g2f, g2)       # c_out is the post-dropout, pre-gate branch value
        dw_ = _undrop(dw_, *ctx.drop2)                            # from here on: gradient of the pre-dropout fc2 output
        df = torch.empty_like(f)
        K.gemm(dw_, w16(w2), df, b_t=True, relu_mask=f, gate=g2f)             # d relu(fc1) (mask, then gate)
        if need[i_w2]:
            grads[i_w2] = _wgrad(dw_, f, w2, gate=g2f)
        if b2 is not None and need[i_b2]:
            grads[i_b2] = _bgrad(dw_, b2, gate=g2f)
        if need[i_w1]:
            grads[i_w1] = _wgrad(df, f_in, w1)
        if b1 is not None and need[i_b1]:
            grads[i_b1] = _bgrad(df, b1)
        if pre_ln:
            dln2 = torch.empty_like(h1)
            K.gemm(df, w16(w1), dln2, b_t=True)
            dh1, grads[i_ln2w], grads[i_ln2b] = _ln_bwd(dln2, h1, ln2g, mean2, rstd2, dres2,
                                                        need[i_ln2w] or need[i_ln2b], ln2_w, ln2_b)
            du = dh1                                                           # grad of u (= h1)
        else:
            dh1 = torch.empty_like(h1)
            K.gemm(df, w16(w1), dh1, b_t=True, residual=dres2)
            du, grads[i_ln1w], grads[i_ln1b] = _ln_bwd(dh1, u, ln1g, mean1, rstd1, None,
                                                       need[i_ln1w] or need[i_ln1b], ln1_w, ln1_b)

        # ---- attention branch
        if g1 is not None and need[i_g1]:
            grads[i_g1] = _scalar_grad(du, a_out, g1f, g1)
        dud = _undrop(du, *ctx.drop1)                                          # gradient of the pre-dropout out_proj output
        d_o = torch.empty_like(o)
        K.gemm(dud, w16(wo), d_o, b_t=True, gate=g1f)
        if need[i_wo]:
            grads[i_wo] = _wgrad(dud, o, wo, gate=g1f)
        if bo is not None and need[i_bo]:
            grads[i_bo] = _bgrad(dud, bo, gate=g1f)
        dq = torch.empty_like(q)
        dkv = torch.empty_like(kv)
        dk, dv = dkv[:, :h], dkv[:, h:]
        K.xattn_bwd(d_o, q, k, v, o, stats, m8, dq, dk, dv, bsz, s, nk, heads, d)
        if need[i_wq]:
            grads[i_wq] = _wgrad(dq, a_in, wq, alpha=scaling)
        if bq is not None and need[i_bq]:
            grads[i_bq] = _bgrad(dq, bq, scale=scaling)
        if need[i_wk]:
            grads[i_wk] = _wgrad(dk, bank2, wk)
        if bk is not None and need[i_bk]:
            grads[i_bk] = _bgrad(dk, bk)
        if need[i_wv]:
            grads[i_wv] = _wgrad(dv, bank2, wv)
        if bv is not None and need[i_bv]:
            grads[i_bv] = _bgrad(dv, bv)
        if need[i_bank]:
            dbank = torch.empty_like(bank2)
            K.gemm(dk, w16(wk), dbank, b_t=True, a1=dv, b1=w16(wv))           # dK Wk + dV Wv in one TMEM tile
            grads[i_bank] = dbank.reshape(ctx.bank_shape).to(ctx.bank_dtype)
        if need[i_x]:
            if pre_ln:
                dln1 = torch.empty_like(x2)
                K.gemm(dq, w16(wq), dln1, b_t=True, alpha=scaling)
                dx, grads[i_ln1w], grads[i_ln1b] = _ln_bwd(dln1, x2, ln1g, mean1, rstd1, du,
                                                           need[i_ln1w] or need[i_ln1b], ln1_w, ln1_b)
            else:
                dx = torch.empty_like(x2)
                K.gemm(dq, w16(wq), dx, b_t=True, alpha=scaling, residual=du)
            grads[i_x] = dx.reshape(ctx.x_shape).to(ctx.x_dtype)
        elif pre_ln and (need[i_ln1w] or need[i_ln1b]):
            dln1 = torch.empty_like(x2)
            K.gemm(dq, w16(wq), dln1, b_t=True, alpha=scaling)
            _, grads[i_ln1w], grads[i_ln1b] = _ln_bwd(dln1, x2, ln1g, mean1, rstd1, du, True, ln1_w, ln1_b)
        return tuple(grads)


def gated_cross_layer(x, bank, mask, ln1_w, ln1_b, wq, bq, wk, bk, wv, bv, wo, bo, g1, ln2_w, ln2_b, w1, b1, w2, b2,
                      g2, heads, eps=1e-5, pre_ln=True, dropout_p=0.0):
    return GatedCrossLayerFn.apply(x, bank, mask, ln1_w, ln1_b, wq, bq, wk, bk, wv, bv, wo, bo, g1,
                                   ln2_w, ln2_b, w1, b1, w2, b2, g2, heads, eps, pre_ln, float(dropout_p))


# --------------------------------------------------------------------------------------------- neighbor bank
class BankPackFn(torch.autograd.Function):
    """Ragged interleave of projected text/image neighbor embeddings into the bank [B,(T+I)*n_tok,H] + byte mask,
    with the position-embedding gather-add and the Laplacian-PE projection fused in.

    model/modelling_cross_attention.py:999-1004, 1022-1027, 1080-1104; model/modelling_self_attention.py:284-315."""

    @staticmethod
    def forward(ctx, text_proj, text_pos_table, text_pos_ids, text_locations,
                image_proj, image_pos_table, image_pos_ids, image_locations, lpe, lpe_weight, lpe_bias, n_tok):
        ref = text_proj if text_proj is not None else image_proj
        bsz = ref.shape[0]
        n_text = text_proj.shape[1] if text_proj is not None else 0
        n_image = image_proj.shape[1] if image_proj is not None else 0
        row_width = ref.shape[-1]
        n_src = n_text + n_image
        dev = ref.device
        a = K.BankArgs()
        keep = []

        def c16(t):
            if t is None:
                return None
            t = t.to(BF16) if t.dtype != BF16 else t
            t = t.contiguous()
            keep.append(t)
            return t

        def ci64(t):
            if t is None:
                return None
            t = t.to(torch.int64).contiguous()
            keep.append(t)
            return t

        tp, ip = c16(text_proj), c16(image_proj)
        tt, it = w16(text_pos_table), w16(image_pos_table)
        tpos, tloc, ipos, iloc = ci64(text_pos_ids), ci64(text_locations), ci64(image_pos_ids), ci64(image_locations)
        if tloc is None and n_text:  # text_only context: identity placement (modelling_cross_attention.py:1072-1078)
            tloc = torch.arange(n_text, device=dev, dtype=torch.int64).repeat(bsz, 1).contiguous()
        a.text_proj, a.text_pos_table, a.text_pos_ids, a.text_locations = K._p(tp), K._p(tt), K._p(tpos), K._p(tloc)
        a.image_proj, a.image_pos_table, a.image_pos_ids, a.image_locations = K._p(ip), K._p(it), K._p(ipos), K._p(iloc)
        a.batch, a.n_text, a.n_image, a.row_width, a.n_tok = bsz, n_text, n_image, row_width, n_tok
        lpe32 = lw = lb = None
        if lpe is not None:
            lpe32 = lpe.to(F32).contiguous()
            lw, lb = w16(lpe_weight), f32(lpe_bias)
            a.lpe, a.lpe_k, a.lpe_weight, a.lpe_bias = K._p(lpe32), lpe32.shape[-1], K._p(lw), K._p(lb)
        bank = torch.empty((bsz, n_src, row_width), dtype=BF16, device=dev)
        mask = torch.empty((bsz, n_src * n_tok), dtype=torch.uint8, device=dev)
        a.bank, a.mask = K._p(bank), K._p(mask)
        K._req_cuda(bank, tp, ip)
        K.bank_pack_fwd(a)
        ctx.save_for_backward(tpos, tloc, ipos, iloc, lpe32, text_pos_table, image_pos_table, lpe_weight, lpe_bias)
        ctx.dims = (bsz, n_text, n_image, row_width, n_tok)
        ctx.dtypes = (text_proj.dtype if text_proj is not None else None,
                      image_proj.dtype if image_proj is not None else None)
        ctx.mark_non_differentiable(mask)
        h = row_width // n_tok
        return bank.reshape(bsz, n_src * n_tok, h), mask

    @staticmethod
    def backward(ctx, d_bank, _d_mask):
        tpos, tloc, ipos, iloc, lpe32, t_table, i_table, lpe_w, lpe_b = ctx.saved_tensors
        bsz, n_text, n_image, row_width, n_tok = ctx.dims
        dev = d_bank.device
        db = d_bank.reshape(bsz, n_text + n_image, row_width)
        db = (db if db.dtype == BF16 else db.to(BF16)).contiguous()
        a = K.BankBwdArgs()
        a.d_bank = K._p(db)
        a.text_pos_ids, a.text_locations, a.image_pos_ids, a.image_locations = K._p(tpos), K._p(tloc), K._p(ipos), K._p(iloc)
        a.batch, a.n_text, a.n_image, a.row_width = bsz, n_text, n_image, row_width
        need = ctx.needs_input_grad
        d_tp = d_ip = d_tt = d_it = d_lw = d_lb = None
        if n_text and need[0]:
            d_tp = torch.empty((bsz, n_text, row_width), dtype=BF16, device=dev)
            a.d_text_proj = K._p(d_tp)
        if n_image and need[4]:
            d_ip = torch.empty((bsz, n_image, row_width), dtype=BF16, device=dev)
            a.d_image_proj = K._p(d_ip)
        if t_table is not None and need[1]:
            d_tt = torch.zeros(t_table.shape, dtype=F32, device=dev)
            a.d_text_pos_table, a.text_pos_rows = K._p(d_tt), t_table.shape[0]
        if i_table is not None and need[5]:
            d_it = torch.zeros(i_table.shape, dtype=F32, device=dev)
            a.d_image_pos_table, a.image_pos_rows = K._p(d_it), i_table.shape[0]
        if lpe32 is not None:
            a.lpe, a.lpe_k = K._p(lpe32), lpe32.shape[-1]
            if need[9]:
                d_lw = torch.zeros(lpe_w.shape, dtype=F32, device=dev)
                a.d_lpe_weight = K._p(d_lw)
            if lpe_b is not None and need[10]:
                d_lb = torch.zeros(lpe_b.shape, dtype=F32, device=dev)
                a.d_lpe_bias = K._p(d_lb)
        K.bank_pack_bwd(a)
        t_dt, i_dt = ctx.dtypes
        if d_tp is not None and t_dt != BF16:
            d_tp = d_tp.to(t_dt)
        if d_ip is not None and i_dt != BF16:
            d_ip = d_ip.to(i_dt)

        def cast(g, p):
            return None if g is None else (g if p.dtype == F32 else g.to(p.dtype))
        return (d_tp, cast(d_tt, t_table) if d_tt is not None else None, None, None,
                d_ip, cast(d_it, i_table) if d_it is not None else None, None, None,
                None, cast(d_lw, lpe_w) if d_lw is not None else None,
                cast(d_lb, lpe_b) if d_lb is not None else None, None)


def bank_pack(text_proj, text_pos_table, text_pos_ids, text_locations,
              image_proj=None, image_pos_table=None, image_pos_ids=None, image_locations=None,
              lpe=None, lpe_weight=None, lpe_bias=None, n_tok=1):
    """text_proj [B,T,n_tok*H], image_proj [B,I,n_tok*H] -> bank [B,(T+I)*n_tok,H] bf16, mask uint8 [B,(T+I)*n_tok]."""
    return BankPackFn.apply(text_proj, text_pos_table, text_pos_ids, text_locations,
                            image_proj, image_pos_table, image_pos_ids, image_locations,
                            lpe, lpe_weight, lpe_bias, n_tok)


# --------------------------------------------------------------------------------------------- GCN
class GCNFn(torch.autograd.Function):
    """2-layer mean-aggregate GCN with a null root node (model/graph.py:17-31):
        Xr = [0; X];  H = relu([Xr, adj Xr] W1^T);  Y = ([H, adj H] W2^T)[:, 1:]
    The concat is never materialised for the GEMM: [X, AX] W^T = X Wa^T + (AX) Wb^T runs as the two operand
    pairs of one tcgen05 GEMM (Wa/Wb are column slices of W, addressed through the TMA descriptor)."""

    @staticmethod
    def forward(ctx, x, adj, w1, w2):
        bsz, n, din = x.shape
        nodes = n + 1
        dh, dout = w1.shape[0], w2.shape[0]
        xb = x.to(BF16).contiguous() if x.dtype != BF16 or not x.is_contiguous() else x
        adj32 = adj.to(F32).contiguous()
        c1 = torch.empty((bsz * nodes, 2 * din), dtype=BF16, device=x.device)
        K.gcn_concat_fwd(xb, adj32, c1, bsz, nodes, din, True)
        w1b, w2b = w16(w1), w16(w2)
        hid = _new(bsz * nodes, dh, xb)
        K.gemm(c1[:, :din], w1b[:, :din], hid, a1=c1[:, din:], b1=w1b[:, din:], relu=True)
        c2 = torch.empty((bsz * nodes, 2 * dh), dtype=BF16, device=x.device)
        K.gcn_concat_fwd(hid, adj32, c2, bsz, nodes, dh, False)
        y = _new(bsz * nodes, dout, xb)
        K.gemm(c2[:, :dh], w2b[:, :dh], y, a1=c2[:, dh:], b1=w2b[:, dh:])
        ctx.save_for_backward(adj32, c1, hid, c2, w1, w2)
        ctx.dims = (bsz, nodes, din, dh, dout)
        ctx.x_dtype = x.dtype
        return y.reshape(bsz, nodes, dout)[:, 1:, :]

    @staticmethod
    def backward(ctx, dy):
        adj32, c1, hid, c2, w1, w2 = ctx.saved_tensors
        bsz, nodes, din, dh, dout = ctx.dims
        dev = dy.device
        dyf = torch.zeros((bsz, nodes, dout), dtype=BF16, device=dev)
        dyf[:, 1:, :] = dy
        dy2 = dyf.reshape(bsz * nodes, dout)
        dx = dw1 = dw2 = None
        if ctx.needs_input_grad[3]:
            dw2 = _wgrad(dy2, c2, w2)
        dc2 = _new(bsz * nodes, 2 * dh, dy2)
        K.gemm(dy2, w16(w2), dc2, b_t=True)
        dhid = _new(bsz * nodes, dh, dy2)
        K.gcn_combine_bwd(dc2, adj32, hid, dhid, bsz, nodes, dh, False)          # ReLU mask applied here
        if ctx.needs_input_grad[2]:
            dw1 = _wgrad(dhid, c1, w1)
        if ctx.needs_input_grad[0]:
            dc1 = _new(bsz * nodes, 2 * din, dy2)
            K.gemm(dhid, w16(w1), dc1, b_t=True)
            dxb = torch.empty((bsz, nodes - 1, din), dtype=BF16, device=dev)
            K.gcn_combine_bwd(dc1, adj32, None, dxb, bsz, nodes, din, True)
            dx = dxb.to(ctx.x_dtype)
        return dx, None, dw1, dw2


def gcn(x, adj, w1, w2):
    return GCNFn.apply(x, adj, w1, w2)


# --------------------------------------------------------------------------------------------- Llama pieces
class RoPEFn(torch.autograd.Function):
    """Rotary position embedding on the q and k thirds of a fused Q|K|V projection [B,S,3H] (HF
    models/llama/modeling_llama.py: apply_rotary_pos_emb / rotate_half).  ``cos_sin`` fp32 [S, head_dim/2, 2].  The rotation
    is orthogonal, so the backward is the same kernel with the sign of sin flipped."""

    @staticmethod
    def forward(ctx, qkv, cos_sin, heads):
        b, s, h3 = qkv.shape
        d = h3 // 3 // heads
        out = _as2d(qkv).clone()
        K.rope_inplace(out, b * s, s, heads, d, 2, cos_sin, inverse=False)
        ctx.save_for_backward(cos_sin)
        ctx.cfg = (b, s, heads, d)
        return out.reshape(b, s, h3)

    @staticmethod
    def backward(ctx, d_out):
        (cos_sin,) = ctx.saved_tensors
        b, s, heads, d = ctx.cfg
        g = _as2d(d_out).clone()
        K.rope_inplace(g, b * s, s, heads, d, 2, cos_sin, inverse=True)
        return g.reshape(d_out.shape), None, None


def rope_qk(qkv, cos_sin, heads):
    return RoPEFn.apply(qkv, cos_sin, heads)


class SwiGLUFn(torch.autograd.Function):
    """h = silu(g) * u over gu = [g | u] (HF LlamaMLP.forward, models/llama/modeling_llama.py:155-165)."""

    @staticmethod
    def forward(ctx, gu):
        gu2 = _as2d(gu)
        f = gu2.shape[1] // 2
        h = _new(gu2.shape[0], f, gu2)
        K.swiglu_fwd(gu2, h)
        ctx.save_for_backward(gu2)
        ctx.shape = gu.shape
        return h.reshape(*gu.shape[:-1], f)

    @staticmethod
    def backward(ctx, dh):
        (gu2,) = ctx.saved_tensors
        dgu = torch.empty_like(gu2)
        K.swiglu_bwd(gu2, _as2d(dh), dgu)
        return dgu.reshape(ctx.shape)


def swiglu(gu):
    return SwiGLUFn.apply(gu)
