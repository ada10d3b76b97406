"""Frozen neighbor encoders on this package's kernels.

The reference runs HF ``RobertaModel`` / ``CLIPVisionModel`` over every neighbor
(model/modelling_cross_attention.py:992, :1018); they are frozen and run under ``no_grad``, so only an inference
forward is needed.  ``roberta_cls_hidden`` / ``clip_pooler_output`` read the weights of the HF modules in place
(same state-dict, nothing is copied or renamed) and run the same arithmetic through libmmgl_b200.so: fused QKV GEMM,
tcgen05 self-attention (bidirectional, key-padding mask), out-projection with the residual in the epilogue, LayerNorm,
FFN with GELU / quick-GELU in the epilogue.  Two things the library path cannot do:
  * only the [CLS] row of the last hidden state is consumed (TextPooler takes ``hidden[:, 0]``; CLIP pools token 0),
    so the LAST layer computes Q, the out-projection, the FFN and the LayerNorms for that single row per sequence;
  * no [B,1,S,S] additive masks, no head-split copies;
  * right-padded text batches are PACKED: the padding positions of every neighbor (data.py:457 pads all of them to
    max_input_length) are dropped before the first layer and the attention kernel runs on a variable-length batch
    (cu_seqlens) -- identical [CLS] outputs, about half the encoder work on WikiWeb2M-shaped batches.
Parity: tests/test_gpu_encoders.py compares with the HF modules' own forward on the same weights.
"""
from __future__ import annotations

import torch

from . import _capi as K
from .ops import BF16, F32, f32, fused_rows, w16


def _gemm(x, w, bias=None, act=0, residual=None):
    y = torch.empty((x.shape[0], w.shape[0]), dtype=BF16, device=x.device)
    K.gemm(x, w, y, bias=bias, relu=act, residual=residual)
    return y


def _ln(x, ln):
    y = torch.empty_like(x)
    mean = torch.empty(x.shape[0], dtype=F32, device=x.device)
    rstd = torch.empty_like(mean)
    K.layernorm_fwd(x, f32(ln.weight), f32(ln.bias), y, mean, rstd, float(ln.eps))
    return y


def _attention(qkv, key_mask, b, s, heads, d, scale):
    h = heads * d
    o = torch.empty((b * s, h), dtype=BF16, device=qkv.device)
    stats = torch.empty((b, heads, s, 2), dtype=F32, device=qkv.device)
    K.attn_fwd(qkv[:, :h], qkv[:, h:2 * h], qkv[:, 2 * h:], key_mask, None, o, stats, b, s, s, heads, d, scale, False)
    return o


def _attention_packed(qkv, cu, total, n, max_len, heads, d, scale):
    """bidirectional attention over a packed variable-length batch: sample i = rows [cu[i], cu[i+1])"""
    h = heads * d
    o = torch.empty((total, h), dtype=BF16, device=qkv.device)
    K.attn_fwd(qkv[:, :h], qkv[:, h:2 * h], qkv[:, 2 * h:], None, None, o, None, n, max_len, max_len, heads, d, scale, False,
               cu_seqlens=cu, total_tokens=total)
    return o


def _pack_plan(attention_mask):
    """Right-padded batches (every row a prefix of ones; data.py:457 pads each neighbor to max_input_length) can drop
    their padding positions altogether: padded keys are masked and padded queries are never read (only [CLS] is), so
    the encoder runs on the real tokens only.  One host read (lengths + a prefix-form flag).  Returns None when the mask
    is not prefix-form or nothing would be saved."""
    n, s = attention_mask.shape
    am = attention_mask != 0
    lens = am.sum(1)
    last = (am * torch.arange(1, s + 1, device=am.device)).amax(1)
    host = torch.cat((lens, ((lens == last) & (lens > 0)).all().reshape(1).to(lens.dtype))).cpu()
    if not bool(host[-1]):
        return None
    lens_h = host[:-1]
    total = int(lens_h.sum())
    if total * 10 > n * s * 9:          # < 10% padding: not worth the gathers
        return None
    cu_h = torch.zeros(n + 1, dtype=torch.int32)
    cu_h[1:] = torch.cumsum(lens_h, 0)
    cu = cu_h.to(am.device, non_blocking=True)
    idx = torch.nonzero_static(am.reshape(-1), size=total).squeeze(1)       # flat positions of the real tokens (no sync)
    return idx, cu, total, int(lens_h.max())


_ACT = {"relu": 1, "gelu": 2, "quick_gelu": 3}

# bench.py --no-packing control: run the text encoder on the padded [N, S] batch like the reference does
PACK_PADDING = True


def _supported(cfg) -> bool:
    d = cfg.hidden_size // cfg.num_attention_heads
    return d in (64, 128) and cfg.hidden_size % 8 == 0 and cfg.hidden_act in _ACT


@torch.no_grad()
def roberta_cls_hidden(model, input_ids, attention_mask, pack_padding=None, plan=None):
    """``model(input_ids, attention_mask).last_hidden_state[:, 0]`` of a HF RobertaModel, [N, hidden] bf16.
    (HF: models/roberta/modeling_roberta.py -- embeddings :70-150, layer :400-470; post-LN blocks.)"""
    cfg = model.config
    if not _supported(cfg):
        raise NotImplementedError(f"mmgl_b200 runs the frozen text encoder on its own kernels: head_dim must be 64 / 128 and "
                                  f"hidden_act one of {sorted(_ACT)} (got {cfg.hidden_size // cfg.num_attention_heads}, "
                                  f"{cfg.hidden_act!r}); there is no library fallback")
    if pack_padding is None:
        pack_padding = PACK_PADDING
    n, s = input_ids.shape
    emb = model.embeddings
    pad = emb.padding_idx
    not_pad = (input_ids != pad).to(torch.int64)
    pos = torch.cumsum(not_pad, dim=1) * not_pad + pad
    heads, h = cfg.num_attention_heads, cfg.hidden_size
    d = h // heads
    # plan: (token index, cu_seqlens, total, longest) made on the host (mmgl_b200.plan: no sync), False = do not pack,
    # None = derive it here from the device mask (one device->host read)
    if plan is None:
        plan = _pack_plan(attention_mask) if pack_padding else None
    elif plan is False:
        plan = None
    if plan is None:
        x = emb.word_embeddings(input_ids) + emb.position_embeddings(pos) + emb.token_type_embeddings.weight[0]
        x = x.reshape(n * s, -1)
        km = (attention_mask != 0).to(torch.uint8).contiguous()
    else:   # embed the real tokens only
        idx, cu, total, max_len = plan
        ids_p, pos_p = input_ids.reshape(-1).index_select(0, idx), pos.reshape(-1).index_select(0, idx)
        x = emb.word_embeddings(ids_p) + emb.position_embeddings(pos_p) + emb.token_type_embeddings.weight[0]
        cls_rows = cu[:-1].long()
    x = _ln(x.to(BF16).contiguous(), emb.LayerNorm)
    act = _ACT[cfg.hidden_act]
    layers = model.encoder.layer
    for li, layer in enumerate(layers):
        a = layer.attention
        w = fused_rows([a.self.query.weight, a.self.key.weight, a.self.value.weight], BF16)
        bias = fused_rows([a.self.query.bias, a.self.key.bias, a.self.value.bias], F32)
        qkv = _gemm(x, w, bias)
        if plan is None:
            ctx = _attention(qkv, km, n, s, heads, d, d ** -0.5)
        else:
            ctx = _attention_packed(qkv, cu, total, n, max_len, heads, d, d ** -0.5)
        if li == len(layers) - 1:   # only [CLS] rows are consumed downstream
            if plan is None:
                ctx = ctx.reshape(n, s, h)[:, 0].contiguous()
                x = x.reshape(n, s, h)[:, 0].contiguous()
            else:
                ctx, x = ctx.index_select(0, cls_rows), x.index_select(0, cls_rows)
        y = _ln(_gemm(ctx, w16(a.output.dense.weight), f32(a.output.dense.bias), residual=x), a.output.LayerNorm)
        f = _gemm(y, w16(layer.intermediate.dense.weight), f32(layer.intermediate.dense.bias), act=act)
        x = _ln(_gemm(f, w16(layer.output.dense.weight), f32(layer.output.dense.bias), residual=y), layer.output.LayerNorm)
    return x


_pad_cache = {}


def _padded_weight(param, w2d, pad):
    key = (id(param), param._version, pad)
    t = _pad_cache.get(key)
    if t is None:
        t = torch.nn.functional.pad(w2d, (0, pad)).contiguous()
        _pad_cache.clear()
        _pad_cache[key] = t
    return t


@torch.no_grad()
def clip_pooler_output(model, pixel_values):
    """``model(pixel_values).pooler_output`` of a HF CLIPVisionModel, [N, hidden] bf16.
    (HF: models/clip/modeling_clip.py -- CLIPVisionEmbeddings, pre-LN CLIPEncoderLayer, post_layernorm on token 0.)"""
    cfg = model.config
    vm = model.vision_model
    p = cfg.patch_size
    n, c, hh, ww = pixel_values.shape
    if not _supported(cfg) or hh % p or ww % p:
        raise NotImplementedError("mmgl_b200 runs the frozen vision tower on its own kernels: head_dim 64 / 128, image size a "
                                  "multiple of the patch size; there is no library fallback")
    gh, gw = hh // p, ww // p
    s = gh * gw + 1
    h, heads = cfg.hidden_size, cfg.num_attention_heads
    d = h // heads
    emb = vm.embeddings
    # stride-p convolution == GEMM over unfolded patches ([c, kh, kw] order of the conv weight)
    kdim = c * p * p
    patches = pixel_values.reshape(n, c, gh, p, gw, p).permute(0, 2, 4, 1, 3, 5).reshape(n * gh * gw, kdim).to(BF16)
    wpe = w16(emb.patch_embedding.weight).reshape(h, kdim)
    if kdim % 8:   # ViT-L/14: 3 * 14 * 14 = 588; the GEMM's TMA rows need a 16-byte pitch -> zero-pad K (cached weight copy)
        pad = 8 - kdim % 8
        patches = torch.nn.functional.pad(patches, (0, pad))
        wpe = _padded_weight(emb.patch_embedding.weight, wpe, pad)
    pe = _gemm(patches.contiguous(), wpe)
    x = torch.empty((n, s, h), dtype=BF16, device=pe.device)
    x[:, 0] = emb.class_embedding.to(BF16)
    x[:, 1:] = pe.reshape(n, gh * gw, h)
    x = x + emb.position_embedding.weight.to(BF16)[None, :s]
    x = _ln(x.reshape(n * s, h), vm.pre_layrnorm)
    act = _ACT[cfg.hidden_act]
    layers = vm.encoder.layers
    for li, layer in enumerate(layers):
        a = layer.self_attn
        y = _ln(x, layer.layer_norm1)
        w = fused_rows([a.q_proj.weight, a.k_proj.weight, a.v_proj.weight], BF16)
        bias = fused_rows([a.q_proj.bias, a.k_proj.bias, a.v_proj.bias], F32)
        ctx = _attention(_gemm(y, w, bias), None, n, s, heads, d, d ** -0.5)
        if li == len(layers) - 1:   # only the class token is pooled
            ctx = ctx.reshape(n, s, h)[:, 0].contiguous()
            x = x.reshape(n, s, h)[:, 0].contiguous()
        x = _gemm(ctx, w16(a.out_proj.weight), f32(a.out_proj.bias), residual=x)
        f = _gemm(_ln(x, layer.layer_norm2), w16(layer.mlp.fc1.weight), f32(layer.mlp.fc1.bias), act=act)
        x = _gemm(f, w16(layer.mlp.fc2.weight), f32(layer.mlp.fc2.bias), residual=x)
    return _ln(x, vm.post_layernorm)
