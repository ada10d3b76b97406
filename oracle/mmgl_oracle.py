"""CPU oracle for the MMGL neighbor-fusion hot path.  TEST INFRASTRUCTURE ONLY.

This file is a functional (no nn.Module) fp32/fp64 restatement, in plain torch
CPU ops, of the reference algorithm on the hot path named by BASELINE.json.  It
exists to CHECK the CUDA kernels; only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` leg may import it.
The product package (``mmgl_b200``) never imports anything under ``oracle/``.

Pinning: every function below is checked against outputs of the *real*
reference modules (imported from /root/reference in the build container by
``tests/golden/make_golden.py``; the resulting vectors are committed under
``tests/golden/*.pt``) in ``tests/test_oracle_golden.py``.  Two pieces cannot be
pinned that way and say so in their docstrings: ``lora_linear`` (the arithmetic
lives in the third-party ``peft`` package, unpinned in requirements.txt:8 and
absent from this image) and LPE *generation* (``utils.compute_LPE`` is undefined
in the reference) -- parity unpinned for those two.

All ``file:line`` citations are relative to /root/reference.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

# ----------------------------------------------------------------------------
# optional emulation of the CUDA path's bf16 STORAGE points
# ----------------------------------------------------------------------------
# The reference arithmetic is fp32 end to end.  The kernels keep fp32 accumulators but STORE activations in bf16
# (LayerNorm outputs, q / k / v, the attention output, FFN hidden, block outputs).  A ReLU whose pre-activation sits
# within one bf16 ulp of zero can then take the other branch -- a discontinuity, so gradients downstream differ from the
# pure-fp32 oracle by O(sqrt(flipped fraction)) however exact the kernels are.  Inside ``with bf16_storage():`` the oracle
# rounds the same tensors at the same points (straight-through: value rounded, gradient untouched), which makes the two
# sides take the same branches and lets the parity tests use a TIGHT tolerance for the kernel algebra; the comparison
# with the unrounded oracle / reference fixtures stays as the (looser) statement of bf16 accuracy.
_STORE = [False]


class bf16_storage:
    def __enter__(self):
        self.prev = _STORE[0]
        _STORE[0] = True
        return self

    def __exit__(self, *exc):
        _STORE[0] = self.prev
        return False


def _st(t: Tensor) -> Tensor:
    if not _STORE[0]:
        return t
    return t + (t.to(torch.bfloat16).to(t.dtype) - t).detach()


# ----------------------------------------------------------------------------
# masks / positions
# ----------------------------------------------------------------------------
def expand_mask(mask: Tensor, dtype: torch.dtype, tgt_len: int) -> Tensor:
    """[B,Nk] {0,1}/bool -> additive [B,1,tgt_len,Nk] (0 / finfo.min).

    model/modelling_cross_attention.py:68-79 (_expand_mask)."""
    b, src = mask.shape
    m = mask[:, None, None, :].expand(b, 1, tgt_len, src).to(dtype)
    inv = 1.0 - m
    return inv.masked_fill(inv.to(torch.bool), torch.finfo(dtype).min)


def causal_mask(bsz: int, tgt_len: int, dtype: torch.dtype) -> Tensor:
    """model/modelling_cross_attention.py:51-65 (_make_causal_mask), no past."""
    mask = torch.full((tgt_len, tgt_len), torch.finfo(dtype).min)
    cond = torch.arange(tgt_len)
    mask.masked_fill_(cond < (cond + 1).view(tgt_len, 1), 0)
    return mask.to(dtype)[None, None].expand(bsz, 1, tgt_len, tgt_len)


def decoder_attention_mask(attention_mask: Tensor, dtype: torch.dtype) -> Tensor:
    """causal + padding mask, model/modelling_cross_attention.py:455-476."""
    b, s = attention_mask.shape
    return expand_mask(attention_mask, dtype, s) + causal_mask(b, s, dtype)


def learned_positions(attention_mask: Tensor) -> Tensor:
    """position ids (before the +2 offset is applied by the embedding lookup).

    model/modelling_cross_attention.py:135-145: cumsum(mask)*mask - 1, then +offset(2)."""
    am = attention_mask.long()
    return (torch.cumsum(am, dim=1) * am).long() - 1 + 2


# ----------------------------------------------------------------------------
# a1: MPTAttention (cross and self branch)
# ----------------------------------------------------------------------------
def mpt_attention(
    x: Tensor,
    kv_src: Tensor,
    add_mask: Optional[Tensor],
    p: Dict[str, Tensor],
    num_heads: int,
    prefix: str = "self_attn.",
) -> Tensor:
    """Multi-head attention exactly as MPTAttention.forward computes it.

    model/modelling_cross_attention.py:179-275.  ``kv_src`` is the neighbor bank
    for the cross branch (:196-200) or ``x`` itself for the self branch (:201-204).
    ``add_mask`` is the additive [B,1,S,Nk] mask (0/finfo.min)."""
    b, s, h = x.shape
    d = h // num_heads
    scaling = d ** -0.5
    q = _st(F.linear(x, p[prefix + "q_proj.weight"], p.get(prefix + "q_proj.bias")) * scaling) # :194
    k = _st(F.linear(kv_src, p[prefix + "k_proj.weight"], p.get(prefix + "k_proj.bias")))    # :198
    v = _st(F.linear(kv_src, p[prefix + "v_proj.weight"], p.get(prefix + "v_proj.bias")))    # :199
    nk = kv_src.shape[1]

    def shape(t, n):                                                                           # :176-177
        return t.view(b, n, num_heads, d).transpose(1, 2).contiguous().view(b * num_heads, n, d)

    q, k, v = shape(q, s), shape(k, nk), shape(v, nk)
    w = torch.bmm(q, k.transpose(1, 2))                                                        # :212
    if add_mask is not None:                                                                   # :220-229
        w = w.view(b, num_heads, s, nk) + add_mask
        w = torch.max(w, torch.tensor(torch.finfo(w.dtype).min))
        w = w.view(b * num_heads, s, nk)
    w = F.softmax(w, dim=-1)                                                                   # :235
    o = torch.bmm(w, v)                                                                        # :258
    o = _st(o.view(b, num_heads, s, d).transpose(1, 2).reshape(b, s, h))                       # :266-271
    return F.linear(o, p[prefix + "out_proj.weight"], p.get(prefix + "out_proj.bias"))        # :273


def xattn_core(q: Tensor, k: Tensor, v: Tensor, mask: Tensor, num_heads: int) -> Tuple[Tensor, Tensor]:
    """The fused-kernel slice of mpt_attention: head-interleaved Q[B,S,H] (already
    scaled), K,V[B,Nk,H], byte mask[B,Nk] -> O[B,S,H] and log-sum-exp [B,nh,S].

    Same arithmetic as model/modelling_cross_attention.py:206-271."""
    b, s, h = q.shape
    nk = k.shape[1]
    d = h // num_heads
    qh = q.view(b, s, num_heads, d).transpose(1, 2)
    kh = k.view(b, nk, num_heads, d).transpose(1, 2)
    vh = v.view(b, nk, num_heads, d).transpose(1, 2)
    w = qh @ kh.transpose(-1, -2)
    add = expand_mask(mask, w.dtype, s)
    w = torch.max(w + add, torch.tensor(torch.finfo(w.dtype).min))
    lse = torch.logsumexp(w, dim=-1)
    o = F.softmax(w, dim=-1) @ vh
    return o.transpose(1, 2).reshape(b, s, h), lse


def attention_core(q: Tensor, k: Tensor, v: Tensor, num_heads: int, scale: float = 1.0,
                   key_mask: Optional[Tensor] = None, causal: bool = False, position_bias: Optional[Tensor] = None,
                   prob_multiplier: Optional[Tensor] = None) -> Tensor:
    """General attention core of the language models on the concat path: Q[B,Sq,H], K,V[B,Sk,H] (heads interleaved),
    byte key mask [B,Sk], additive position bias [1|B, nh, Sq, Sk], optional dropout multiplier [B,nh,Sq,Sk] applied to
    the probabilities.  Restates HF T5Attention (HF: models/t5/modeling_t5.py:253-345 -- scores = q k^T with NO scaling,
    += position_bias (+ mask), softmax in fp32, dropout, @ v) and the OPT / MPT self branch
    (model/modelling_cross_attention.py:201-275: scaled q, additive finfo.min masks, clamp)."""
    b, sq, h = q.shape
    sk = k.shape[1]
    d = h // num_heads
    qh = q.view(b, sq, num_heads, d).transpose(1, 2)
    kh = k.view(b, sk, num_heads, d).transpose(1, 2)
    vh = v.view(b, sk, num_heads, d).transpose(1, 2)
    w = (qh * scale) @ kh.transpose(-1, -2)
    if position_bias is not None:
        w = w + position_bias
    add = torch.zeros(b, 1, sq, sk, dtype=w.dtype)
    if key_mask is not None:
        add = add + expand_mask(key_mask, w.dtype, sq)
    if causal:
        add = add + causal_mask(b, sq, w.dtype)
    w = torch.max(w + add, torch.tensor(torch.finfo(w.dtype).min))
    pr = F.softmax(w, dim=-1)
    if prob_multiplier is not None:
        pr = pr * prob_multiplier
    return (pr @ vh).transpose(1, 2).reshape(b, sq, h)


def t5_relative_position_bucket(relative_position: Tensor, bidirectional: bool, num_buckets: int = 32,
                                max_distance: int = 128) -> Tensor:
    """HF: models/t5/modeling_t5.py T5Attention._relative_position_bucket (relative_position = key - query)."""
    rp = relative_position
    buckets = torch.zeros_like(rp)
    if bidirectional:
        num_buckets //= 2
        buckets = buckets + (rp > 0).long() * num_buckets
        rp = rp.abs()
    else:
        rp = -torch.min(rp, torch.zeros_like(rp))
    max_exact = num_buckets // 2
    is_small = rp < max_exact
    large = max_exact + (torch.log(rp.float() / max_exact) / math.log(max_distance / max_exact)
                         * (num_buckets - max_exact)).long()
    large = torch.min(large, torch.full_like(large, num_buckets - 1))
    return buckets + torch.where(is_small, rp, large)


def t5_position_bias(table: Tensor, sq: int, sk: int, bidirectional: bool, num_buckets: int = 32,
                     max_distance: int = 128) -> Tensor:
    """[1, nh, sq, sk] additive bias from the [num_buckets, nh] embedding table (T5Attention.compute_bias)."""
    ctx = torch.arange(sq)[:, None]
    mem = torch.arange(sk)[None, :]
    bucket = t5_relative_position_bucket(mem - ctx, bidirectional, num_buckets, max_distance)
    return table[bucket].permute(2, 0, 1)[None]


# ----------------------------------------------------------------------------
# dropout (nn.functional.dropout, model/modelling_cross_attention.py:332, :356)
# ----------------------------------------------------------------------------
def dropout_multiplier(seed: int, p: float, m: int, n: int) -> Tensor:
    """[m,n] fp32 multiplier (0 or 1/(1-p')) of the CUDA path's counter-based dropout, restated with numpy integers.

    torch's Philox stream cannot be reproduced by another implementation, so parity for dropout is (i) this mask
    restatement, bit-exact, and (ii) the reference semantics y = x * keep / (1-p) given the same keep mask.
    Element (r, c): g = r*ceil(n/8) + c//8; word = splitmix64(seed ^ (2g + (c%8)//4)); lane = (c%8)%4;
    keep iff ((word >> 16*lane) & 0xFFFF) >= round(p*65536); p' = round(p*65536)/65536."""
    import numpy as np
    thresh = int(p * 65536.0 + 0.5)
    if thresh == 0:
        return torch.ones(m, n)
    groups = (n + 7) // 8
    r = np.arange(m, dtype=np.uint64)[:, None]
    c = np.arange(n, dtype=np.uint64)[None, :]
    g = r * np.uint64(groups) + c // np.uint64(8)
    z = np.uint64(seed & 0xFFFFFFFFFFFFFFFF) ^ (np.uint64(2) * g + (c % np.uint64(8)) // np.uint64(4))
    with np.errstate(over="ignore"):
        z = z + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    lane = (c % np.uint64(8)) % np.uint64(4)
    bits = (z >> (np.uint64(16) * lane)) & np.uint64(0xFFFF)
    keep = bits >= np.uint64(thresh)
    scale = 65536.0 / (65536.0 - thresh)
    return torch.from_numpy(keep.astype(np.float32) * np.float32(scale))


# ----------------------------------------------------------------------------
# a2: MPTDecoderLayer
# ----------------------------------------------------------------------------
def mpt_decoder_layer(
    x: Tensor,
    p: Dict[str, Tensor],
    num_heads: int,
    *,
    cross_attention: bool,
    add_mask: Optional[Tensor] = None,
    bank: Optional[Tensor] = None,
    bank_add_mask: Optional[Tensor] = None,
    do_layer_norm_before: bool = True,
    flamingo: bool = True,
    eps: float = 1e-5,
    drop1: Optional[Tensor] = None,
    drop2: Optional[Tensor] = None,
) -> Tensor:
    """model/modelling_cross_attention.py:304-375.  Dropout (:332, :356) is the identity (eval) unless the
    multipliers ``drop1``/``drop2`` ([B*S,H], 0 or 1/(1-p)) are supplied -- see dropout_multiplier.

    cross_attention=True, flamingo=True: tanh-gated residuals with scalars
    ``gating1``/``gating2`` (:298-302, :334-335, :358-359)."""
    h = x.shape[-1]
    residual = x
    hs = x
    if do_layer_norm_before:                                                                   # :319-320
        hs = _st(F.layer_norm(hs, (h,), p["self_attn_layer_norm.weight"], p["self_attn_layer_norm.bias"], eps))
    if cross_attention:
        hs = mpt_attention(hs, bank, bank_add_mask, p, num_heads)                              # :323-331
    else:
        hs = mpt_attention(hs, hs, add_mask, p, num_heads)
    if drop1 is not None:                                                                      # :332
        hs = hs * drop1.view(hs.shape)
    if cross_attention and flamingo:                                                           # :334-337
        hs = residual + torch.tanh(p["gating1"]) * hs
    else:
        hs = residual + hs
    hs = _st(hs)
    if not do_layer_norm_before:                                                               # :340-341
        hs = _st(F.layer_norm(hs, (h,), p["self_attn_layer_norm.weight"], p["self_attn_layer_norm.bias"], eps))
    shape = hs.shape
    hs = hs.reshape(-1, h)
    residual = hs
    if do_layer_norm_before:                                                                   # :349-350
        hs = _st(F.layer_norm(hs, (h,), p["final_layer_norm.weight"], p["final_layer_norm.bias"], eps))
    hs = _st(F.relu(F.linear(hs, p["fc1.weight"], p.get("fc1.bias"))))                         # :352-353 (OPT: relu)
    hs = F.linear(hs, p["fc2.weight"], p.get("fc2.bias"))                                      # :355
    if drop2 is not None:                                                                      # :356
        hs = hs * drop2.view(hs.shape)
    if cross_attention and flamingo:                                                           # :358-361
        hs = (residual + torch.tanh(p["gating2"]) * hs).view(shape)
    else:
        hs = (residual + hs).view(shape)
    hs = _st(hs)
    if not do_layer_norm_before:                                                               # :364-365
        hs = _st(F.layer_norm(hs, (h,), p["final_layer_norm.weight"], p["final_layer_norm.bias"], eps))
    return hs


def sub(p: Dict[str, Tensor], prefix: str) -> Dict[str, Tensor]:
    n = len(prefix)
    return {k[n:]: v for k, v in p.items() if k.startswith(prefix)}


# ----------------------------------------------------------------------------
# a4 + a12: decoder interleave loop and the LM head / loss
# ----------------------------------------------------------------------------
def mpt_causal_lm(
    p: Dict[str, Tensor],
    cfg: Dict,
    input_ids: Tensor,
    attention_mask: Tensor,
    labels: Optional[Tensor],
    bank: Optional[Tensor],
    bank_mask: Optional[Tensor],
) -> Tuple[Optional[Tensor], Tensor]:
    """MPTForCausalLM.forward: model/modelling_cross_attention.py:478-653 (decoder,
    interleave rule :613-625) and :826-836 (lm_head, shifted CE over all positions).

    ``p`` uses the reference state-dict keys (``model.decoder.*``, ``lm_head.weight``).
    cfg: num_heads, num_layers, neighbor_layer_wise, do_layer_norm_before, vocab."""
    nh = cfg["num_heads"]
    pre_ln = cfg.get("do_layer_norm_before", True)
    dec = "model.decoder."
    x = F.embedding(input_ids, p[dec + "embed_tokens.weight"])                                # :527
    dtype = x.dtype
    b, s = input_ids.shape
    add_mask = decoder_attention_mask(attention_mask, dtype)                                   # :542-544
    bank_add = expand_mask(bank_mask, dtype, s) if bank_mask is not None else None             # :545-546
    pos = F.embedding(learned_positions(attention_mask), p[dec + "embed_positions.weight"])   # :548
    if dec + "project_in.weight" in p:
        x = F.linear(x, p[dec + "project_in.weight"])                                          # :550-551
    x = _st(x + pos)                                                                           # :553
    nlw = cfg["neighbor_layer_wise"]
    for idx in range(cfg["num_layers"]):                                                       # :576
        x = mpt_decoder_layer(x, sub(p, f"{dec}layers.{idx}."), nh, cross_attention=False,
                              add_mask=add_mask, do_layer_norm_before=pre_ln)
        if bank is not None and (idx + 1) % nlw == 0:                                          # :613
            k = (idx + 1) // nlw - 1
            x = mpt_decoder_layer(x, sub(p, f"{dec}neighbor_layers.{k}."), nh, cross_attention=True,
                                  bank=bank, bank_add_mask=bank_add, do_layer_norm_before=pre_ln,
                                  flamingo=cfg.get("flamingo", True))
    if dec + "final_layer_norm.weight" in p:                                                   # :635-636
        x = _st(F.layer_norm(x, (x.shape[-1],), p[dec + "final_layer_norm.weight"], p[dec + "final_layer_norm.bias"]))
    if dec + "project_out.weight" in p:
        x = F.linear(x, p[dec + "project_out.weight"])                                         # :638-639
    logits = F.linear(x, p["lm_head.weight"])                                                  # :826
    loss = None
    if labels is not None:                                                                     # :828-836
        sl = logits[..., :-1, :].contiguous()
        tl = labels[..., 1:].contiguous()
        loss = F.cross_entropy(sl.view(-1, sl.shape[-1]), tl.view(-1))
    return loss, logits


# ----------------------------------------------------------------------------
# a5: neighbor projections, a6: bank packing
# ----------------------------------------------------------------------------
def text_pooler(hidden: Tensor, w: Tensor, b: Tensor) -> Tensor:
    """tanh(Linear(h[:,0])): model/modelling_cross_attention.py:888-893."""
    return torch.tanh(F.linear(hidden[:, 0], w, b))


def neighbor_projection(
    pooled: Tensor, w: Tensor, b: Tensor, pos_table: Optional[Tensor], pos_ids: Optional[Tensor], n_tokens: int
) -> Tensor:
    """pooled [B,N,E] -> [B,N,n_tokens,H]: Linear(E -> n_tokens*H) + pos-emb gather add.

    model/modelling_cross_attention.py:997-1004 (text) / :1020-1027 (visual);
    twins at model/modelling_self_attention.py:170-177, :193-200."""
    bsz, n, e = pooled.shape
    y = F.linear(pooled.reshape(-1, e), w, b)
    if pos_table is not None and pos_ids is not None:
        y = y + F.embedding(pos_ids.reshape(-1), pos_table)
    return y.reshape(bsz, n, n_tokens, -1)


def pack_bank(
    text_embeds: Tensor, text_pos_ids: Tensor, text_locations: Tensor,
    visual_embeds: Tensor, visual_pos_ids: Tensor, image_locations: Tensor,
) -> Tuple[Tensor, Tensor]:
    """Interleave text/image neighbor embeddings by ``*_locations`` into
    bank [B,(T+I)*n_tok,H] and bool mask [B,(T+I)*n_tok] (= pos_id > 0).

    model/modelling_cross_attention.py:1080-1104 (twin: modelling_self_attention.py:284-308)."""
    b, t, n_tok, h = text_embeds.shape
    i = visual_embeds.shape[1]
    bidx = torch.arange(b)[:, None]
    bank = torch.zeros((b, t + i, n_tok, h), dtype=text_embeds.dtype)
    bank[bidx, text_locations] = text_embeds
    bank[bidx, image_locations] = visual_embeds
    mask = torch.zeros((b, t + i, n_tok), dtype=torch.bool)
    mask[bidx, text_locations] = (text_pos_ids > 0).unsqueeze(-1).expand(-1, -1, n_tok)
    mask[bidx, image_locations] = (visual_pos_ids > 0).unsqueeze(-1).expand(-1, -1, n_tok)
    return bank.reshape(b, -1, h), mask.reshape(b, -1)


def pack_bank_text_only(text_embeds: Tensor, text_pos_ids: Tensor) -> Tuple[Tensor, Tensor]:
    """context == text_only: model/modelling_cross_attention.py:1072-1078."""
    b, t, n_tok, h = text_embeds.shape
    mask = torch.repeat_interleave(text_pos_ids > 0, repeats=n_tok, dim=1)
    return text_embeds.reshape(b, t * n_tok, h), mask


# ----------------------------------------------------------------------------
# a9: Laplacian-PE projection, a10: GCN
# ----------------------------------------------------------------------------
def lpe_add(bank: Tensor, lpe: Tensor, w: Tensor, b: Tensor, n_tok: int) -> Tensor:
    """bank += Linear(k -> n_tok*H)(lpe)[:, 1:]  (root row dropped).

    model/modelling_self_attention.py:311-315."""
    bsz, nk, h = bank.shape
    n = nk // n_tok
    e = F.linear(lpe, w, b).reshape(bsz, n + 1, n_tok, h)
    return bank + e[:, 1:].reshape(bsz, -1, h)


def gcn_forward(x: Tensor, adj: Tensor, w1: Tensor, w2: Tensor) -> Tensor:
    """2-layer mean-aggregate GCN with a null root node: model/graph.py:17-31.

    x [B,N,Din], adj [B,N+1,N+1]; w1 [Dh, 2*Din], w2 [Dout, 2*Dh]  ->  [B,N,Dout]."""
    root = torch.zeros((x.shape[0], 1, x.shape[2]), dtype=x.dtype)
    x = torch.cat((root, x), dim=1)
    agg = torch.bmm(adj, x)
    x = F.relu(F.linear(torch.cat((x, agg), dim=-1), w1))
    agg = torch.bmm(adj, x)
    x = F.linear(torch.cat((x, agg), dim=-1), w2)
    return x[:, 1:, :]


def gnn_add(bank: Tensor, graph: Tensor, w1: Tensor, w2: Tensor, n_tok: int) -> Tensor:
    """bank viewed [B,N,n_tok*H] + GCN(bank, graph): model/modelling_self_attention.py:316-320."""
    bsz, nk, h = bank.shape
    n = nk // n_tok
    flat = bank.reshape(bsz, n, n_tok * h)
    return (flat + gcn_forward(flat, graph, w1, w2)).reshape(bsz, nk, h)


# ----------------------------------------------------------------------------
# a7: concat path (wrapper logic only; the LM itself is HF library code)
# ----------------------------------------------------------------------------
def concat_neighbors(
    input_embs: Tensor, attention_mask: Tensor, labels: Tensor, bank: Tensor, bank_mask: Tensor, decoder_only: bool
) -> Tuple[Tensor, Tensor, Tensor]:
    """model/modelling_self_attention.py:323-330: neighbors are appended AFTER the
    input tokens, mask concatenated (float), labels padded with -100."""
    embs = torch.cat((input_embs, bank), dim=1)
    mask = torch.cat((attention_mask, bank_mask.to(attention_mask.dtype)), dim=1)
    if decoder_only:
        pad = -100 * torch.ones((bank.shape[0], bank.shape[1]), dtype=labels.dtype)
        labels = torch.cat((labels, pad), dim=1)
    return embs, mask, labels


# ----------------------------------------------------------------------------
# a8: LoRA  (PARITY UNPINNED: peft is not in this image; published algorithm restated)
# ----------------------------------------------------------------------------
def lora_linear(x: Tensor, w: Tensor, bias: Optional[Tensor], a: Tensor, b: Tensor, alpha: float, r: int) -> Tensor:
    """y = x W^T + bias + (alpha / r) * (x A^T) B^T   (dropout = 0).

    LoRA as published (Hu et al. 2021) and as configured at
    model/modelling_self_attention.py:80-87 (r=lora_r, lora_alpha, bias="none"); init rule
    model/modelling_cross_attention.py:719-724 (A kaiming-uniform a=sqrt(5), B zeros).
    peft (requirements.txt:8, unpinned, absent) -> parity unpinned."""
    return F.linear(x, w, bias) + (alpha / r) * F.linear(F.linear(x, a), b)


def lora_init(r: int, in_features: int, out_features: int, gen: torch.Generator) -> Tuple[Tensor, Tensor]:
    a = torch.empty(r, in_features)
    bound = math.sqrt(6.0 / ((1 + 5.0) * in_features))  # kaiming_uniform_(a=sqrt(5))
    a.uniform_(-bound, bound, generator=gen)
    return a, torch.zeros(out_features, r)


# ----------------------------------------------------------------------------
# whole wrapper (a5+a6+a4+a12) given pooled encoder features
# ----------------------------------------------------------------------------
def cross_attention_model_from_pooled(
    p: Dict[str, Tensor], cfg: Dict, batch: Dict[str, Tensor], text_pooled: Tensor, visual_pooled: Tensor
) -> Tuple[Tensor, Tensor]:
    """CrossAttentionModel.forward (model/modelling_cross_attention.py:1038-1114) downstream of the
    frozen encoders: ``text_pooled`` = text_pooler(RoBERTa(...)) [B,T,E], ``visual_pooled`` =
    CLIPVision(...).pooler_output [B,I,E].  Optional ``lpe``/``graph`` add (SURVEY D9 extension)."""
    n_tok = cfg["n_tokens"]
    te = neighbor_projection(text_pooled, p["text_embeddings.weight"], p["text_embeddings.bias"],
                             p["text_position_embeddings.weight"], batch["neighbor_pos_ids"], n_tok)
    ve = neighbor_projection(visual_pooled, p["visual_embeddings.weight"], p["visual_embeddings.bias"],
                             p["visual_position_embeddings.weight"], batch["neighbor_images_pos_ids"], n_tok)
    bank, mask = pack_bank(te, batch["neighbor_pos_ids"], batch["text_locations"],
                           ve, batch["neighbor_images_pos_ids"], batch["image_locations"])
    if "lpe" in batch and "lpe_embeddings.weight" in p:
        bank = lpe_add(bank, batch["lpe"], p["lpe_embeddings.weight"], p["lpe_embeddings.bias"], n_tok)
    if "graph" in batch and "gnn.w1.weight" in p:
        bank = gnn_add(bank, batch["graph"], p["gnn.w1.weight"], p["gnn.w2.weight"], n_tok)
    bank = _st(bank)
    return mpt_causal_lm(sub(p, "lm."), cfg, batch["input_ids"], batch["attention_mask"], batch["labels"], bank, mask)
