"""The language model of the concatenated-embedding path on this package's kernels.

The reference's SelfAttentionModel hands ``inputs_embeds`` (token embeddings ++ neighbor bank) to a HuggingFace T5 or
OPT model (model/modelling_self_attention.py:261, :280, :332), optionally wrapped by peft LoRA (:79-87).  The functions
here run that SAME arithmetic -- the layer stacks of HF ``T5ForConditionalGeneration`` / ``OPTForCausalLM`` -- through
libmmgl_b200.so, reading the weights of the HF modules in place (nothing is copied or renamed, so state-dict keys and
checkpoints stay those of the reference):

  * projections, FFN, lm_head: the tcgen05 GEMM (bias / ReLU / dropout / residual in the epilogue); LoRA'd q / v through
    ``ops.lora_linear`` (the rank-r product accumulates into the same TMEM tile);
  * attention: ``ops.attention`` -- T5: unscaled scores, bucketed relative-position bias (bidirectional in the encoder,
    causal in the decoder), dropout on the probabilities, decoder cross-attention over the encoder output with the
    encoder's key padding; OPT: scaled, causal + key padding;
  * T5LayerNorm -> ``ops.rms_norm``; nn.LayerNorm -> ``ops.layer_norm``; loss -> ``ops.cross_entropy``.

HF arithmetic restated (HF: models/t5/modeling_t5.py -- T5Stack.forward :637-790, T5Block :424-497, T5Attention
:253-345, T5LayerFF :146-150, T5ForConditionalGeneration.forward :992-1130; models/opt/modeling_opt.py -- OPTDecoder
:320-400, OPTDecoderLayer :202-254).  Parity: tests/test_gpu_lm.py compares loss, logits and gradients with the HF
modules' own fp32 forward / backward on the same weights.

``supports(lm)`` says whether a model can run here; SelfAttentionModel raises when it cannot (head_dim not in {64, 128},
gated-GELU T5 variants, LayerDrop): there is no HF / eager fallback.  Prefix tuning is implemented for OPT
(``opt_forward(prefix_kv=...)``: the causal kernel takes keys = virtual tokens + own positions) and for T5
(``t5_forward(prefix_kv=...)``: the decoder's self-attention, the effective behaviour of peft + the reference-era HF).
"""
from __future__ import annotations

import math

import torch

from . import ops

BF16 = torch.bfloat16


class LMOutput(dict):
    """loss / logits with attribute, key and index access (the fields run_generation.py reads, :466-474)."""

    def __init__(self, loss=None, logits=None):
        super().__init__(loss=loss, logits=logits)
        self.loss, self.logits = loss, logits

    def __getitem__(self, k):
        if isinstance(k, int):
            return [v for v in (self.loss, self.logits) if v is not None][k]
        return super().__getitem__(k)


def _proj(lin, x, alpha=1.0):
    """nn.Linear or LoRALinear (self_attention.LoRALinear: base_layer + lora_A / lora_B) through the GEMM kernel."""
    base = getattr(lin, "base_layer", None)
    if base is None:
        return ops.linear(x, lin.weight, lin.bias, alpha=alpha)
    if alpha != 1.0:
        raise NotImplementedError("scaled LoRA projection")
    return lin(x)


def _weight(lin):
    return getattr(lin, "base_layer", lin).weight


# ------------------------------------------------------------------------------------------------- T5
def _t5_bucket(rel, bidirectional, num_buckets, max_distance):
    """T5Attention._relative_position_bucket (HF t5 :189-233); rel = key position - query position."""
    buckets = torch.zeros_like(rel)
    if bidirectional:
        num_buckets //= 2
        buckets = buckets + (rel > 0).long() * num_buckets
        rel = rel.abs()
    else:
        rel = -torch.min(rel, torch.zeros_like(rel))
    max_exact = num_buckets // 2
    is_small = rel < max_exact
    large = max_exact + (torch.log(rel.float() / max_exact) / math.log(max_distance / max_exact)
                         * (num_buckets - max_exact)).long()
    large = torch.min(large, torch.full_like(large, num_buckets - 1))
    return buckets + torch.where(is_small, rel, large)


def t5_rel_bias(attn, sq, sk, shift=0):
    """The position bias of T5Attention.compute_bias (:236-251) in the kernel's compact form: it depends on
    key - query only, so [heads, sq + sk - 1] fp32 (entry d + sq - 1 = bias at distance d) replaces [1,nh,sq,sk].
    ``shift``: the queries sit ``shift`` positions after key 0 (virtual-token K / V of prefix tuning in front of the
    decoder's own keys: HF computes the bias for the full key length and keeps the last ``sq`` query rows)."""
    dev = attn.relative_attention_bias.weight.device
    rel = torch.arange(-(sq - 1), sk, device=dev) - shift
    bucket = _t5_bucket(rel, not attn.is_decoder, attn.relative_attention_num_buckets, attn.relative_attention_max_distance)
    w = attn.relative_attention_bias.weight
    if not w.requires_grad:
        w = w.detach()
    return w.float()[bucket].t().contiguous()     # differentiable when the table is trainable (peft "none")


def _t5_attention(attn, x, kv, key_mask, rel_bias, causal, p_drop, prefix=None):
    """T5Attention.forward (:253-345) for training: q / k / v without bias or scaling, o projection.  ``prefix`` = (k, v)
    [n_virtual, inner_dim] of prefix tuning: the virtual tokens' keys / values go in front of the projected ones, exactly
    what HF does with a pre-filled cache (``curr_past_key_values.update`` concatenates along the key axis)."""
    q = _proj(attn.q, x)
    k = _proj(attn.k, kv)
    v = _proj(attn.v, kv)
    if prefix is not None:
        b = k.shape[0]
        k = torch.cat((prefix[0].to(BF16)[None].expand(b, -1, -1), k), dim=1)
        v = torch.cat((prefix[1].to(BF16)[None].expand(b, -1, -1), v), dim=1)
    return ops.attention(q, k, v, key_mask=key_mask, rel_bias=rel_bias, heads=attn.n_heads, causal=causal, scale=1.0,
                         dropout_p=p_drop)


def _t5_stack(stack, h, key_mask, enc=None, enc_mask=None, p=0.0, prefix_kv=None):
    """T5Stack.forward (:637-790): dropout(embeds) -> blocks -> final RMSNorm -> dropout.  ``prefix_kv`` (decoder only)
    [n_virtual, layers, 2, inner_dim]: the virtual K / V of prefix tuning go in front of every decoder layer's SELF-attention
    keys.  peft's get_prompt also hands the same tensors over as cross-attention "past" K / V, but the transformers of the
    reference's era (4.2x, T5Attention.project: "checking that the sequence_length of the past_key_value is the same as the
    provided key_value_states to support prefix tuning") discards them there and projects the encoder output as usual --
    while transformers 5.x would attend to the 20 virtual tokens INSTEAD of the encoder (a pre-filled cross-attention cache
    counts as final).  The era's behaviour is the one restated: cross-attention is untouched."""
    cfg = stack.config
    eps = cfg.layer_norm_epsilon
    b, s = h.shape[:2]
    h = ops.dropout(h, p)
    self0 = stack.block[0].layer[0].SelfAttention
    n_pre = 0 if prefix_kv is None else prefix_kv.shape[0]
    bias = t5_rel_bias(self0, s, s + n_pre, shift=n_pre)     # shared by every layer of the stack (:768-774)
    if n_pre:
        ones = torch.ones((b, n_pre), dtype=torch.uint8, device=h.device)
        self_mask = ones if key_mask is None else torch.cat((ones, key_mask), dim=1)
        if key_mask is None:
            self_mask = torch.cat((ones, torch.ones((b, s), dtype=torch.uint8, device=h.device)), dim=1)
    else:
        self_mask = key_mask
    for li, block in enumerate(stack.block):
        pre = None if not n_pre else (prefix_kv[:, li, 0], prefix_kv[:, li, 1])
        sa = block.layer[0]
        xn, h = ops.rms_norm_fork(h, sa.layer_norm.weight, eps)     # (norm, residual): one backward kernel for both
        a = _t5_attention(sa.SelfAttention, xn, xn, self_mask, bias, stack.is_decoder, p, prefix=pre)
        h = ops.linear(a, _weight(sa.SelfAttention.o), None, residual=h, dropout_p=p)
        if stack.is_decoder:
            ca = block.layer[1]
            xn, h = ops.rms_norm_fork(h, ca.layer_norm.weight, eps)
            a = _t5_attention(ca.EncDecAttention, xn, enc, enc_mask, None, False, p)
            h = ops.linear(a, _weight(ca.EncDecAttention.o), None, residual=h, dropout_p=p)
        ff = block.layer[-1]
        dense = ff.DenseReluDense
        xn, h = ops.rms_norm_fork(h, ff.layer_norm.weight, eps)
        h = ops.mlp(xn, dense.wi.weight, None, dense.wo.weight, None, residual=h, dropout_p=p, hidden_dropout_p=p)
    return ops.dropout(ops.rms_norm(h, stack.final_layer_norm.weight, eps), p)


def _t5_supported(lm) -> bool:
    cfg = lm.config
    if cfg.d_kv not in (64, 128) or cfg.d_model % 8 or cfg.d_ff % 8 or getattr(cfg, "is_gated_act", False):
        return False
    if getattr(cfg, "dense_act_fn", "relu") != "relu":
        return False
    return True


def _shift_right(lm, labels):
    """T5PreTrainedModel._shift_right (:595-614)."""
    cfg = lm.config
    dec = labels.new_zeros(labels.shape)
    dec[..., 1:] = labels[..., :-1].clone()
    dec[..., 0] = cfg.decoder_start_token_id
    return dec.masked_fill(dec == -100, cfg.pad_token_id)


def t5_forward(lm, input_ids=None, attention_mask=None, inputs_embeds=None, labels=None, prefix_kv=None):
    """T5ForConditionalGeneration.forward (:992-1130) for training: returns loss (mean CE, ignore_index -100) and
    logits [B, S_dec, V].  ``prefix_kv`` [n_virtual, decoder layers, 2, inner_dim]: prefix tuning
    (model/modelling_self_attention.py:88-92; peft hands the virtual K / V to the DECODER as past_key_values and leaves the
    encoder alone; they act on the decoder's self-attention, see _t5_stack; semantics restated, peft absent: parity unpinned
    against peft itself, pinned against HF's own forward fed the same tensors as the self-attention half of an
    EncoderDecoderCache)."""
    if labels is None:
        raise ValueError("t5_forward is the training forward: labels are required")
    p = lm.config.dropout_rate if lm.training else 0.0
    if inputs_embeds is None:
        inputs_embeds = lm.shared(input_ids)
    b, s_enc = inputs_embeds.shape[:2]
    if attention_mask is None:
        attention_mask = torch.ones(b, s_enc, dtype=torch.long, device=inputs_embeds.device)
    enc_mask = (attention_mask != 0).to(torch.uint8).contiguous()
    enc = _t5_stack(lm.encoder, inputs_embeds.to(BF16), enc_mask, p=p)
    dec_in = lm.shared(_shift_right(lm, labels)).to(BF16)
    dec = _t5_stack(lm.decoder, dec_in, None, enc=enc, enc_mask=enc_mask, p=p, prefix_kv=prefix_kv)
    scale = lm.model_dim ** -0.5 if getattr(lm.config, "scale_decoder_outputs", lm.config.tie_word_embeddings) else 1.0
    logits = ops.linear(dec, lm.lm_head.weight, None, alpha=scale)       # (x * s) W^T == s * (x W^T)
    loss = ops.cross_entropy(logits, labels.to(logits.device), ignore_index=-100)
    return LMOutput(loss=loss, logits=logits)


# ------------------------------------------------------------------------------------------------- OPT
def _opt_supported(lm) -> bool:
    cfg = lm.config
    d = cfg.hidden_size // cfg.num_attention_heads
    return (d in (64, 128) and cfg.hidden_size % 8 == 0 and cfg.ffn_dim % 8 == 0 and cfg.activation_function == "relu"
            and float(cfg.attention_dropout) == 0.0 and float(getattr(cfg, "layerdrop", 0.0)) == 0.0)


def opt_forward(lm, input_ids=None, attention_mask=None, inputs_embeds=None, labels=None, prefix_kv=None):
    """OPTForCausalLM.forward for training: learned positions from the attention mask (HF opt :56-70), pre- or post-LN
    decoder layers (:202-254) with causal + key-padding attention, final LayerNorm, lm_head, shifted CE
    (ignore_index -100: the concat path pads the bank positions of the labels with -100,
    model/modelling_self_attention.py:327-330).

    ``prefix_kv`` [n_virtual, layers, 2, H] (peft prefix tuning, model/modelling_self_attention.py:88-92; semantics
    restated from the published method -- peft is absent, parity unpinned): layer l attends over
    keys = [prefix_kv[:, l, 0] ; k_proj(x)], values = [prefix_kv[:, l, 1] ; v_proj(x)]; the virtual tokens are always
    visible (mask 1, before every query) and shift the learned positions of the real tokens by n_virtual, exactly as HF
    does when it is handed them as past_key_values."""
    dec = lm.model.decoder
    cfg = lm.config
    p = cfg.dropout if lm.training else 0.0
    if inputs_embeds is None:
        inputs_embeds = dec.embed_tokens(input_ids)
    b, s = inputs_embeds.shape[:2]
    if attention_mask is None:
        attention_mask = torch.ones(b, s, dtype=torch.long, device=inputs_embeds.device)
    am = attention_mask.long()
    n_pre = 0 if prefix_kv is None else prefix_kv.shape[0]
    if n_pre:
        full = torch.cat((torch.ones(b, n_pre, dtype=am.dtype, device=am.device), am), dim=1)
        pos = dec.embed_positions(full)[:, n_pre:]            # positions continue after the virtual tokens
    else:
        full = am
        pos = dec.embed_positions(am)
    if dec.project_in is not None:
        inputs_embeds = ops.linear(inputs_embeds, dec.project_in.weight)
    h = (inputs_embeds + pos.to(inputs_embeds.dtype)).to(BF16)
    key_mask = (full != 0).to(torch.uint8).contiguous()
    heads = cfg.num_attention_heads
    scale = (cfg.hidden_size // heads) ** -0.5
    for li, layer in enumerate(dec.layers):
        a = layer.self_attn
        ln1, ln2 = layer.self_attn_layer_norm, layer.final_layer_norm
        x, h = ops.layer_norm_fork(h, ln1.weight, ln1.bias, ln1.eps) if layer.do_layer_norm_before else (h, h)
        if n_pre:
            pk = prefix_kv[:, li, 0].to(BF16)[None].expand(b, -1, -1)
            pv = prefix_kv[:, li, 1].to(BF16)[None].expand(b, -1, -1)
            k = torch.cat((pk, _proj(a.k_proj, x)), dim=1)
            v = torch.cat((pv, _proj(a.v_proj, x)), dim=1)
            o = ops.attention(_proj(a.q_proj, x), k, v, key_mask=key_mask, heads=heads, causal=True, scale=scale)
        else:
            qkv = torch.cat([_proj(a.q_proj, x), _proj(a.k_proj, x), _proj(a.v_proj, x)], dim=-1)
            o = ops.self_attention(qkv, key_mask, heads, causal=True, scale=scale)
        h = ops.linear(o, a.out_proj.weight, a.out_proj.bias, residual=h, dropout_p=p)
        if not layer.do_layer_norm_before:
            h = ops.layer_norm(h, ln1.weight, ln1.bias, ln1.eps)
        x, h = ops.layer_norm_fork(h, ln2.weight, ln2.bias, ln2.eps) if layer.do_layer_norm_before else (h, h)
        h = ops.mlp(x, layer.fc1.weight, layer.fc1.bias, layer.fc2.weight, layer.fc2.bias, residual=h, dropout_p=p)
        if not layer.do_layer_norm_before:
            h = ops.layer_norm(h, ln2.weight, ln2.bias, ln2.eps)
    if dec.final_layer_norm is not None:
        ln = dec.final_layer_norm
        h = ops.layer_norm(h, ln.weight, ln.bias, ln.eps)
    if dec.project_out is not None:
        h = ops.linear(h, dec.project_out.weight)
    logits = ops.linear(h, lm.lm_head.weight)
    loss = None
    if labels is not None:
        loss = ops.shifted_cross_entropy(logits, labels.to(logits.device), ignore_index=-100)
    return LMOutput(loss=loss, logits=logits)


# ------------------------------------------------------------------------------------------------- dispatch
def supports(lm) -> bool:
    name = type(lm).__name__
    if name == "T5ForConditionalGeneration":
        return _t5_supported(lm)
    if name == "OPTForCausalLM":
        return _opt_supported(lm)
    return False


def forward(lm, **kw):
    """Run a HF T5ForConditionalGeneration / OPTForCausalLM training forward on this package's kernels."""
    if type(lm).__name__ == "T5ForConditionalGeneration":
        return t5_forward(lm, **kw)
    return opt_forward(lm, **kw)


__all__ = ["forward", "supports", "t5_forward", "opt_forward", "t5_rel_bias", "LMOutput"]
