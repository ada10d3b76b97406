"""nn.Module surface of MMGL's neighbor-fusion path, backed by the sm_100a kernels in libmmgl_b200.so.

Same class names, constructor arguments, ``forward`` keyword arguments and state-dict keys as the reference's
``model/modelling_cross_attention.py`` and ``model/graph.py`` (file:line citations are relative to the reference
root), so ``language_modelling/run_generation.py``'s loop -- ``DDP(model)``, ``outputs = model(**batch)``,
``outputs.loss.backward()`` -- runs unchanged with ``from mmgl_b200 import CrossAttentionModel, SelfAttentionModel``.

What runs where
  * gated cross-attention layers (``MPTDecoderLayer(cross_attention=True)``), the neighbor projections, the ragged
    bank packing (+ Laplacian-PE projection), the GCN, LoRA linears and every nn.Linear / LayerNorm / FFN of the
    frozen decoder layers: this package's CUDA kernels (``ops``), forward and backward.
  * the causal self-attention core of the frozen OPT layers (``ops.self_attention``), the frozen RoBERTa / CLIP
    encoders (``encoders``: HF weights read in place), the lm_head projection and the shifted cross-entropy: this
    package's kernels as well.
There is no CPU path and no library fallback: modules raise on CPU tensors, and shapes the kernels do not cover (head
dims other than 64 / 128, neighbor banks longer than ``ops.xattn_max_keys(head_dim)`` rows, arbitrary 4-D attention
masks handed in by other callers) are rejected in the constructors / at the call with a message that says so.

Reference defects that are deliberately NOT inherited: D1 (``neighbor_layer_wise`` vs
``num_neighbor_layers``), D2 (``neighbor_mode == "embedding"`` with flamingo means cross-attention), D4 (bank is
allocated in the compute dtype), D10 (``train()`` returns ``self``), D12 (``lm_head`` tied to ``embed_tokens``).
"""
from __future__ import annotations

import os
from types import SimpleNamespace
from typing import Optional

import torch
import torch.nn as nn

from . import configs, encoders, ops

BF16 = torch.bfloat16


# ------------------------------------------------------------------------------------------------- outputs
class CausalLMOutput(dict):
    """Minimal stand-in for HF's CausalLMOutputWithPast: attribute + key + index access to loss / logits."""

    def __init__(self, loss=None, logits=None):
        super().__init__(loss=loss, logits=logits)
        self.loss, self.logits = loss, logits

    def __getitem__(self, k):
        if isinstance(k, int):
            return [v for v in (self.loss, self.logits) if v is not None][k]
        return super().__getitem__(k)


# ------------------------------------------------------------------------------------------------- config
class MPTConfig:
    """OPT hyper-parameters + the MMGL knobs (model/modelling_cross_attention.py:82-121)."""

    def __init__(self, args, opt_config):
        nlw = getattr(args, "neighbor_layer_wise", None)
        if nlw is None:  # D1: the CLI only defines num_neighbor_layers
            nnl = int(getattr(args, "num_neighbor_layers", 4) or 4)
            nlw = max(1, opt_config.num_hidden_layers // max(1, nnl))
        self.neighbor_layer_wise = int(nlw)
        mode = getattr(args, "neighbor_mode", "cross_attention")
        self.peft_type = getattr(args, "peft_type", "flamingo")
        if mode == "embedding":  # D2
            mode = "cross_attention"
        self.neighbor_mode = mode
        self.lora_r = getattr(args, "lora_r", 64)
        self.lora_alpha = getattr(args, "lora_alpha", 1)
        self.lora_dropout = getattr(args, "lora_dropout", 0.0)
        for k in ("vocab_size", "max_position_embeddings", "num_attention_heads", "word_embed_proj_dim", "ffn_dim",
                  "hidden_size", "num_hidden_layers", "dropout", "attention_dropout", "activation_function",
                  "init_std", "layerdrop", "do_layer_norm_before", "enable_bias", "layer_norm_elementwise_affine",
                  "pad_token_id", "bos_token_id", "eos_token_id"):
            setattr(self, k, getattr(opt_config, k))
        self._remove_final_layer_norm = getattr(opt_config, "_remove_final_layer_norm", False)
        if self.activation_function != "relu":
            raise NotImplementedError("mmgl_b200 fuses ReLU into the FFN GEMM epilogues; OPT checkpoints use relu")
        if self.attention_dropout != 0.0:
            raise NotImplementedError("attention-probability dropout is not implemented (OPT uses 0.0)")


class MPTLearnedPositionalEmbedding(nn.Embedding):
    """OPT's learned positions with the +2 offset (model/modelling_cross_attention.py:124-145)."""

    def __init__(self, num_embeddings: int, embedding_dim: int):
        self.offset = 2
        super().__init__(num_embeddings + self.offset, embedding_dim)

    def forward(self, attention_mask, past_key_values_length: int = 0):
        am = attention_mask.long()
        positions = (torch.cumsum(am, dim=1) * am) - 1
        return super().forward(positions[:, past_key_values_length:] + self.offset)


class KeyPaddingCausalMask:
    """Compact form of the reference's additive decoder mask (causal AND key-not-padding,
    model/modelling_cross_attention.py:455-476): the [B,S] byte key mask; the [B,1,S,S] tensor is never built."""

    def __init__(self, key_mask):
        self.key_mask = key_mask


# ------------------------------------------------------------------------------------------------- attention
class MPTAttention(nn.Module):
    """model/modelling_cross_attention.py:148-275.  Cross branch: q/k/v projections (bias and the d^-1/2 scale in
    the GEMM epilogue) + the fused attention core + out_proj.  Self branch (frozen OPT layers): projections through
    the same GEMM kernel (one fused QKV GEMM when frozen), causal / key-padded core through ops.self_attention."""

    def __init__(self, config, cross_attention):
        super().__init__()
        self.embed_dim = config.hidden_size
        self.num_heads = config.num_attention_heads
        self.head_dim = self.embed_dim // self.num_heads
        if self.head_dim * self.num_heads != self.embed_dim:
            raise ValueError(f"embed_dim {self.embed_dim} not divisible by num_heads {self.num_heads}")
        if self.head_dim not in (64, 128):
            raise ValueError(f"mmgl_b200's attention kernels cover head_dim 64 and 128; hidden_size {self.embed_dim} / "
                             f"{self.num_heads} heads gives {self.head_dim} (no library fallback is provided)")
        self.scaling = self.head_dim ** -0.5
        bias = config.enable_bias
        self.k_proj = nn.Linear(self.embed_dim, self.embed_dim, bias=bias)
        self.q_proj = nn.Linear(self.embed_dim, self.embed_dim, bias=bias)
        self.out_proj = nn.Linear(self.embed_dim, self.embed_dim, bias=bias)
        self.v_proj = nn.Linear(self.embed_dim, self.embed_dim, bias=bias)
        self.cross_attention = cross_attention
        self.peft_type = config.peft_type

    @staticmethod
    def _key_mask(attention_mask, b, s):
        """The self-attention kernel takes the mask in its compact form: a [B,S] key-padding mask + a causal flag.
        MPTDecoder passes exactly that (``KeyPaddingCausalMask``); None means causal without padding; anything else
        (an arbitrary 4-D mask handed in by other callers) is rejected: there is no library fallback."""
        if attention_mask is None:
            return None, s > 1
        if isinstance(attention_mask, KeyPaddingCausalMask):
            return attention_mask.key_mask, True
        raise NotImplementedError(
            "mmgl_b200 self-attention takes the decoder mask in compact form (modules.KeyPaddingCausalMask([B,S] key mask), "
            "or None for causal without padding); arbitrary [B,1,S,S] masks are not supported and there is no fallback")

    def forward(self, hidden_states, attention_mask=None, neighbor_embeds=None, neighbor_attention_mask=None,
                layer_head_mask=None, past_key_value=None, output_attentions=False, residual=None, dropout_p=0.0):
        """Returns (attn_output, None, None) like the reference.  ``residual``/``dropout_p`` (extensions) fuse the
        hidden dropout and the residual add of the enclosing layer into the out_proj epilogue."""
        if layer_head_mask is not None or past_key_value is not None or output_attentions:
            raise NotImplementedError("head masks, KV caches and attention-weight outputs are not on the training path")
        if self.cross_attention:
            q = ops.linear(hidden_states, self.q_proj.weight, self.q_proj.bias, alpha=self.scaling)
            k = ops.linear(neighbor_embeds, self.k_proj.weight, self.k_proj.bias)
            v = ops.linear(neighbor_embeds, self.v_proj.weight, self.v_proj.bias)
            o = ops.xattn_core(q, k, v, neighbor_attention_mask, self.num_heads)
        else:
            b, s, _ = hidden_states.shape
            projs = (self.q_proj, self.k_proj, self.v_proj)
            if not any(p.weight.requires_grad for p in projs):
                # frozen block: one N = 3H GEMM over the cached row-concatenation Wq|Wk|Wv
                w = ops.fused_rows([p.weight for p in projs], BF16)
                bias = ops.fused_rows([p.bias for p in projs], torch.float32) if self.q_proj.bias is not None else None
                qkv = ops.linear(hidden_states, w, bias)
            else:
                qkv = torch.cat([ops.linear(hidden_states, p.weight, p.bias) for p in projs], dim=-1)
            key_mask, causal = self._key_mask(attention_mask, b, s)
            o = ops.self_attention(qkv, key_mask, self.num_heads, causal=causal, scale=self.scaling)
        out = ops.linear(o, self.out_proj.weight, self.out_proj.bias, residual=residual, dropout_p=dropout_p)
        return out, None, None


class MPTDecoderLayer(nn.Module):
    """model/modelling_cross_attention.py:278-375.  ``cross_attention=True`` is the Flamingo-style tanh-gated block
    (one fused forward/backward schedule, ops.GatedCrossLayerFn); otherwise the plain (frozen) OPT block."""

    def __init__(self, config, cross_attention=False):
        super().__init__()
        self.embed_dim = config.hidden_size
        self.self_attn = MPTAttention(config, cross_attention)
        self.do_layer_norm_before = config.do_layer_norm_before
        self.dropout = config.dropout
        affine = config.layer_norm_elementwise_affine
        self.self_attn_layer_norm = nn.LayerNorm(self.embed_dim, elementwise_affine=affine)
        self.fc1 = nn.Linear(self.embed_dim, config.ffn_dim, bias=config.enable_bias)
        self.fc2 = nn.Linear(config.ffn_dim, self.embed_dim, bias=config.enable_bias)
        self.final_layer_norm = nn.LayerNorm(self.embed_dim, elementwise_affine=affine)
        self.cross_attention = cross_attention
        self.peft_type = config.peft_type
        if self.cross_attention and self.peft_type == "flamingo":
            self.gating1 = nn.Parameter(torch.tensor(0.0))
            self.gating2 = nn.Parameter(torch.tensor(0.0))

    def forward(self, hidden_states, attention_mask=None, neighbor_embeds=None, neighbor_attention_mask=None,
                layer_head_mask=None, past_key_value=None, output_attentions=False, use_cache=False):
        if layer_head_mask is not None or past_key_value is not None or output_attentions or use_cache:
            raise NotImplementedError("head masks, KV caches and attention-weight outputs are not on the training path")
        p = self.dropout if self.training else 0.0
        a = self.self_attn
        ln1, ln2 = self.self_attn_layer_norm, self.final_layer_norm
        if self.cross_attention:
            flamingo = self.peft_type == "flamingo"
            y = ops.gated_cross_layer(
                hidden_states, neighbor_embeds, neighbor_attention_mask, ln1.weight, ln1.bias,
                a.q_proj.weight, a.q_proj.bias, a.k_proj.weight, a.k_proj.bias, a.v_proj.weight, a.v_proj.bias,
                a.out_proj.weight, a.out_proj.bias, self.gating1 if flamingo else None,
                ln2.weight, ln2.bias, self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias,
                self.gating2 if flamingo else None, a.num_heads, ln1.eps, self.do_layer_norm_before, p)
            return (y,)
        x = hidden_states
        if self.do_layer_norm_before:
            h, x_res = ops.layer_norm_fork(x, ln1.weight, ln1.bias, ln1.eps)
            h1, _, _ = a(h, attention_mask=attention_mask, residual=x_res, dropout_p=p)
            f_in, h1_res = ops.layer_norm_fork(h1, ln2.weight, ln2.bias, ln2.eps)
            y = ops.mlp(f_in, self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias, residual=h1_res, dropout_p=p)
        else:
            u, _, _ = a(x, attention_mask=attention_mask, residual=x, dropout_p=p)
            h1 = ops.layer_norm(u, ln1.weight, ln1.bias, ln1.eps)
            w = ops.mlp(h1, self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias, residual=h1, dropout_p=p)
            y = ops.layer_norm(w, ln2.weight, ln2.bias, ln2.eps)
        return (y,)


# ------------------------------------------------------------------------------------------------- decoder / LM
class MPTDecoder(nn.Module):
    """model/modelling_cross_attention.py:400-653: OPT decoder with a gated cross-attention layer after every
    ``neighbor_layer_wise``-th self-attention layer (:437-442, :613-625)."""

    def __init__(self, config: MPTConfig):
        super().__init__()
        self.config = config
        self.padding_idx = config.pad_token_id
        self.embed_tokens = nn.Embedding(config.vocab_size, config.word_embed_proj_dim, self.padding_idx)
        self.embed_positions = MPTLearnedPositionalEmbedding(config.max_position_embeddings, config.hidden_size)
        if config.word_embed_proj_dim != config.hidden_size:
            self.project_out = nn.Linear(config.hidden_size, config.word_embed_proj_dim, bias=False)
            self.project_in = nn.Linear(config.word_embed_proj_dim, config.hidden_size, bias=False)
        else:
            self.project_out = self.project_in = None
        if config.do_layer_norm_before and not config._remove_final_layer_norm:
            self.final_layer_norm = nn.LayerNorm(config.hidden_size,
                                                 elementwise_affine=config.layer_norm_elementwise_affine)
        else:
            self.final_layer_norm = None
        self.cross_attention = config.neighbor_mode == "cross_attention"
        self.neighbor_layer_wise = config.neighbor_layer_wise
        self.layers = nn.ModuleList()
        self.neighbor_layers = nn.ModuleList()
        for l in range(config.num_hidden_layers):
            self.layers.append(MPTDecoderLayer(config))
            if self.cross_attention and (l + 1) % self.neighbor_layer_wise == 0:
                self.neighbor_layers.append(MPTDecoderLayer(config, cross_attention=True))
        self.apply(self._init_weights)

    def _init_weights(self, module):  # :384-393
        std = self.config.init_std
        if isinstance(module, nn.Linear):
            module.weight.data.normal_(mean=0.0, std=std)
            if module.bias is not None:
                module.bias.data.zero_()
        elif isinstance(module, nn.Embedding):
            module.weight.data.normal_(mean=0.0, std=std)
            if module.padding_idx is not None:
                module.weight.data[module.padding_idx].zero_()

    def forward(self, input_ids=None, attention_mask=None, inputs_embeds=None, neighbor_embeds=None,
                neighbor_attention_mask=None):
        if (input_ids is None) == (inputs_embeds is None):
            raise ValueError("specify exactly one of input_ids / inputs_embeds")
        if inputs_embeds is None:
            inputs_embeds = self.embed_tokens(input_ids)
        bsz, seq = inputs_embeds.shape[:2]
        if attention_mask is None:
            attention_mask = torch.ones(bsz, seq, device=inputs_embeds.device, dtype=torch.long)
        # causal AND key-not-padding in compact form, shared by all frozen layers (:542-544 builds the additive twin)
        allowed = KeyPaddingCausalMask((attention_mask != 0).to(torch.uint8).contiguous())
        pos = self.embed_positions(attention_mask)
        if self.project_in is not None:
            inputs_embeds = ops.linear(inputs_embeds, self.project_in.weight)
        h = (inputs_embeds + pos).to(BF16)
        if neighbor_embeds is not None and neighbor_embeds.dtype != BF16:
            neighbor_embeds = neighbor_embeds.to(BF16)
        for idx, layer in enumerate(self.layers):
            with ops.nvtx_range(f"lm_layer_{idx}"):
                h = layer(h, attention_mask=allowed)[0]
            if self.cross_attention and neighbor_embeds is not None and (idx + 1) % self.neighbor_layer_wise == 0:
                k = (idx + 1) // self.neighbor_layer_wise - 1
                with ops.nvtx_range(f"gated_xattn_layer_{k}"):
                    h = self.neighbor_layers[k](h, neighbor_embeds=neighbor_embeds,
                                                neighbor_attention_mask=neighbor_attention_mask)[0]
        if self.final_layer_norm is not None:
            ln = self.final_layer_norm
            h = ops.layer_norm(h, ln.weight, ln.bias, ln.eps)
        if self.project_out is not None:
            h = ops.linear(h, self.project_out.weight)
        return h


class MPTModel(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.decoder = MPTDecoder(config)

    def forward(self, **kw):
        return self.decoder(**kw)


def mark_only_peft_as_trainable(model):
    """Freeze everything, re-enable the cross-attention layers (model/modelling_cross_attention.py:731-737)."""
    for p in model.parameters():
        p.requires_grad = False
    for m in model.modules():
        if isinstance(m, MPTDecoderLayer) and m.cross_attention:
            for p in m.parameters():
                p.requires_grad = True


class MPTForCausalLM(nn.Module):
    """model/modelling_cross_attention.py:739-876.  ``lm_head`` is tied to ``embed_tokens`` (D12)."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.model = MPTModel(config)
        self.lm_head = nn.Linear(config.word_embed_proj_dim, config.vocab_size, bias=False)
        self.lm_head.weight = self.model.decoder.embed_tokens.weight
        if config.peft_type != "none":
            mark_only_peft_as_trainable(self.model)

    def get_input_embeddings(self):
        return self.model.decoder.embed_tokens

    def forward(self, input_ids=None, attention_mask=None, inputs_embeds=None, labels=None, neighbor_embeds=None,
                neighbor_attention_mask=None, **unused):
        h = self.model.decoder(input_ids=input_ids, attention_mask=attention_mask, inputs_embeds=inputs_embeds,
                               neighbor_embeds=neighbor_embeds, neighbor_attention_mask=neighbor_attention_mask)
        logits = ops.linear(h, self.lm_head.weight)                                               # :826
        loss = None
        if labels is not None:                                                                    # :828-836
            loss = ops.shifted_cross_entropy(logits, labels.to(logits.device))
        return CausalLMOutput(loss=loss, logits=logits)


# ------------------------------------------------------------------------------------------------- graph PE
class GCN(nn.Module):
    """2-layer mean-aggregate GCN with a null root node (model/graph.py:6-31)."""

    def __init__(self, input_dim, output_dim, hidden_dim=128, use_bias=False):
        super().__init__()
        if use_bias:
            raise NotImplementedError("the reference never enables the GCN bias (model/graph.py:8)")
        self.w1 = nn.Linear(2 * input_dim, hidden_dim, bias=False)
        self.w2 = nn.Linear(2 * hidden_dim, output_dim, bias=False)

    def forward(self, X, adj):
        return ops.gcn(X, adj, self.w1.weight, self.w2.weight)


class TextPooler(nn.Module):
    """tanh(Linear(h[:, 0])) (model/modelling_cross_attention.py:879-893)."""

    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.hidden_size)
        self.activation = nn.Tanh()

    def forward(self, hidden_states):
        return self.pool_cls(hidden_states[:, 0])

    def pool_cls(self, cls_hidden):
        return self.activation(ops.linear(cls_hidden, self.dense.weight, self.dense.bias))


# ------------------------------------------------------------------------------------------------- loaders
def _load_or_init(kind: str, name, auto_cls_name: str):
    """``from_pretrained(name)`` exactly as the reference does (model/modelling_cross_attention.py:920, :934, :953-954;
    model/modelling_self_attention.py:68, :72, :111, :125): a local ``save_pretrained`` directory or a checkpoint name
    found in the local HF cache.  When the weights cannot be found (no network, empty cache) the call RAISES unless the
    caller opted in to random initialisation from the built-in architecture config -- ``MMGL_ALLOW_RANDOM_INIT=1`` in
    the environment, which bench.py and the synthetic-data tests set -- and then it says so loudly.  Passing a HF config
    object instead of a name is the explicit form of the same opt-in (tests)."""
    import transformers
    cls = getattr(transformers, auto_cls_name)
    if isinstance(name, transformers.PretrainedConfig):   # extension: explicit config object -> random init
        return cls(name)
    if isinstance(name, str) and os.path.isdir(name):
        return cls.from_pretrained(name)
    err = None
    try:
        return cls.from_pretrained(name, local_files_only=os.environ.get("HF_HUB_OFFLINE", "0") == "1"
                                   or os.environ.get("TRANSFORMERS_OFFLINE", "0") == "1")
    except Exception as e:  # noqa: BLE001 -- hub / cache / network errors come in many types
        err = e
    if os.environ.get("MMGL_ALLOW_RANDOM_INIT", "0") != "1":
        raise RuntimeError(
            f"mmgl_b200: pretrained weights for {name!r} were not found ({type(err).__name__}: {err}).  The reference "
            f"always trains from from_pretrained() weights; set MMGL_ALLOW_RANDOM_INIT=1 to build a RANDOM-initialised "
            f"{auto_cls_name} of the same architecture instead (synthetic benchmarks / tests only).") from err
    import warnings
    cfg = {"lm": configs.lm_config, "text": configs.text_config, "visual": configs.visual_config}[kind](name)
    warnings.warn(f"mmgl_b200: {name!r} is RANDOM-INITIALISED from its architecture config (MMGL_ALLOW_RANDOM_INIT=1); "
                  f"no pretrained weights were loaded", RuntimeWarning, stacklevel=2)
    return cls(cfg)


class _NeighborEncoderMixin:
    """Frozen neighbor encoders + trainable projections shared by both wrappers
    (model/modelling_cross_attention.py:914-940, 978-1027; model/modelling_self_attention.py:106-132, 154-200)."""

    def _build_encoders(self, args, embed_dim, with_text, with_visual, with_pos):
        self.text_model = None
        if with_text:
            if "clip" in str(args.text_model):
                raise NotImplementedError(
                    "CLIP TEXT towers as the neighbor text encoder (model/modelling_cross_attention.py:919) are not built on "
                    "the package's kernels yet (RoBERTa is); there is no library fallback")
            else:
                self.text_model = _load_or_init("text", args.text_model, "RobertaModel")
                self.text_pooler = TextPooler(self.text_model.config)
            self.text_embeddings = nn.Linear(self.text_model.config.hidden_size, embed_dim * args.n_text_tokens)
            if with_pos:
                self.text_position_embeddings = nn.Embedding(args.max_output_length + 1, embed_dim * args.n_text_tokens)
            self.text_model.eval()
            for p in self.text_model.parameters():
                p.requires_grad = False
        self.visual_model = None
        if with_visual:
            self.visual_model = _load_or_init("visual", args.visual_model, "CLIPVisionModel")
            self.visual_embeddings = nn.Linear(self.visual_model.config.hidden_size, embed_dim * args.n_visual_tokens)
            if with_pos:
                self.visual_position_embeddings = nn.Embedding(args.max_output_length + 1,
                                                               embed_dim * args.n_visual_tokens)
            self.visual_model.eval()
            for p in self.visual_model.parameters():
                p.requires_grad = False

    def encode_text(self, input_ids, attention_mask, pack=None):
        """frozen text encoder (+ trainable pooler) -> pooled features [B*T, E]
        (model/modelling_cross_attention.py:988-996).  ``pack``: host-made packing plan (mmgl_b200.plan), or None."""
        l = input_ids.shape[-1]
        ids2, am2 = input_ids.reshape(-1, l), attention_mask.reshape(-1, l)
        # frozen RoBERTa on this package's kernels; only the [CLS] row is consumed (TextPooler: hidden[:, 0])
        if pack is not None:
            return self.text_pooler.pool_cls(encoders.roberta_cls_hidden(self.text_model, ids2, am2, plan=pack))
        return self.text_pooler.pool_cls(encoders.roberta_cls_hidden(self.text_model, ids2, am2))

    def encode_images(self, pixel_values):
        """frozen CLIP vision tower -> pooler_output [B*I, E] (model/modelling_cross_attention.py:1015-1019)."""
        return encoders.clip_pooler_output(self.visual_model, pixel_values.reshape(-1, *pixel_values.shape[2:]))

    skip_padding_neighbors = True

    def _needed(self, pos_ids):
        """Neighbors whose frozen-encoder output can influence the step: the valid ones (pos_id > 0).  A padding
        neighbor's bank rows are masked, get softmax weight exactly 0 and gradient exactly 0 (SURVEY invariant I2),
        so its encoder pass is skipped and its projection rows are left zero -- loss and gradients are identical.
        Exception kept for parity: a sample with NO valid neighbor attends uniformly over its masked rows in the
        reference, so all of its neighbors stay 'needed'.  Not applied with position_type == "gnn" (the GCN mixes rows
        before the mask acts).  One host sync per call (index list)."""
        if pos_ids is None or not self.skip_padding_neighbors or getattr(self, "position_type", "none") == "gnn":
            return None   # the GCN aggregates bank rows BEFORE masking: an edge into a padding slot would see W*enc+b
        valid = pos_ids > 0
        need = valid | ~valid.any(dim=1, keepdim=True)
        if bool(need.all()):
            return None
        return need.reshape(-1).nonzero(as_tuple=False).squeeze(1)

    @staticmethod
    def _scatter_rows(y, idx, total):
        if idx is None:
            return y
        full = torch.zeros((total, y.shape[-1]), dtype=y.dtype, device=y.device)
        return full.index_copy(0, idx, y)

    def text_projection(self, input_ids, attention_mask, pos_ids=None, plan=None):
        """pooled [B*T, E] -> Linear(E -> n_tok*H): [B, T, n_tok*H] (the position-embedding add is fused into the
        bank packing kernel; :997).  With a host-made ``plan`` (mmgl_b200.plan.NeighborPlan) no device->host read happens."""
        b, n, l = input_ids.shape
        use_plan = plan is not None and self._plan_usable()
        idx = plan.text_idx if use_plan else self._needed(pos_ids)
        ids2, am2 = input_ids.reshape(-1, l), attention_mask.reshape(-1, l)
        if idx is not None:
            ids2, am2 = ids2.index_select(0, idx), am2.index_select(0, idx)
        pack = None
        if use_plan:
            pack = (plan.tok_idx, plan.cu, plan.total, plan.max_len) if plan.pack else False
        pooled = self.encode_text(ids2, am2, pack) if use_plan else self.encode_text(ids2, am2)
        y = ops.linear(pooled, self.text_embeddings.weight, self.text_embeddings.bias)
        return self._scatter_rows(y, idx, b * n).reshape(b, n, -1)

    def _plan_usable(self):
        """a plan made with the default rules is only valid while the module still follows them"""
        return self.skip_padding_neighbors and getattr(self, "position_type", "none") != "gnn" and encoders.PACK_PADDING

    def visual_projection(self, pixel_values, pos_ids=None, plan=None):
        b, n = pixel_values.shape[:2]
        idx = plan.image_idx if (plan is not None and self._plan_usable()) else self._needed(pos_ids)
        px = pixel_values.reshape(b * n, 1, *pixel_values.shape[2:])
        if idx is not None:
            px = px.index_select(0, idx)
        y = ops.linear(self.encode_images(px), self.visual_embeddings.weight, self.visual_embeddings.bias)
        return self._scatter_rows(y, idx, b * n).reshape(b, n, -1)

    def _table(self, name):
        emb = getattr(self, name, None)
        return None if emb is None else emb.weight

    def build_bank(self, neighbor_input_ids, neighbor_attention_mask, neighbor_pos_ids, text_locations,
                   neighbor_images=None, neighbor_images_pos_ids=None, image_locations=None, lpe=None,
                   use_pos_tables=True, plan=None):
        """bank [B,(T+I)*n_tok,H] bf16 + byte mask [B,(T+I)*n_tok]
        (model/modelling_cross_attention.py:1072-1104; model/modelling_self_attention.py:263-315)."""
        if neighbor_images is not None and self.n_text_tokens != self.n_visual_tokens:
            raise ValueError("the packed bank needs n_text_tokens == n_visual_tokens (reference :1093-1098)")
        tp = self.text_projection(neighbor_input_ids, neighbor_attention_mask, neighbor_pos_ids, plan)
        ip = self.visual_projection(neighbor_images, neighbor_images_pos_ids, plan) if neighbor_images is not None else None
        lpe_lin = getattr(self, "lpe_embeddings", None) if lpe is not None else None
        return ops.bank_pack(
            tp, self._table("text_position_embeddings") if use_pos_tables else None, neighbor_pos_ids, text_locations,
            ip, self._table("visual_position_embeddings") if (use_pos_tables and ip is not None) else None,
            neighbor_images_pos_ids, image_locations,
            lpe=lpe if lpe_lin is not None else None,
            lpe_weight=lpe_lin.weight if lpe_lin is not None else None,
            lpe_bias=lpe_lin.bias if lpe_lin is not None else None, n_tok=self.n_text_tokens)

    def _freeze_modes(self):
        if getattr(self.args, "freeze_lm", False):
            self.lm.eval()
        if self.text_model is not None:
            self.text_model.eval()
        if self.visual_model is not None:
            self.visual_model.eval()


# ------------------------------------------------------------------------------------------------- wrapper A
class CrossAttentionModel(nn.Module, _NeighborEncoderMixin):
    """Drop-in for the reference's CrossAttentionModel (model/modelling_cross_attention.py:896-1114).

    ``args`` attributes read: context, neighbor_mode, n_text_tokens, n_visual_tokens, model_name_or_path,
    text_model, visual_model, max_output_length, freeze_lm, peft_type, lora_*, neighbor_layer_wise |
    num_neighbor_layers, and (extension) position_type / max_text_neighbors / max_image_neighbors for
    graph positional encodings on the bank."""

    def __init__(self, args, tokenizer=None):
        super().__init__()
        self.args = args
        self.context = args.context
        mode = args.neighbor_mode
        if mode == "embedding" and getattr(args, "peft_type", "flamingo") in ("flamingo", "none"):
            mode = "cross_attention"  # D2
        self.neighbor_mode = mode
        self.n_text_tokens = args.n_text_tokens
        self.n_visual_tokens = args.n_visual_tokens
        self.tokenizer = tokenizer
        self.position_type = getattr(args, "position_type", "none")
        self.initialize_lm(args)
        self.input_embeddings = self.lm.get_input_embeddings()
        h = self.input_embeddings.embedding_dim
        self._build_encoders(args, h, with_text=self.context != "section_only",
                             with_visual=self.context in ("section_all", "all"), with_pos=True)
        if self.position_type == "laplacian":
            k = 1 + args.max_text_neighbors + args.max_image_neighbors - 5
            self.lpe_embeddings = nn.Linear(k, h * args.n_text_tokens)
        elif self.position_type == "gnn":
            d = h * args.n_text_tokens
            self.gnn = GCN(input_dim=d, output_dim=d, hidden_dim=self.text_model.config.hidden_size)
        if getattr(args, "freeze_lm", False):
            self.lm.eval()
            for p in self.lm.parameters():
                p.requires_grad = False
        # the cross-attention core keeps a head's whole bank on chip: reject configurations it cannot hold here, not at
        # the first forward
        n_nbrs = (getattr(args, "max_text_neighbors", None) or 0) + (getattr(args, "max_image_neighbors", None) or 0)
        cfg = self.lm.config
        limit = ops.xattn_max_keys(cfg.hidden_size // cfg.num_attention_heads)
        if self.neighbor_mode == "cross_attention" and n_nbrs * self.n_text_tokens > limit:
            raise ValueError(f"neighbor bank of {n_nbrs} neighbors x {self.n_text_tokens} tokens = "
                             f"{n_nbrs * self.n_text_tokens} rows exceeds the cross-attention kernel's limit of {limit} "
                             f"rows at head_dim {cfg.hidden_size // cfg.num_attention_heads}")

    def initialize_lm(self, args):
        """OPT weights copied into the self-attention layers, cross layers fresh (:951-976)."""
        name = args.model_name_or_path.replace("mpt", "opt") if isinstance(args.model_name_or_path, str) else args.model_name_or_path
        opt_model = _load_or_init("lm", name, "OPTForCausalLM")
        mpt = MPTForCausalLM(MPTConfig(args, opt_model.config))
        src, dst = opt_model.model.decoder, mpt.model.decoder
        dst.embed_tokens.load_state_dict(src.embed_tokens.state_dict())
        dst.embed_positions.load_state_dict(src.embed_positions.state_dict())
        if dst.project_in is not None:
            dst.project_in.load_state_dict(src.project_in.state_dict())
            dst.project_out.load_state_dict(src.project_out.state_dict())
        if dst.final_layer_norm is not None:
            dst.final_layer_norm.load_state_dict(src.final_layer_norm.state_dict())
        for i in range(len(src.layers)):
            dst.layers[i].load_state_dict(src.layers[i].state_dict(), strict=False)
        self.lm = mpt

    def train(self, mode=True):
        super().train(mode)
        self._freeze_modes()
        return self  # D10

    def forward(self, input_ids, attention_mask, labels, images=None, image_positions=None, neighbor_input_ids=None,
                neighbor_attention_mask=None, neighbor_pos_ids=None, text_locations=None, neighbor_images=None,
                neighbor_images_pos_ids=None, image_locations=None, lpe=None, graph=None, neighbor_plan=None):
        """``neighbor_plan`` (extension, optional): mmgl_b200.plan.attach_plan(batch) made on the host by the data pipeline;
        with it the step issues no device->host read."""
        if self.neighbor_mode == "raw" or self.context == "section_only":
            bank = mask = None                                                                     # :1068-1071
        elif self.neighbor_mode == "cross_attention" and self.context == "text_only":
            with ops.nvtx_range("neighbor_bank"):
                bank, mask = self.build_bank(neighbor_input_ids, neighbor_attention_mask, neighbor_pos_ids, None,
                                             plan=neighbor_plan)
        elif self.neighbor_mode == "cross_attention" and self.context in ("section_all", "all"):
            with ops.nvtx_range("neighbor_bank"):
                bank, mask = self.build_bank(neighbor_input_ids, neighbor_attention_mask, neighbor_pos_ids,
                                             text_locations, neighbor_images, neighbor_images_pos_ids, image_locations,
                                             lpe=lpe if self.position_type == "laplacian" else None, plan=neighbor_plan)
            if self.position_type == "gnn" and graph is not None:                                  # D9 extension
                b, nk, h = bank.shape
                flat = bank.reshape(b, nk // self.n_text_tokens, self.n_text_tokens * h)
                bank = (flat + self.gnn(flat, graph)).reshape(b, nk, h)
        else:
            raise ValueError(f"Neighbor mode: {self.neighbor_mode} and context: {self.context} are not supported.")
        with ops.nvtx_range("lm_forward_and_loss"):
            return self.lm(input_ids=input_ids, attention_mask=attention_mask, labels=labels, neighbor_embeds=bank,
                           neighbor_attention_mask=mask)


def prepare_for_training(model: nn.Module, device="cuda") -> nn.Module:
    """Move to ``device``; store FROZEN parameters in bf16 (they are only ever read by bf16 kernels) and keep
    trainable ones in fp32 (master weights; kernels read cached bf16 shadows, gradients arrive in fp32)."""
    model.to(device)
    for p in model.parameters():
        if not p.requires_grad and p.is_floating_point():
            p.data = p.data.to(BF16)
    for b in model.buffers():
        if b.is_floating_point():
            b.data = b.data.to(BF16)
    return model
