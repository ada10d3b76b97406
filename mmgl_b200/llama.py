"""Llama-2 with interleaved gated cross-attention layers (BASELINE configs[4]) -- an EXTENSION: the
reference dispatches only t5 / opt / mpt (language_modelling/run_generation.py:286-301) and its MPT classes hard-code
OPT's learned positions / LayerNorm / ReLU, so there is no reference Llama wrapper.  What is built here is the survey's
definition: a frozen HF ``LlamaForCausalLM`` whose decoder layers run on this package's kernels, with the reference's
``MPTDecoderLayer(cross_attention=True)`` (model/modelling_cross_attention.py:278-375; dimension-agnostic: LayerNorm,
ReLU FFN, tanh gates) inserted after every ``neighbor_layer_wise``-th layer, fed by the same neighbor bank as
``CrossAttentionModel`` (model/modelling_cross_attention.py:1038-1114).

Frozen Llama layer on the kernels (HF models/llama/modeling_llama.py: LlamaDecoderLayer :292-334, LlamaAttention :225-290,
LlamaMLP :171-184, LlamaRMSNorm :53-70, rotary embedding :73-170): RMSNorm -> one Q|K|V GEMM -> RoPE in place on q, k
(``mmgl_rope_inplace``) -> causal + key-padding attention (head_dim 128) -> o_proj with the residual in the epilogue ->
RMSNorm -> one gate|up GEMM -> SwiGLU (``mmgl_swiglu_fwd``) -> down_proj with the residual in the epilogue.  The weights
of the HF module are read in place (state-dict keys stay HF's).  Parity: tests/test_gpu_llama.py against HF's own fp32
forward / backward (gates at 0: invariant I1) and against HF layers + the oracle's gated layer (gates live).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from .modules import (BF16, GCN, CausalLMOutput, MPTDecoderLayer, _load_or_init, _NeighborEncoderMixin)

_rope_cache = {}


def rope_table(lm, seq, device):
    """cos / sin of HF's own LlamaRotaryEmbedding for positions 0..seq-1 as fp32 [seq, head_dim / 2, 2] (the two halves
    of HF's [seq, head_dim] tables are identical: emb = cat(freqs, freqs))."""
    key = (id(lm), seq, str(device))
    t = _rope_cache.get(key)
    if t is None:
        pos = torch.arange(seq, device=device)[None]
        cos, sin = lm.model.rotary_emb(torch.zeros(1, dtype=torch.float32, device=device), position_ids=pos)
        half = cos.shape[-1] // 2
        t = torch.stack((cos[0, :, :half].float(), sin[0, :, :half].float()), dim=-1).contiguous()
        _rope_cache[key] = t
    return t


def supports(lm) -> bool:
    cfg = lm.config
    d = getattr(cfg, "head_dim", None) or cfg.hidden_size // cfg.num_attention_heads
    return (type(lm).__name__ == "LlamaForCausalLM" and d in (64, 128) and cfg.num_key_value_heads == cfg.num_attention_heads
            and cfg.hidden_act == "silu" and cfg.hidden_size % 8 == 0 and cfg.intermediate_size % 8 == 0
            and float(getattr(cfg, "attention_dropout", 0.0)) == 0.0)


def llama_layer(layer, h, key_mask, cos_sin, heads, eps):
    """One frozen LlamaDecoderLayer (training forward; autograd carries the input gradient through the same kernels)."""
    a, mlp = layer.self_attn, layer.mlp
    d = h.shape[-1] // heads
    x, res = ops.rms_norm_fork(h, layer.input_layernorm.weight, eps)
    w = ops.fused_rows([a.q_proj.weight, a.k_proj.weight, a.v_proj.weight], BF16)
    bias = ops.fused_rows([a.q_proj.bias, a.k_proj.bias, a.v_proj.bias], torch.float32) if a.q_proj.bias is not None else None
    qkv = ops.rope_qk(ops.linear(x, w, bias), cos_sin, heads)
    o = ops.self_attention(qkv, key_mask, heads, causal=True, scale=d ** -0.5)
    h = ops.linear(o, a.o_proj.weight, a.o_proj.bias, residual=res)
    x, res = ops.rms_norm_fork(h, layer.post_attention_layernorm.weight, eps)
    wgu = ops.fused_rows([mlp.gate_proj.weight, mlp.up_proj.weight], BF16)
    bgu = ops.fused_rows([mlp.gate_proj.bias, mlp.up_proj.bias], torch.float32) if mlp.gate_proj.bias is not None else None
    return ops.linear(ops.swiglu(ops.linear(x, wgu, bgu)), mlp.down_proj.weight, mlp.down_proj.bias, residual=res)


class _GatedConfig:
    """the fields MPTDecoderLayer reads, at the Llama's width"""

    def __init__(self, cfg, dropout, peft_type):
        self.hidden_size, self.num_attention_heads, self.ffn_dim = cfg.hidden_size, cfg.num_attention_heads, cfg.intermediate_size
        self.enable_bias, self.dropout, self.do_layer_norm_before = True, dropout, True
        self.layer_norm_elementwise_affine, self.peft_type = True, peft_type


class GatedLlamaForCausalLM(nn.Module):
    """Frozen HF Llama + gated cross-attention layers after every ``neighbor_layer_wise``-th decoder layer."""

    def __init__(self, llama, args):
        super().__init__()
        if not supports(llama):
            raise NotImplementedError("mmgl_b200 runs HF LlamaForCausalLM with head_dim 64 / 128, no grouped-query attention and a "
                                      "SiLU MLP on its kernels; this configuration is not covered (there is no library fallback)")
        self.model = llama
        cfg = llama.config
        self.config = cfg
        nlw = getattr(args, "neighbor_layer_wise", None)
        if nlw is None:
            nlw = max(1, cfg.num_hidden_layers // max(1, int(getattr(args, "num_neighbor_layers", 4) or 4)))
        self.neighbor_layer_wise = int(nlw)
        gcfg = _GatedConfig(cfg, float(getattr(args, "neighbor_dropout", 0.1)), getattr(args, "peft_type", "flamingo"))
        self.neighbor_layers = nn.ModuleList(MPTDecoderLayer(gcfg, cross_attention=True)
                                             for _ in range(cfg.num_hidden_layers // self.neighbor_layer_wise))
        std = getattr(cfg, "initializer_range", 0.02)
        for m in self.neighbor_layers.modules():
            if isinstance(m, nn.Linear):
                m.weight.data.normal_(mean=0.0, std=std)
                if m.bias is not None:
                    m.bias.data.zero_()
        for p in self.model.parameters():                                   # mark_only_peft_as_trainable (:731-737)
            p.requires_grad = False

    def get_input_embeddings(self):
        return self.model.get_input_embeddings()

    def forward(self, input_ids=None, attention_mask=None, inputs_embeds=None, labels=None, neighbor_embeds=None,
                neighbor_attention_mask=None, **unused):
        lm, cfg = self.model, self.config
        if inputs_embeds is None:
            inputs_embeds = lm.model.embed_tokens(input_ids)
        b, s = inputs_embeds.shape[:2]
        if attention_mask is None:
            attention_mask = torch.ones(b, s, dtype=torch.long, device=inputs_embeds.device)
        key_mask = (attention_mask != 0).to(torch.uint8).contiguous()
        heads, eps = cfg.num_attention_heads, cfg.rms_norm_eps
        cos_sin = rope_table(lm, s, inputs_embeds.device)
        h = inputs_embeds.to(BF16)
        if neighbor_embeds is not None and neighbor_embeds.dtype != BF16:
            neighbor_embeds = neighbor_embeds.to(BF16)
        for idx, layer in enumerate(lm.model.layers):
            h = llama_layer(layer, h, key_mask, cos_sin, heads, eps)
            if neighbor_embeds is not None and (idx + 1) % self.neighbor_layer_wise == 0:
                k = (idx + 1) // self.neighbor_layer_wise - 1
                h = self.neighbor_layers[k](h, neighbor_embeds=neighbor_embeds, neighbor_attention_mask=neighbor_attention_mask)[0]
        h = ops.rms_norm(h, lm.model.norm.weight, eps)
        logits = ops.linear(h, lm.lm_head.weight)
        loss = ops.shifted_cross_entropy(logits, labels.to(logits.device)) if labels is not None else None
        return CausalLMOutput(loss=loss, logits=logits)


class LlamaCrossAttentionModel(nn.Module, _NeighborEncoderMixin):
    """``CrossAttentionModel`` (model/modelling_cross_attention.py:896-1114) with a Llama language model: same constructor
    arguments, forward keyword arguments and neighbor-bank construction; ``args.model_name_or_path`` names a Llama."""

    def __init__(self, args, tokenizer=None):
        super().__init__()
        self.args = args
        self.context = args.context
        self.neighbor_mode = "cross_attention" if args.neighbor_mode == "embedding" else args.neighbor_mode
        self.n_text_tokens, self.n_visual_tokens = args.n_text_tokens, args.n_visual_tokens
        self.tokenizer = tokenizer
        self.position_type = getattr(args, "position_type", "none")
        self.lm = GatedLlamaForCausalLM(_load_or_init("lm", args.model_name_or_path, "LlamaForCausalLM"), args)
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
        n_nbrs = (getattr(args, "max_text_neighbors", None) or 0) + (getattr(args, "max_image_neighbors", None) or 0)
        cfg = self.lm.config
        limit = ops.xattn_max_keys(cfg.hidden_size // cfg.num_attention_heads)
        if n_nbrs * self.n_text_tokens > limit:
            raise ValueError(f"neighbor bank of {n_nbrs * self.n_text_tokens} rows exceeds the cross-attention kernel's limit of {limit}")

    def train(self, mode=True):
        super().train(mode)
        self.lm.model.eval()
        self._freeze_modes()
        return self

    def forward(self, input_ids, attention_mask, labels, images=None, image_positions=None, neighbor_input_ids=None,
                neighbor_attention_mask=None, neighbor_pos_ids=None, text_locations=None, neighbor_images=None,
                neighbor_images_pos_ids=None, image_locations=None, lpe=None, graph=None, neighbor_plan=None):
        if self.neighbor_mode == "raw" or self.context == "section_only":
            bank = mask = None
        elif self.context == "text_only":
            bank, mask = self.build_bank(neighbor_input_ids, neighbor_attention_mask, neighbor_pos_ids, None, plan=neighbor_plan)
        elif self.context in ("section_all", "all"):
            bank, mask = self.build_bank(neighbor_input_ids, neighbor_attention_mask, neighbor_pos_ids, text_locations,
                                         neighbor_images, neighbor_images_pos_ids, image_locations,
                                         lpe=lpe if self.position_type == "laplacian" else None, plan=neighbor_plan)
            if self.position_type == "gnn" and graph is not None:
                b, nk, h = bank.shape
                flat = bank.reshape(b, nk // self.n_text_tokens, self.n_text_tokens * h)
                bank = (flat + self.gnn(flat, graph)).reshape(b, nk, h)
        else:
            raise ValueError(f"Neighbor mode: {self.neighbor_mode} and context: {self.context} are not supported.")
        return self.lm(input_ids=input_ids, attention_mask=attention_mask, labels=labels, neighbor_embeds=bank,
                       neighbor_attention_mask=mask)
