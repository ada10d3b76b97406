"""SelfAttentionModel: the concatenated-embedding path (reference: model/modelling_self_attention.py:48-336).

Neighbor embeddings (projected, packed, optionally shifted by the Laplacian-PE projection or the GCN) are appended
to the token embeddings and the whole sequence runs through a T5 / OPT language model.  On this path the package's CUDA
kernels are: the neighbor projections, the ragged bank packing with the fused position-embedding / LPE add
(ops.bank_pack), the GCN (ops.gcn), the LoRA linears on the LM's q / v projections (ops.lora_linear: the rank-r update
accumulates into the same TMEM tile as the base product) and the language model's own layer stack -- attention with the
T5 relative-position bias or the OPT causal / padding mask, RMSNorm / LayerNorm, FFN, lm_head and the loss
(mmgl_b200/lm.py, which reads the weights of the HF module in place).  Models that file cannot run (gated-GELU T5,
head dims other than 64 / 128) are rejected with an error: there is no HF / eager fallback.

peft is not importable in this image and its source is absent, so LoRA / prompt / prefix tuning are restated from
their published definitions (parity unpinned, see oracle/mmgl_oracle.py:lora_linear); module and state-dict names follow
peft's layout (``lm.base_model.model.<path>.q.lora_A.default.weight`` ...) so reference checkpoints load.

Reference defects not inherited: D3 (``session``/``session_all`` typos: the documented names
``section_only`` / ``section_all`` are accepted), D4 (bank dtype), D7 (LoRA targets the LM's real q / v projection
names), D10 (``train()`` returns self).
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import lm as lm_kernels
from . import ops
from .modules import BF16, GCN, _load_or_init, _NeighborEncoderMixin

_TEXT_ONLY = ("session", "section_only", "text_only")
_ALL = ("session_all", "section_all", "all")


class LoRALinear(nn.Module):
    """y = x W^T + b + (alpha / r) * (dropout(x) A^T) B^T around a frozen nn.Linear (peft LoRA layer layout)."""

    def __init__(self, base: nn.Linear, r: int, alpha: float, dropout: float):
        super().__init__()
        self.base_layer = base
        for p in base.parameters():
            p.requires_grad = False
        self.lora_A = nn.ModuleDict({"default": nn.Linear(base.in_features, r, bias=False)})
        self.lora_B = nn.ModuleDict({"default": nn.Linear(r, base.out_features, bias=False)})
        nn.init.kaiming_uniform_(self.lora_A["default"].weight, a=math.sqrt(5))   # modelling_cross_attention.py:719-724
        nn.init.zeros_(self.lora_B["default"].weight)
        self.scaling = alpha / r
        self.lora_dropout = dropout

    def forward(self, x):
        a, b = self.lora_A["default"].weight, self.lora_B["default"].weight
        base = self.base_layer
        if self.training and self.lora_dropout > 0.0:
            side = ops.linear(ops.linear(F.dropout(x, self.lora_dropout), a), b, alpha=self.scaling)
            return ops.linear(x, base.weight, base.bias, residual=side)
        return ops.lora_linear(x, base.weight, base.bias, a, b, self.scaling)


class _PeftHolder(nn.Module):
    def __init__(self, model):
        super().__init__()
        self.model = model


class _PeftShim(nn.Module):
    """Gives the adapted LM peft's attribute / state-dict layout: ``lm.base_model.model.<hf path>``."""

    def __init__(self, model, prompt_tokens: int = 0, prefix_tokens: int = 0):
        super().__init__()
        self.base_model = _PeftHolder(model)
        self.num_virtual_tokens = prompt_tokens
        self.num_prefix_tokens = prefix_tokens
        if prompt_tokens:
            dim = model.get_input_embeddings().embedding_dim
            self.prompt_encoder = nn.ModuleDict({"default": nn.ModuleDict({"embedding": nn.Embedding(prompt_tokens, dim)})})
        if prefix_tokens:
            # peft PrefixEncoder without projection: Embedding(num_virtual_tokens, num_layers * 2 * token_dim); row t holds,
            # layer by layer, the key then the value of virtual token t (peft's view(.., layers * 2, heads, head_dim))
            cfg = model.config
            if cfg.is_encoder_decoder:
                # seq2seq (T5): peft sizes the table for num_transformer_submodules = 2 (2 x 20 rows) but get_prompt only
                # ever reads the first num_virtual_tokens rows (decoder self-attention prefix, see lm._t5_stack)
                self.prefix_layers, self.prefix_dim = cfg.num_decoder_layers, cfg.num_heads * cfg.d_kv
                rows = 2 * prefix_tokens
            else:
                self.prefix_layers, self.prefix_dim = cfg.num_hidden_layers, cfg.hidden_size
                rows = prefix_tokens
            self.prompt_encoder = nn.ModuleDict({"default": nn.ModuleDict(
                {"embedding": nn.Embedding(rows, self.prefix_layers * 2 * self.prefix_dim)})})

    def get_input_embeddings(self):
        return self.base_model.model.get_input_embeddings()

    def forward(self, input_ids=None, attention_mask=None, inputs_embeds=None, labels=None, **kw):
        lm = self.base_model.model
        if self.num_prefix_tokens:    # prefix tuning: per-layer K / V of the virtual tokens in front of every layer's keys
            if not lm_kernels.supports(lm):
                raise NotImplementedError("prefix tuning needs a language model the package's kernels can run")
            w = self.prompt_encoder["default"]["embedding"].weight[: self.num_prefix_tokens]
            prefix = w.view(self.num_prefix_tokens, self.prefix_layers, 2, self.prefix_dim)
            fwd = lm_kernels.t5_forward if lm.config.is_encoder_decoder else lm_kernels.opt_forward
            return fwd(lm, input_ids=input_ids, attention_mask=attention_mask, inputs_embeds=inputs_embeds, labels=labels,
                       prefix_kv=prefix)
        if self.num_virtual_tokens:   # prompt tuning: learned embeddings prepended to the (encoder) input
            if inputs_embeds is None:
                inputs_embeds = lm.get_input_embeddings()(input_ids)
                input_ids = None
            b = inputs_embeds.shape[0]
            prompt = self.prompt_encoder["default"]["embedding"].weight.to(inputs_embeds.dtype)[None].expand(b, -1, -1)
            inputs_embeds = torch.cat((prompt, inputs_embeds), 1)
            ones = torch.ones((b, self.num_virtual_tokens), dtype=attention_mask.dtype, device=attention_mask.device)
            attention_mask = torch.cat((ones, attention_mask), 1)
            if labels is not None and not lm.config.is_encoder_decoder:
                pad = torch.full((b, self.num_virtual_tokens), -100, dtype=labels.dtype, device=labels.device)
                labels = torch.cat((pad, labels), 1)
        return run_language_model(lm, input_ids=input_ids, attention_mask=attention_mask,
                                  inputs_embeds=inputs_embeds, labels=labels, **kw)


def run_language_model(lm, **kw):
    """The LM call of the concat path (model/modelling_self_attention.py:246, :261, :280, :332): the layer stack of the
    HF T5 / OPT model runs on this package's kernels (mmgl_b200/lm.py).  The peft shim forwards to its base model through
    this function.  A model the kernels cannot run, or a call without labels (the training forward computes the loss
    in-kernel), raises: there is no HF / eager fallback."""
    if isinstance(lm, _PeftShim):
        return lm(**kw)
    if not lm_kernels.supports(lm):
        raise NotImplementedError(
            f"mmgl_b200 cannot run {type(lm).__name__} with this configuration on its kernels (supported: HF T5 with a ReLU "
            f"FFN and OPT, head_dim 64 or 128, no LayerDrop / attention dropout for OPT); there is no library fallback")
    if kw.get("labels") is None:
        raise NotImplementedError("mmgl_b200 runs the training forward (loss computed in-kernel): labels are required")
    return lm_kernels.forward(lm, **kw)


def apply_lora(model: nn.Module, r: int, alpha: float, dropout: float) -> int:
    """Wrap every attention q / v projection of a HF T5 or OPT model in LoRALinear (intent of
    model/modelling_self_attention.py:79-87; D7).  Returns the number of adapted linears."""
    targets = ("q", "v", "q_proj", "v_proj")
    n = 0
    for parent in list(model.modules()):
        for name, child in list(parent.named_children()):
            if name in targets and isinstance(child, nn.Linear):
                setattr(parent, name, LoRALinear(child, r, alpha, dropout))
                n += 1
    return n


class SelfAttentionModel(nn.Module, _NeighborEncoderMixin):
    """Drop-in for the reference's SelfAttentionModel (constructor args and forward kwargs identical)."""

    def __init__(self, args, tokenizer=None):
        super().__init__()
        self.args = args
        self.context = args.context
        self.decoder_only = args.decoder_only
        self.neighbor_mode = args.neighbor_mode
        self.position_type = args.position_type
        self.n_text_tokens = args.n_text_tokens
        self.n_visual_tokens = args.n_visual_tokens
        self.tokenizer = tokenizer

        name = str(args.model_name_or_path if isinstance(args.model_name_or_path, str) else
                   type(args.model_name_or_path).__name__).lower()
        if "t5" in name:
            model = _load_or_init("lm", args.model_name_or_path, "T5ForConditionalGeneration")
        elif "opt" in name:
            model = _load_or_init("lm", args.model_name_or_path, "OPTForCausalLM")
        else:
            raise ValueError(f"SelfAttentionModel does not support {args.model_name_or_path}.")
        if args.peft_type == "none":
            self.lm = model
        elif args.peft_type == "lora":
            for p in model.parameters():
                p.requires_grad = False
            apply_lora(model, args.lora_r, args.lora_alpha, args.lora_dropout)
            head = model.get_output_embeddings()           # modules_to_save=["lm_head"]: a trainable, untied copy
            head.weight = nn.Parameter(head.weight.detach().clone())
            self.lm = _PeftShim(model)
        elif args.peft_type == "prompt":
            for p in model.parameters():
                p.requires_grad = False
            self.lm = _PeftShim(model, prompt_tokens=20)
        elif args.peft_type == "prefix":                   # :88-92 PrefixTuningConfig(num_virtual_tokens=20)
            for p in model.parameters():
                p.requires_grad = False
            self.lm = _PeftShim(model, prefix_tokens=20)
        else:
            raise ValueError(f"SelfAttentionModel does not support {args.peft_type}.")
        self.input_embeddings = self.lm.get_input_embeddings()
        h = self.input_embeddings.embedding_dim

        with_pos = args.position_type != "none"
        self._build_encoders(args, h, with_text=self.neighbor_mode == "embedding",
                             with_visual=self.context in _ALL, with_pos=with_pos)
        if self.position_type == "laplacian":
            if self.context in ("section_only", "text_only") or self.neighbor_mode == "raw":
                raise ValueError(f"[Laplacian PE] neighbor mode: {self.neighbor_mode} and context: {self.context} are not supported.")
            k = 1 + args.max_text_neighbors + args.max_image_neighbors - 5
            self.lpe_embeddings = nn.Linear(k, h * args.n_text_tokens)
        if self.position_type == "gnn":
            d = h * args.n_text_tokens
            self.gnn = GCN(input_dim=d, output_dim=d, hidden_dim=self.text_model.config.hidden_size)
        if getattr(args, "freeze_lm", False):
            self.lm.eval()
            for p in self.lm.parameters():
                p.requires_grad = False

    def train(self, mode=True):
        super().train(mode)
        self._freeze_modes()
        return self

    def _run_lm(self, **kw):
        return run_language_model(self.lm, **kw)

    def forward(self, input_ids, attention_mask, labels, images=None, image_positions=None, neighbor_input_ids=None,
                neighbor_attention_mask=None, neighbor_pos_ids=None, text_locations=None, neighbor_images=None,
                neighbor_images_pos_ids=None, image_locations=None, lpe=None, graph=None, neighbor_plan=None):
        if self.neighbor_mode == "raw" and self.context in _TEXT_ONLY:                               # :244-246
            return self._run_lm(input_ids=input_ids, attention_mask=attention_mask, labels=labels)
        if self.neighbor_mode == "raw" and self.context in _ALL:                                    # :248-261
            embs = self.input_embeddings(input_ids).clone()
            vis = self.visual_projection(images).to(embs.dtype)
            b, _, h = embs.shape
            bidx = torch.arange(b, device=embs.device)[:, None]
            embs[bidx, image_positions] = vis.reshape(b, -1, h)
            if self.decoder_only:
                labels = labels.clone()
                labels[bidx, image_positions] = -100
            return self._run_lm(inputs_embeds=embs, attention_mask=attention_mask, labels=labels)
        if self.neighbor_mode != "embedding" or self.context not in _TEXT_ONLY + _ALL:
            raise ValueError(f"Neighbor mode: {self.neighbor_mode} and context: {self.context} are not supported.")

        use_pos = self.position_type != "none"
        if self.context in _TEXT_ONLY:                                                              # :263-280
            bank, mask = self.build_bank(neighbor_input_ids, neighbor_attention_mask, neighbor_pos_ids, None,
                                         use_pos_tables=use_pos, plan=neighbor_plan)
        else:                                                                                       # :282-320
            is_all = self.context == "all"
            bank, mask = self.build_bank(neighbor_input_ids, neighbor_attention_mask, neighbor_pos_ids, text_locations,
                                         neighbor_images, neighbor_images_pos_ids, image_locations,
                                         lpe=lpe if (is_all and self.position_type == "laplacian") else None,
                                         use_pos_tables=use_pos, plan=neighbor_plan)
            if is_all and self.position_type == "gnn":
                b, nk, h = bank.shape
                flat = bank.reshape(b, nk // self.n_text_tokens, self.n_text_tokens * h)
                bank = (flat + self.gnn(flat, graph)).reshape(b, nk, h)
        embs = self.input_embeddings(input_ids)                                                     # :323-330
        embs = torch.cat((embs.to(bank.dtype), bank), dim=1)
        attention_mask = torch.cat((attention_mask, mask.to(attention_mask.dtype)), dim=1)
        if self.decoder_only:
            pad = torch.full((bank.shape[0], bank.shape[1]), -100, dtype=labels.dtype, device=labels.device)
            labels = torch.cat((labels, pad), dim=1)
        return self._run_lm(inputs_embeds=embs, attention_mask=attention_mask, labels=labels)
