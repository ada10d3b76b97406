"""CPU reference training step of the neighbor-fusion path.  TEST / BASELINE INFRASTRUCTURE ONLY.

One optimisation step of the reference's ``train_loop`` body (language_modelling/run_generation.py:462-494) on host
cores in fp32: frozen HF encoders (library code, as in the reference) -> the oracle's functional restatement of
CrossAttentionModel.forward (oracle/mmgl_oracle.py, pinned to the reference by tests/test_oracle_golden.py) ->
``loss.backward()`` -> AdamW.  ``bench.py`` times it as ``cpu_baseline`` and as ``--impl reference`` (kind "port":
/root/reference itself does not exist on the GPU box, and is pure Python over the same torch/HF calls anyway).
Nothing in the product package imports this file.
"""
from __future__ import annotations

import os
import time
from typing import Dict

import torch

from . import mmgl_oracle as O


def build_cpu_reference(lm_cfg, text_cfg, visual_cfg, args, seed: int = 0):
    """Random-init fp32 CPU weights in the reference's state-dict layout + the frozen HF encoders.
    Returns (params, cfg, text_model, visual_model); trainable tensors have requires_grad=True."""
    from transformers import CLIPVisionModel, OPTForCausalLM, RobertaModel
    torch.manual_seed(seed)
    g = torch.Generator().manual_seed(seed)
    opt = OPTForCausalLM(lm_cfg).float().eval()
    text_model = RobertaModel(text_cfg).float().eval()
    visual_model = CLIPVisionModel(visual_cfg).float().eval()
    h, f = lm_cfg.hidden_size, lm_cfg.ffn_dim
    p: Dict[str, torch.Tensor] = {}
    for k, v in opt.state_dict().items():
        p["lm." + k] = v.detach()
    p["lm.lm_head.weight"] = p["lm.model.decoder.embed_tokens.weight"]
    nlw = args.neighbor_layer_wise
    n_cross = lm_cfg.num_hidden_layers // nlw
    std = lm_cfg.init_std

    def lin(out_f, in_f, prefix):
        p[prefix + ".weight"] = (torch.randn(out_f, in_f, generator=g) * std).requires_grad_(True)
        p[prefix + ".bias"] = torch.zeros(out_f, requires_grad=True)

    for i in range(n_cross):
        pre = f"lm.model.decoder.neighbor_layers.{i}."
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            lin(h, h, pre + "self_attn." + n)
        lin(f, h, pre + "fc1")
        lin(h, f, pre + "fc2")
        for n in ("self_attn_layer_norm", "final_layer_norm"):
            p[pre + n + ".weight"] = torch.ones(h, requires_grad=True)
            p[pre + n + ".bias"] = torch.zeros(h, requires_grad=True)
        p[pre + "gating1"] = torch.tensor(0.5, requires_grad=True)   # live gates (0.0 at the reference init kills the branch)
        p[pre + "gating2"] = torch.tensor(0.5, requires_grad=True)
    e_t, e_v = text_cfg.hidden_size, visual_cfg.hidden_size
    n_tok = args.n_text_tokens
    lin(e_t, e_t, "text_pooler.dense")
    lin(n_tok * h, e_t, "text_embeddings")
    lin(n_tok * h, e_v, "visual_embeddings")
    p["text_position_embeddings.weight"] = (torch.randn(args.max_output_length + 1, n_tok * h, generator=g)).requires_grad_(True)
    p["visual_position_embeddings.weight"] = (torch.randn(args.max_output_length + 1, n_tok * h, generator=g)).requires_grad_(True)
    cfg = dict(num_heads=lm_cfg.num_attention_heads, num_layers=lm_cfg.num_hidden_layers, neighbor_layer_wise=nlw,
               do_layer_norm_before=lm_cfg.do_layer_norm_before, n_tokens=n_tok, flamingo=True)
    return p, cfg, text_model, visual_model


def train_step(p, cfg, batch, text_model, visual_model, optimizer) -> float:
    b, t, l = batch["neighbor_input_ids"].shape
    i = batch["neighbor_images"].shape[1]
    with torch.no_grad():
        enc = text_model(input_ids=batch["neighbor_input_ids"].reshape(-1, l),
                         attention_mask=batch["neighbor_attention_mask"].reshape(-1, l)).last_hidden_state
        vis = visual_model(batch["neighbor_images"].reshape(-1, *batch["neighbor_images"].shape[2:])).pooler_output
    pooled = O.text_pooler(enc, p["text_pooler.dense.weight"], p["text_pooler.dense.bias"])
    loss, _ = O.cross_attention_model_from_pooled(p, cfg, batch, pooled.reshape(b, t, -1), vis.reshape(b, i, -1))
    loss.backward()
    optimizer.step()
    optimizer.zero_grad(set_to_none=True)
    return float(loss.detach())


def time_steps(p, cfg, batches, text_model, visual_model, steps: int, warmup: int, threads: int = 0):
    """Returns (seconds per step, losses).  ``batches`` is a list cycled over; one batch = the bounded sample."""
    torch.set_num_threads(threads or os.cpu_count() or 1)
    params = [v for v in p.values() if v.requires_grad]
    opt = torch.optim.AdamW(params, lr=1e-4)
    losses = []
    for s in range(warmup):
        losses.append(train_step(p, cfg, batches[s % len(batches)], text_model, visual_model, opt))
    t0 = time.perf_counter()
    for s in range(steps):
        losses.append(train_step(p, cfg, batches[s % len(batches)], text_model, visual_model, opt))
    return (time.perf_counter() - t0) / max(1, steps), losses
