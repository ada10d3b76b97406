"""Import and drive the byte-compiled REAL reference (``oracle/_ref``, built by oracle/build_ref.py) -- TEST / BASELINE
INFRASTRUCTURE ONLY; nothing in the product package imports this file.

The reference modules run UNMODIFIED.  What the harness supplies around them (the reference needs a network and a
dataset that do not exist here):
  * ``from_pretrained``: the names ``AutoConfig / AutoModelForCausalLM / RobertaModel / CLIPVisionModel`` in the loaded
    module's namespace are bound to shims that return RANDOM-INITIALISED HF models of the named architecture
    (mmgl_b200/configs.py holds the published hyper-parameters) -- the same weights policy as the GPU arm;
  * ``args``: a namespace with ``neighbor_layer_wise`` and ``neighbor_mode="cross_attention"`` (reference defects D1 / D2,
    SURVEY section 0: the CLI cannot produce a working namespace for this path by itself);
  * ``lm_head`` is frozen (D12: under transformers 5.x it comes out untied and trainable, which the 4.x-era reference
    never intended; leaving it trainable would only slow the baseline down);
  * bf16 (GPU baseline only): ``model.bfloat16()`` as run_generation.py:306-307 does, with the default dtype switched to
    bf16 around the forward so that the ``torch.zeros`` bank of :1095 matches (defect D4: the stock bf16 path raises).
"""
from __future__ import annotations

import contextlib
import importlib.machinery
import importlib.util
import json
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


def available() -> bool:
    return os.path.exists(os.path.join(REF_DIR, "model", "modelling_cross_attention.bytecode"))


def manifest() -> dict:
    with open(os.path.join(REF_DIR, "MANIFEST.json")) as f:
        return json.load(f)


def _load_pyc(name: str, rel: str):
    path = os.path.join(REF_DIR, rel)
    loader = importlib.machinery.SourcelessFileLoader(name, path)
    spec = importlib.util.spec_from_loader(name, loader, origin=path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod   # HF PreTrainedModel.__init__ looks cls.__module__ up in sys.modules
    loader.exec_module(mod)
    return mod


_cache = {}


def cross_attention_module():
    if "xa" not in _cache:
        _cache["xa"] = _load_pyc("mmgl_ref_xattn", "model/modelling_cross_attention.bytecode")
    return _cache["xa"]


class _Pretrained:
    """``X.from_pretrained(name, ...)`` -> random-init model / config of the named architecture."""

    def __init__(self, make):
        self._make = make

    def from_pretrained(self, name, *a, **kw):
        return self._make(name, kw.get("config"))


def _install_shims(xa):
    from transformers import CLIPVisionModel, OPTForCausalLM, RobertaModel
    from mmgl_b200 import configs
    xa.AutoConfig = _Pretrained(lambda name, cfg: configs.lm_config(name))
    xa.AutoModelForCausalLM = _Pretrained(lambda name, cfg: OPTForCausalLM(cfg if cfg is not None else configs.lm_config(name)))
    xa.RobertaModel = _Pretrained(lambda name, cfg: RobertaModel(configs.text_config(name)))
    xa.CLIPVisionModel = _Pretrained(lambda name, cfg: CLIPVisionModel(configs.visual_config(name)))


def reference_args(w: dict, n_cross_layers: int = 4):
    """argparse-like namespace for the reference's CrossAttentionModel at workload ``w`` (bench.py WORKLOADS entry)."""
    from mmgl_b200 import configs
    layers = configs.lm_config(w["lm"]).num_hidden_layers
    return types.SimpleNamespace(
        context="all", neighbor_mode="cross_attention", peft_type="flamingo", n_text_tokens=4, n_visual_tokens=4,
        model_name_or_path=w["lm"], text_model=w["text"], visual_model=w["visual"], max_output_length=w["s_out"],
        freeze_lm=False, neighbor_layer_wise=max(1, layers // n_cross_layers), lora_r=64, lora_alpha=1, lora_dropout=0.0)


def build_cross_attention_model(w: dict, seed: int = 0, gate: float = 0.5):
    """The reference's CrossAttentionModel (fp32, CPU), random-init, gates live, lm_head frozen; train mode."""
    xa = cross_attention_module()
    _install_shims(xa)
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(sys.stderr):   # the constructor prints one line per copied layer
        model = xa.CrossAttentionModel(reference_args(w), tokenizer=None)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if "gating" in n:
                p.fill_(gate)
    model.lm.lm_head.weight.requires_grad = False
    model.train()
    return model


@contextlib.contextmanager
def default_dtype(dtype):
    old = torch.get_default_dtype()
    torch.set_default_dtype(dtype)
    try:
        yield
    finally:
        torch.set_default_dtype(old)


def train_step(model, batch, optimizer, dtype=torch.float32):
    """The body of the reference's train_loop (language_modelling/run_generation.py:462-494) at accumulation 1."""
    keys = ("input_ids", "attention_mask", "labels", "neighbor_input_ids", "neighbor_attention_mask", "neighbor_pos_ids",
            "text_locations", "neighbor_images", "neighbor_images_pos_ids", "image_locations")
    kw = {k: batch[k] for k in keys}
    if dtype != torch.float32:
        kw["neighbor_images"] = kw["neighbor_images"].to(dtype)
    with default_dtype(dtype):
        out = model(**kw)
    loss = out.loss
    loss.backward()
    optimizer.step()
    optimizer.zero_grad(set_to_none=True)
    return loss
