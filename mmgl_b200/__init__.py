"""mmgl_b200: B200 (sm_100a) kernels behind MMGL's neighbor-fusion training step, behind the reference's own
nn.Module surface (``from mmgl_b200 import CrossAttentionModel, SelfAttentionModel`` mirrors the reference's
``from model import ...``, model/__init__.py:1-2)."""

__all__ = ["CrossAttentionModel", "SelfAttentionModel", "LlamaCrossAttentionModel", "MPTDecoderLayer", "MPTForCausalLM",
           "MPTConfig", "GCN"]


def __getattr__(name):  # lazy: importing the package must not import torch / load the .so
    if name in __all__:
        from . import modules
        if name == "SelfAttentionModel":
            from . import self_attention
            return self_attention.SelfAttentionModel
        if name == "LlamaCrossAttentionModel":
            from . import llama
            return llama.LlamaCrossAttentionModel
        return getattr(modules, name)
    raise AttributeError(name)
