"""Architecture hyper-parameters of the public checkpoints BASELINE.json names, as HF config objects.

There is no network and no HF cache in the build/bench images, so with ``MMGL_ALLOW_RANDOM_INIT=1`` (bench.py, tests)
models are random-initialised from these configs (values are the published ``config.json`` of each checkpoint).
Without that opt-in ``args.model_name_or_path`` etc. go through ``from_pretrained`` as in the reference
(model/modelling_cross_attention.py:953-954) and a missing checkpoint is an error (modules._load_or_init).
"""
from __future__ import annotations

import os


def _opt(hidden, layers, heads, ffn, proj=None, pre_ln=True):
    from transformers import OPTConfig
    return OPTConfig(vocab_size=50272, hidden_size=hidden, num_hidden_layers=layers, ffn_dim=ffn,
                     num_attention_heads=heads, max_position_embeddings=2048, word_embed_proj_dim=proj or hidden,
                     do_layer_norm_before=pre_ln, dropout=0.1, attention_dropout=0.0, activation_function="relu",
                     pad_token_id=1, bos_token_id=2, eos_token_id=2)


def lm_config(name: str):
    key = os.path.basename(name.rstrip("/")).lower().replace("mpt", "opt")
    if key == "opt-125m":
        return _opt(768, 12, 12, 3072)
    if key == "opt-350m":
        return _opt(1024, 24, 16, 4096, proj=512, pre_ln=False)
    if key == "opt-1.3b":
        return _opt(2048, 24, 32, 8192)
    # opt-2.7b (head_dim 80) is not listed: the attention kernels cover head_dim 64 / 128 only
    if key == "opt-6.7b":
        return _opt(4096, 32, 32, 16384)
    if key in ("llama-2-7b", "llama-2-7b-hf"):
        from transformers import LlamaConfig
        return LlamaConfig(vocab_size=32000, hidden_size=4096, intermediate_size=11008, num_hidden_layers=32,
                           num_attention_heads=32, num_key_value_heads=32, max_position_embeddings=4096, rms_norm_eps=1e-5,
                           rope_theta=10000.0, hidden_act="silu", tie_word_embeddings=False, pad_token_id=0, bos_token_id=1,
                           eos_token_id=2)
    if key in ("t5-base", "t5-small", "t5-large"):
        from transformers import T5Config
        dims = {"t5-small": (512, 64, 2048, 6, 8), "t5-base": (768, 64, 3072, 12, 12), "t5-large": (1024, 64, 4096, 24, 16)}
        d_model, d_kv, d_ff, layers, heads = dims[key]
        return T5Config(vocab_size=32128, d_model=d_model, d_kv=d_kv, d_ff=d_ff, num_layers=layers,
                        num_decoder_layers=layers, num_heads=heads, relative_attention_num_buckets=32,
                        dropout_rate=0.1, feed_forward_proj="relu", decoder_start_token_id=0, pad_token_id=0,
                        eos_token_id=1)
    raise KeyError(f"no built-in config for {name!r}")


def text_config(name: str):
    key = os.path.basename(name.rstrip("/")).lower()
    from transformers import RobertaConfig
    if key == "roberta-base":
        return RobertaConfig(vocab_size=50265, hidden_size=768, num_hidden_layers=12, num_attention_heads=12,
                             intermediate_size=3072, max_position_embeddings=514, type_vocab_size=1,
                             layer_norm_eps=1e-5, pad_token_id=1, bos_token_id=0, eos_token_id=2)
    if key == "roberta-large":
        return RobertaConfig(vocab_size=50265, hidden_size=1024, num_hidden_layers=24, num_attention_heads=16,
                             intermediate_size=4096, max_position_embeddings=514, type_vocab_size=1,
                             layer_norm_eps=1e-5, pad_token_id=1, bos_token_id=0, eos_token_id=2)
    raise KeyError(f"no built-in config for {name!r}")


def visual_config(name: str):
    key = os.path.basename(name.rstrip("/")).lower()
    from transformers import CLIPVisionConfig
    if key == "clip-vit-base-patch16":
        return CLIPVisionConfig(hidden_size=768, intermediate_size=3072, num_hidden_layers=12, num_attention_heads=12,
                                image_size=224, patch_size=16, hidden_act="quick_gelu", projection_dim=512)
    if key == "clip-vit-base-patch32":
        return CLIPVisionConfig(hidden_size=768, intermediate_size=3072, num_hidden_layers=12, num_attention_heads=12,
                                image_size=224, patch_size=32, hidden_act="quick_gelu", projection_dim=512)
    if key == "clip-vit-large-patch14":
        return CLIPVisionConfig(hidden_size=1024, intermediate_size=4096, num_hidden_layers=24, num_attention_heads=16,
                                image_size=224, patch_size=14, hidden_act="quick_gelu", projection_dim=768)
    raise KeyError(f"no built-in config for {name!r}")
