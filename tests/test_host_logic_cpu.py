"""Host-side logic of the product path that needs no GPU: index / bucket arithmetic that the CUDA kernels are driven by.
Integer work: bit-exact.  (The kernels themselves are covered by the `-m gpu` tests.)"""
import types

import pytest
import torch


def test_t5_bucket_matches_hf_relative_position_bucket():
    """lm._t5_bucket restates T5Attention._relative_position_bucket (HF models/t5/modeling_t5.py:189-233)."""
    from transformers.models.t5.modeling_t5 import T5Attention
    from mmgl_b200 import lm as L
    rel = torch.arange(-700, 701)
    for bidirectional in (True, False):
        for nb, md in ((32, 128), (16, 64)):
            want = T5Attention._relative_position_bucket(rel, bidirectional=bidirectional, num_buckets=nb, max_distance=md)
            got = L._t5_bucket(rel, bidirectional, nb, md)
            assert torch.equal(got, want), (bidirectional, nb, md)


@pytest.mark.parametrize("decoder", [False, True])
def test_t5_rel_bias_vector_is_the_dense_bias_by_distance(decoder):
    """t5_rel_bias returns [heads, sq + sk - 1] with entry d + sq - 1 = bias at distance d = key - query; expanding it
    must give T5Attention.compute_bias (:236-251) exactly, and it stays differentiable for a trainable table."""
    from transformers import T5Config
    from transformers.models.t5.modeling_t5 import T5Attention
    from mmgl_b200 import lm as L
    torch.manual_seed(0)
    cfg = T5Config(d_model=64, d_kv=16, num_heads=4, is_decoder=decoder)
    cfg.is_decoder = decoder
    attn = T5Attention(cfg, has_relative_attention_bias=True, layer_idx=0)
    with torch.no_grad():
        attn.relative_attention_bias.weight.normal_()
    sq, sk = 37, 37
    dense = attn.compute_bias(sq, sk)[0]                      # [heads, sq, sk]
    vec = L.t5_rel_bias(attn, sq, sk)
    assert vec.shape == (4, sq + sk - 1) and vec.requires_grad
    idx = torch.arange(sk)[None, :] - torch.arange(sq)[:, None] + sq - 1
    assert torch.equal(vec[:, idx], dense)
    vec.sum().backward()
    assert attn.relative_attention_bias.weight.grad is not None
    attn.relative_attention_bias.weight.requires_grad_(False)
    assert not L.t5_rel_bias(attn, sq, sk).requires_grad


def test_pack_plan_drops_right_padding_only():
    """encoders._pack_plan (frozen RoBERTa on real tokens only; data.py:457 right-pads every neighbor): prefix-form masks
    give the flat indices of the real tokens and their cumulative lengths; anything else (left padding, holes, an empty
    row, < 10% padding) must refuse, so the dense path runs."""
    from mmgl_b200.encoders import _pack_plan
    lens = torch.tensor([5, 12, 1, 9])
    s = 12
    am = (torch.arange(s)[None, :] < lens[:, None]).long()
    idx, cu, total, longest = _pack_plan(am)
    assert total == int(lens.sum()) and longest == 12
    assert cu.dtype == torch.int32 and cu.tolist() == [0, 5, 17, 18, 27]
    assert torch.equal(idx, torch.nonzero(am.reshape(-1)).squeeze(1))
    hole = am.clone()
    hole[1, 3] = 0
    assert _pack_plan(hole) is None
    left = am.flip(1)
    assert _pack_plan(left) is None
    empty = am.clone()
    empty[2] = 0
    assert _pack_plan(empty) is None
    assert _pack_plan(torch.ones(4, s, dtype=torch.long)) is None          # nothing to save


def test_lm_output_is_indexable_like_hf_outputs():
    """run_generation.py reads outputs.loss / outputs[1] (:466-474)."""
    from mmgl_b200.lm import LMOutput
    loss, logits = torch.tensor(1.5), torch.zeros(2, 3)
    out = LMOutput(loss=loss, logits=logits)
    assert out.loss is loss and out["logits"] is logits and out[0] is loss and out[1] is logits
    assert LMOutput(logits=logits)[0] is logits


def test_apply_lora_adapts_q_and_v_of_every_attention_and_keeps_peft_key_names():
    """Intent of model/modelling_self_attention.py:79-87: rank-r adapters on the q and v projections of every
    attention of the LM, base weights frozen, B = 0 at init (the adapted model starts as the base model), state-dict
    keys in peft's layout (`<path>.base_layer.weight`, `<path>.lora_A.default.weight`, `<path>.lora_B.default.weight`)."""
    from transformers import OPTConfig, OPTForCausalLM, T5Config, T5ForConditionalGeneration
    from mmgl_b200.self_attention import LoRALinear, _PeftShim, apply_lora
    t5 = T5ForConditionalGeneration(T5Config(vocab_size=64, d_model=32, d_kv=8, d_ff=64, num_layers=2,
                                             num_decoder_layers=2, num_heads=4, decoder_start_token_id=0))
    n = apply_lora(t5, 4, 1.0, 0.0)
    assert n == 2 * (2 + 2 + 2)                       # encoder self, decoder self, decoder cross: q and v each, 2 layers
    opt = OPTForCausalLM(OPTConfig(vocab_size=64, hidden_size=32, ffn_dim=64, num_hidden_layers=3, num_attention_heads=4,
                                   max_position_embeddings=32, word_embed_proj_dim=32))
    assert apply_lora(opt, 4, 1.0, 0.0) == 2 * 3
    for model in (t5, opt):
        adapted = [(k, m) for k, m in model.named_modules() if isinstance(m, LoRALinear)]
        for name, m in adapted:
            assert name.rsplit(".", 1)[-1] in ("q", "v", "q_proj", "v_proj")
            assert not any(p.requires_grad for p in m.base_layer.parameters())
            assert m.lora_A["default"].weight.shape == (4, m.base_layer.in_features)
            assert float(m.lora_B["default"].weight.detach().abs().max()) == 0.0
            assert m.scaling == 0.25
        keys = _PeftShim(model).state_dict().keys()
        name = adapted[0][0]
        for suffix in ("base_layer.weight", "lora_A.default.weight", "lora_B.default.weight"):
            assert f"base_model.model.{name}.{suffix}" in keys


def test_peft_shim_prompt_and_prefix_tables_have_pefts_shapes():
    """Prompt tuning: Embedding(20, hidden); prefix tuning (OPT): Embedding(20, layers * 2 * hidden), both under
    `prompt_encoder.default.embedding` (peft's PromptEmbedding / PrefixEncoder without projection); T5 (seq2seq) prefix
    tuning: Embedding(2 * 20, decoder layers * 2 * heads * d_kv) as peft sizes it for num_transformer_submodules = 2."""
    from transformers import OPTConfig, OPTForCausalLM, T5Config, T5ForConditionalGeneration
    from mmgl_b200.self_attention import _PeftShim
    opt = OPTForCausalLM(OPTConfig(vocab_size=64, hidden_size=32, ffn_dim=64, num_hidden_layers=3, num_attention_heads=4,
                                   max_position_embeddings=32, word_embed_proj_dim=32))
    assert _PeftShim(opt, prompt_tokens=20).state_dict()["prompt_encoder.default.embedding.weight"].shape == (20, 32)
    assert _PeftShim(opt, prefix_tokens=20).state_dict()["prompt_encoder.default.embedding.weight"].shape == (20, 3 * 2 * 32)
    t5 = T5ForConditionalGeneration(T5Config(vocab_size=64, d_model=32, d_kv=8, d_ff=64, num_layers=1,
                                             num_decoder_layers=3, num_heads=4, decoder_start_token_id=0))
    assert _PeftShim(t5, prefix_tokens=20).state_dict()["prompt_encoder.default.embedding.weight"].shape == (40, 3 * 2 * 32)


def test_lm_support_matrix():
    """lm.supports: HF T5 (ReLU FFN, head dim 64 / 128) and OPT run on the package's kernels; gated-GELU T5 and other
    head dims go to the HF forward."""
    from transformers import OPTConfig, OPTForCausalLM, T5Config, T5ForConditionalGeneration
    from mmgl_b200 import lm as L

    def t5(**kw):
        base = dict(vocab_size=64, d_model=128, d_kv=64, d_ff=64, num_layers=1, num_decoder_layers=1, num_heads=2,
                    decoder_start_token_id=0)
        base.update(kw)
        return T5ForConditionalGeneration(T5Config(**base))

    assert L.supports(t5())
    assert L.supports(t5(d_kv=128))
    assert not L.supports(t5(d_kv=32))
    assert not L.supports(t5(feed_forward_proj="gated-gelu"))
    opt = OPTForCausalLM(OPTConfig(vocab_size=64, hidden_size=128, ffn_dim=64, num_hidden_layers=1, num_attention_heads=2,
                                   max_position_embeddings=32, word_embed_proj_dim=128))
    assert L.supports(opt)
    assert not L.supports(torch.nn.Linear(4, 4))


def test_cfg2_interleave_positions_and_trainable_parameter_count():
    """MPTDecoder interleave loop (model/modelling_cross_attention.py:576-627) at the cfg2 dimensions (OPT-1.3B, 4
    neighbor layers), built on the meta device: one gated cross-attention layer follows each of OPT layers 5 / 11 / 17 /
    23, only they train (`mark_only_peft_as_trainable`, :731-737), lm_head is tied to the embedding (D12), and
    the gradient set DDP all-reduces adds up to the 216,736,520 parameters SURVEY 8 (a11) counts."""
    from transformers import OPTConfig
    from mmgl_b200 import modules as M
    opt_cfg = OPTConfig(vocab_size=50272, hidden_size=2048, ffn_dim=8192, num_hidden_layers=24, num_attention_heads=32,
                        max_position_embeddings=2048, word_embed_proj_dim=2048)
    args = types.SimpleNamespace(neighbor_layer_wise=None, num_neighbor_layers=4, neighbor_mode="cross_attention",
                                 peft_type="flamingo", lora_r=64, lora_alpha=1, lora_dropout=0.0)
    with torch.device("meta"):
        model = M.MPTForCausalLM(M.MPTConfig(args, opt_cfg))
    layers = [m for m in model.modules() if isinstance(m, M.MPTDecoderLayer)]
    cross = [m for m in layers if m.cross_attention]
    assert len(layers) - len(cross) == 24 and len(cross) == 4
    assert model.lm_head.weight is model.model.decoder.embed_tokens.weight
    trainable = sum(p.numel() for p in model.parameters() if p.requires_grad)
    # 4 gated layers x 50,358,274 ("50.36 M params / layer"); the wrapper adds the two neighbor projections
    # Linear(768 -> 4 x 2048), their two Embedding(129, 8192) position tables and the TextPooler dense = 15,303,424
    per_layer = sum(p.numel() for p in cross[0].parameters())
    assert per_layer == 50_358_274 and trainable == 4 * per_layer
    wrapper = 2 * (768 * 8192 + 8192) + 2 * 129 * 8192 + (768 * 768 + 768)
    assert trainable + wrapper == 216_736_520
    assert all(p.requires_grad for m in cross for p in m.parameters())
    assert not any(p.requires_grad for m in layers if not m.cross_attention for p in m.parameters())
    # D1: the CLI defines only num_neighbor_layers; one gated layer follows every 24 / 4 = 6th OPT layer (indices 5..23)
    dec = model.model.decoder
    assert dec.neighbor_layer_wise == 6 and len(dec.layers) == 24 and len(dec.neighbor_layers) == 4
    assert [i for i in range(24) if (i + 1) % dec.neighbor_layer_wise == 0] == [5, 11, 17, 23]


def test_pretrained_names_are_not_silently_random_initialised(monkeypatch):
    """a hub name whose weights are absent must RAISE (the reference always calls from_pretrained,
    model/modelling_cross_attention.py:953-954); random init only behind MMGL_ALLOW_RANDOM_INIT=1, with a warning."""
    from mmgl_b200 import modules
    monkeypatch.delenv("MMGL_ALLOW_RANDOM_INIT", raising=False)
    with pytest.raises(RuntimeError, match="MMGL_ALLOW_RANDOM_INIT"):
        modules._load_or_init("lm", "facebook/opt-125m", "OPTForCausalLM")
    monkeypatch.setenv("MMGL_ALLOW_RANDOM_INIT", "1")
    with pytest.warns(RuntimeWarning, match="RANDOM-INITIALISED"):
        m = modules._load_or_init("text", "roberta-base", "RobertaModel")
    assert m.config.hidden_size == 768


def test_pretrained_directory_is_loaded(tmp_path):
    """a local save_pretrained directory loads its weights (reference behaviour), no opt-in needed"""
    from transformers import OPTConfig, OPTForCausalLM
    from mmgl_b200 import modules
    cfg = OPTConfig(vocab_size=64, hidden_size=64, num_hidden_layers=1, ffn_dim=64, num_attention_heads=1,
                    max_position_embeddings=32, word_embed_proj_dim=64)
    src = OPTForCausalLM(cfg)
    src.save_pretrained(tmp_path)
    got = modules._load_or_init("lm", str(tmp_path), "OPTForCausalLM")
    assert torch.equal(got.model.decoder.layers[0].fc1.weight, src.model.decoder.layers[0].fc1.weight)


def test_unsupported_head_dim_is_rejected_at_construction():
    """head_dim 80 (opt-2.7b) can never run on the kernels -> constructor error, not a first-forward error."""
    from mmgl_b200 import modules, ops
    cfg = types.SimpleNamespace(hidden_size=160, num_attention_heads=2, enable_bias=True, peft_type="flamingo")
    with pytest.raises(ValueError, match="head_dim 64 and 128"):
        modules.MPTAttention(cfg, cross_attention=True)
    assert ops.xattn_max_keys(64) == 256 and ops.xattn_max_keys(128) == 128 and ops.xattn_max_keys(80) == 0
    with pytest.raises(KeyError):
        from mmgl_b200 import configs
        configs.lm_config("facebook/opt-2.7b")


def test_weight_shadow_cache_sees_data_writes_after_invalidate():
    """p.data.add_() does not bump p._version; ops.invalidate_weight_cache() is the documented remedy (called
    by train.optimizer_step), while version-bumping updates refresh by themselves and a re-pointed .data is noticed."""
    from mmgl_b200 import ops
    p = torch.nn.Parameter(torch.ones(4, 4))
    a = ops.w16(p)
    assert ops.w16(p) is a                               # cached
    with torch.no_grad():
        p.add_(1.0)                                      # bumps _version
    b = ops.w16(p)
    assert float(b[0, 0]) == 2.0
    p.data.add_(1.0)                                     # silent: same version, same storage
    assert ops.w16(p) is b
    ops.invalidate_weight_cache(trainable_only=True)
    assert float(ops.w16(p)[0, 0]) == 3.0
    p.data = torch.full((4, 4), 7.0)                     # new storage -> stamp changes
    assert float(ops.w16(p)[0, 0]) == 7.0
    frozen = torch.nn.Parameter(torch.ones(2, 2), requires_grad=False)
    f = ops.fused_rows([frozen, frozen], torch.bfloat16)
    ops.invalidate_weight_cache(trainable_only=True)
    assert ops.fused_rows([frozen, frozen], torch.bfloat16) is f   # frozen entries survive an optimizer step
    ops.invalidate_weight_cache()
    assert ops.fused_rows([frozen, frozen], torch.bfloat16) is not f


def test_gnn_disables_the_padding_neighbor_skip():
    """with position_type == 'gnn' the GCN mixes bank rows before the mask acts, so padding neighbors keep their
    W*enc+b rows (reference semantics) instead of zeros."""
    from mmgl_b200.modules import _NeighborEncoderMixin
    m = _NeighborEncoderMixin()
    pos = torch.tensor([[1, 2, 0], [1, 0, 0]])
    m.position_type = "none"
    assert m._needed(pos).tolist() == [0, 1, 3]
    m.position_type = "gnn"
    assert m._needed(pos) is None


def test_host_plan_equals_the_device_side_bookkeeping():
    """mmgl_b200.plan.make_plan (data-pipeline side, no sync) must reproduce what the module derives from the device tensors:
    modules._needed (valid neighbors) and encoders._pack_plan (real-token index, cu_seqlens, totals).  Integer work: exact."""
    from mmgl_b200 import encoders, plan, synth
    from mmgl_b200.modules import _NeighborEncoderMixin
    for seed in range(6):
        b = synth.make_batch(synth.BatchSpec(batch=5, max_input_length=64, max_output_length=16, text_neighbors=6,
                                             image_neighbors=3, image_size=8), seed=seed)
        p = plan.make_plan(b)
        m = _NeighborEncoderMixin()
        m.position_type = "none"
        idx = m._needed(b["neighbor_pos_ids"])
        want_idx = idx if idx is not None else torch.arange(b["neighbor_pos_ids"].numel())
        assert torch.equal(p.text_idx, want_idx)
        iidx = m._needed(b["neighbor_images_pos_ids"])
        assert torch.equal(p.image_idx, iidx if iidx is not None else torch.arange(b["neighbor_images_pos_ids"].numel()))
        am2 = b["neighbor_attention_mask"].reshape(-1, 64).index_select(0, want_idx)
        dev = encoders._pack_plan(am2)
        assert (dev is not None) == p.pack
        if dev is not None:
            tok, cu, total, max_len = dev
            assert torch.equal(tok, p.tok_idx) and torch.equal(cu, p.cu) and (total, max_len) == (p.total, p.max_len)
    moved = p.to("cpu").pin_memory() if torch.cuda.is_available() else p.to("cpu")
    assert moved.total == p.total and torch.equal(moved.text_idx, p.text_idx)


def test_fused_adamw_has_no_cpu_fallback():
    """optim.FusedAdamW is a CUDA-kernel optimizer: a CPU parameter must raise, not run an eager update."""
    import pytest
    import torch
    from mmgl_b200.optim import FusedAdamW
    p = torch.nn.Parameter(torch.randn(4))
    opt = FusedAdamW([p], lr=1e-3)
    p.grad = torch.randn(4)
    with pytest.raises(NotImplementedError):
        opt.step()
    with pytest.raises(NotImplementedError):
        FusedAdamW([p], amsgrad=True)


def test_optimizer_step_drops_trainable_weight_shadows():
    """torch.optim.AdamW(fused=True) updates parameters without bumping ``_version``; the bf16 shadow cache (ops.w16) must
    not survive an optimizer step of any optimizer (global post-step hook in ops.py), while frozen shadows stay cached."""
    import torch
    from mmgl_b200 import ops
    w = torch.nn.Parameter(torch.randn(8, 4))
    frozen = torch.nn.Parameter(torch.randn(8, 4), requires_grad=False)
    s0, f0 = ops.w16(w), ops.w16(frozen)
    assert ops.w16(w) is s0 and ops.w16(frozen) is f0              # cached
    opt = torch.optim.AdamW([w], lr=0.1, fused=True)
    v0 = w._version
    w.grad = torch.ones_like(w)
    opt.step()
    s1 = ops.w16(w)
    assert s1 is not s0, "stale shadow served after optimizer.step()"
    assert torch.equal(s1, w.detach().to(torch.bfloat16))
    assert ops.w16(frozen) is f0                                   # frozen operands (fused Wq|Wk|Wv rows etc.) survive
    if w._version == v0:                                            # the reason the hook exists (torch 2.11: true)
        assert not torch.equal(s0, s1)
