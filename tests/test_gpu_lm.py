"""The language model of the concat path on the package's kernels (mmgl_b200/lm.py) against the
HuggingFace modules' OWN fp32 forward / backward on the same weights -- HF T5 / OPT is what the reference's
SelfAttentionModel runs (model/modelling_self_attention.py:68, :72, :332), so it is the oracle for this row.

LoRA (peft is absent from the image, parity unpinned there): the fp32 reference wraps q / v in a plain-torch LoRA
(y = W x + (alpha / r) B A x) holding the same matrices as the product's LoRALinear.

Tolerances: bf16 activations through 4-6 layers -- logits 2e-2 rel-L2, loss 2e-2 abs, gradients 6e-2 rel-L2."""
import copy
import types

import pytest
import torch
import torch.nn as nn

from util import BF16, Report

pytestmark = pytest.mark.gpu


class RefLoRA(nn.Module):
    def __init__(self, base, a, b, scaling):
        super().__init__()
        self.base, self.scaling = base, scaling
        self.a, self.b = nn.Parameter(a.clone()), nn.Parameter(b.clone())

    def forward(self, x):
        return self.base(x) + self.scaling * ((x @ self.a.t()) @ self.b.t())


def _lorafy_pair(product, reference, r, alpha, gen):
    """apply_lora on the product model (random non-zero B so the adapters are live) and the same matrices as RefLoRA
    modules on the fp32 reference model.  Returns [(product LoRALinear, RefLoRA)]."""
    from mmgl_b200.self_attention import LoRALinear, apply_lora
    for p in product.parameters():
        p.requires_grad = False
    for p in reference.parameters():
        p.requires_grad = False
    apply_lora(product, r, alpha, 0.0)
    pairs = []
    ref_mods = dict(reference.named_modules())
    for name, m in list(product.named_modules()):
        if isinstance(m, LoRALinear):
            with torch.no_grad():
                m.lora_B["default"].weight.copy_(torch.randn(m.lora_B["default"].weight.shape, generator=gen) * 0.05)
            parent_name, _, child = name.rpartition(".")
            ref_parent = ref_mods[parent_name]
            ref = RefLoRA(getattr(ref_parent, child), m.lora_A["default"].weight.detach().float(),
                          m.lora_B["default"].weight.detach().float(), m.scaling)
            setattr(ref_parent, child, ref)
            pairs.append((m, ref))
    return pairs


def _finish(rep, out, ref_out, pairs, emb, ref_emb, head, ref_head):
    rep.scalar("loss", out.loss, ref_out.loss, 0.0, 2e-2)
    rep.close("logits", out.logits, ref_out.logits, 2e-2)
    rep.close("d inputs_embeds", emb.grad, ref_emb.grad, 6e-2)
    rep.close("d lm_head", head.grad, ref_head.grad, 6e-2)
    ga = torch.cat([m.lora_A["default"].weight.grad.flatten().float() for m, _ in pairs])
    gb = torch.cat([m.lora_B["default"].weight.grad.flatten().float() for m, _ in pairs])
    ra = torch.cat([r.a.grad.flatten() for _, r in pairs])
    rb = torch.cat([r.b.grad.flatten() for _, r in pairs])
    rep.close("d lora_A (all adapters)", ga, ra, 6e-2)
    rep.close("d lora_B (all adapters)", gb, rb, 6e-2)
    rep.finish()


def test_t5_lora_matches_hf():
    from transformers import T5Config, T5ForConditionalGeneration
    from mmgl_b200 import lm as L
    torch.manual_seed(0)
    gen = torch.Generator().manual_seed(1)
    cfg = T5Config(vocab_size=384, d_model=128, d_kv=64, d_ff=256, num_layers=2, num_decoder_layers=2, num_heads=2,
                   decoder_start_token_id=0, dropout_rate=0.0)
    base = T5ForConditionalGeneration(cfg)
    with torch.no_grad():
        for n, p in base.named_parameters():
            if "relative_attention_bias" in n:
                p.normal_(0, 1.0)          # live position bias (random init is tiny)
            elif "layer_norm" in n:
                p.uniform_(0.7, 1.3)
    reference = copy.deepcopy(base).float()
    product = base
    pairs = _lorafy_pair(product, reference, 8, 1.0, gen)
    # modules_to_save = ["lm_head"]: a trainable, untied copy on both sides
    for m in (product, reference):
        m.lm_head.weight = nn.Parameter(m.lm_head.weight.detach().clone())
    product.cuda().eval()
    reference.cuda().eval()
    assert L.supports(product)
    b, s_enc, s_dec = 2, 200, 40
    emb = (torch.randn(b, s_enc, cfg.d_model, generator=gen) * 1.0).cuda()
    am = torch.ones(b, s_enc, dtype=torch.long)
    am[0, 150:] = 0
    am[1, 90:120] = 0                      # masked bank slots in the middle (concat path: tokens ++ bank mask)
    labels = torch.randint(1, cfg.vocab_size, (b, s_dec), generator=gen)
    labels[0, 30:] = -100
    am, labels = am.cuda(), labels.cuda()
    x = emb.to(BF16).requires_grad_(True)
    out = L.forward(product, inputs_embeds=x, attention_mask=am, labels=labels)
    out.loss.backward()
    xr = emb.to(BF16).float().requires_grad_(True)
    ref_out = reference(inputs_embeds=xr, attention_mask=am, labels=labels)
    ref_out.loss.backward()
    _finish(Report(), out, ref_out, pairs, x, xr, product.lm_head.weight, reference.lm_head.weight)


def test_t5_full_finetune_matches_hf():
    """peft_type "none" (model/modelling_self_attention.py:79 is skipped): every T5 weight trains, including the
    relative-position tables of encoder and decoder -- their gradient comes from the dQ kernel's diagonal sums."""
    from transformers import T5Config, T5ForConditionalGeneration
    from mmgl_b200 import lm as L
    torch.manual_seed(3)
    gen = torch.Generator().manual_seed(4)
    cfg = T5Config(vocab_size=384, d_model=128, d_kv=64, d_ff=256, num_layers=2, num_decoder_layers=2, num_heads=2,
                   decoder_start_token_id=0, dropout_rate=0.0)
    product = T5ForConditionalGeneration(cfg)
    with torch.no_grad():
        for n, p in product.named_parameters():
            if "relative_attention_bias" in n:
                p.normal_(0, 1.0)
            elif "layer_norm" in n:
                p.uniform_(0.7, 1.3)
    reference = copy.deepcopy(product).float()
    product.cuda().eval()
    reference.cuda().eval()
    assert L.supports(product)
    b, s_enc, s_dec = 2, 300, 140                   # 3 x 3 and 2 x 2 score blocks: diagonals span several blocks
    emb = torch.randn(b, s_enc, cfg.d_model, generator=gen).cuda()
    am = torch.ones(b, s_enc, dtype=torch.long)
    am[0, 210:] = 0
    am[1, 100:140] = 0
    labels = torch.randint(1, cfg.vocab_size, (b, s_dec), generator=gen)
    labels[1, 120:] = -100
    am, labels = am.cuda(), labels.cuda()
    out = L.forward(product, inputs_embeds=emb.to(BF16), attention_mask=am, labels=labels)
    out.loss.backward()
    ref_out = reference(inputs_embeds=emb.to(BF16).float(), attention_mask=am, labels=labels)
    ref_out.loss.backward()
    rep = Report()
    rep.scalar("loss", out.loss, ref_out.loss, 0.0, 2e-2)
    rep.close("logits", out.logits, ref_out.logits, 2e-2)
    ref_params = dict(reference.named_parameters())
    seen = 0
    for n, p in product.named_parameters():
        assert p.grad is not None, n
        rep.close("d " + n, p.grad, ref_params[n].grad, 6e-2)
        seen += "relative_attention_bias" in n
    assert seen == 2
    rep.finish()


def test_opt_lora_matches_hf():
    from transformers import OPTConfig, OPTForCausalLM
    from mmgl_b200 import lm as L
    torch.manual_seed(0)
    gen = torch.Generator().manual_seed(2)
    cfg = OPTConfig(vocab_size=384, hidden_size=128, num_hidden_layers=3, ffn_dim=256, num_attention_heads=2,
                    max_position_embeddings=256, word_embed_proj_dim=128, dropout=0.0)
    base = OPTForCausalLM(cfg)
    reference = copy.deepcopy(base).float()
    product = base
    pairs = _lorafy_pair(product, reference, 8, 1.0, gen)
    for m in (product, reference):
        m.lm_head.weight = nn.Parameter(m.lm_head.weight.detach().clone())
    product.cuda().eval()
    reference.cuda().eval()
    assert L.supports(product)
    b, s = 2, 150
    emb = (torch.randn(b, s, cfg.hidden_size, generator=gen) * 0.5).cuda()
    am = torch.ones(b, s, dtype=torch.long)
    am[0, 100:118] = 0                     # right padding of the section, then the bank
    am[1, 140:] = 0
    labels = torch.randint(1, cfg.vocab_size, (b, s), generator=gen)
    labels[:, 118:] = -100                 # bank positions carry no loss (modelling_self_attention.py:327-330)
    am, labels = am.cuda(), labels.cuda()
    x = emb.to(BF16).requires_grad_(True)
    out = L.forward(product, inputs_embeds=x, attention_mask=am, labels=labels)
    out.loss.backward()
    xr = emb.to(BF16).float().requires_grad_(True)
    ref_out = reference(inputs_embeds=xr, attention_mask=am, labels=labels)
    ref_out.loss.backward()
    _finish(Report(), out, ref_out, pairs, x, xr, product.lm_head.weight, reference.lm_head.weight)


@pytest.mark.parametrize("lm", ["t5", "opt"])
def test_self_attention_model_runs_the_lm_on_the_package_kernels(lm, monkeypatch):
    """End to end through the wrapper: with LoRA the LM's layer stack runs in libmmgl_b200.so (attention kernel launches
    are counted), dropout on, loss finite, and the kernel path agrees with the HF module's own forward (run by the test
    itself: the product has no such fallback) at dropout 0."""
    from transformers import CLIPVisionConfig, OPTConfig, RobertaConfig, T5Config
    from mmgl_b200 import _capi, synth
    from mmgl_b200.self_attention import SelfAttentionModel
    torch.manual_seed(0)
    if lm == "t5":
        lm_cfg = T5Config(vocab_size=512, d_model=128, d_kv=64, d_ff=256, num_layers=2, num_decoder_layers=2, num_heads=2,
                          decoder_start_token_id=0, dropout_rate=0.1)
    else:
        lm_cfg = OPTConfig(vocab_size=512, hidden_size=128, num_hidden_layers=2, ffn_dim=256, num_attention_heads=2,
                           max_position_embeddings=512, word_embed_proj_dim=128, dropout=0.1)
    txt = RobertaConfig(vocab_size=512, hidden_size=128, num_hidden_layers=1, num_attention_heads=2,
                        intermediate_size=256, max_position_embeddings=80, pad_token_id=1)
    vis = CLIPVisionConfig(hidden_size=128, intermediate_size=256, num_hidden_layers=1, num_attention_heads=2,
                           image_size=32, patch_size=16)
    args = types.SimpleNamespace(context="all", decoder_only=lm == "opt", neighbor_mode="embedding", position_type="none",
                                 n_text_tokens=2, n_visual_tokens=2, model_name_or_path=lm_cfg, peft_type="lora",
                                 text_model=txt, visual_model=vis, max_output_length=16, freeze_lm=False,
                                 max_text_neighbors=3, max_image_neighbors=2, lora_r=8, lora_alpha=1, lora_dropout=0.0)
    model = SelfAttentionModel(args, tokenizer=None).cuda()
    spec = synth.BatchSpec(batch=2, max_input_length=48, max_output_length=16, text_neighbors=3, image_neighbors=2,
                           vocab_size=512, neighbor_vocab_size=512, image_size=32, decoder_only=lm == "opt")
    batch = synth.to_device(synth.make_batch(spec, seed=3), torch.device("cuda"))
    model.train()
    n0 = _capi.launch_count()
    out = model(**batch)
    out.loss.backward()
    assert torch.isfinite(out.loss)
    assert _capi.launch_count() - n0 > 60, "the LM did not run on the package's kernels"
    # invariant I7: T5 -> (B, S_out, V); OPT -> (B, S_in + S_out + Nk, V), Nk = (T + I) * n_tokens
    want = (2, 16, 512) if lm == "t5" else (2, 48 + 16 + (3 + 2) * 2, 512)
    assert tuple(out.logits.shape) == want, out.logits.shape
    got = {n for n, p in model.named_parameters() if p.grad is not None and float(p.grad.abs().max()) > 0}
    # at init B = 0, so dA is exactly zero; B, the neighbor projection and the lm_head copy must all train
    assert any("lora_B" in n for n in got) and any("lm_head" in n for n in got), got
    if lm == "t5":
        assert any("text_embeddings" in n for n in got), got
    # (decoder-only: the reference appends the bank AFTER the section tokens, modelling_self_attention.py:323-330, so
    # under the causal mask no loss position can see it and the neighbor projection gets no gradient -- reproduced)
    # kernel path vs the HF forward of the same wrapper, dropout off
    model.eval()
    from mmgl_b200 import self_attention as SA

    def hf_forward(lm_, **kw):   # test-side oracle: the HF module's own forward (the product has no such path)
        if isinstance(lm_, SA._PeftShim):
            return lm_(**kw)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            return lm_(**kw)
    with torch.no_grad():
        a = model(**batch)
        monkeypatch.setattr(SA, "run_language_model", hf_forward)
        bref = model(**batch)
    rep = Report()
    rep.scalar("loss (kernels vs HF forward)", a.loss, bref.loss, 0.0, 3e-2)
    rep.close("logits (kernels vs HF forward)", a.logits, bref.logits, 3e-2)
    rep.finish()


def test_opt_prefix_tuning_matches_hf_past_key_values():
    """Prefix tuning on OPT (peft PrefixTuningConfig, model/modelling_self_attention.py:88-92; peft absent -> semantics
    restated): the per-layer virtual-token K / V are what HF's own forward consumes when handed as past_key_values, so
    HF fp32 with a DynamicCache built from the same prefix table is the oracle -- loss, logits, d prefix, d inputs."""
    from transformers import DynamicCache, OPTConfig, OPTForCausalLM
    from mmgl_b200 import lm as L
    torch.manual_seed(0)
    gen = torch.Generator().manual_seed(4)
    cfg = OPTConfig(vocab_size=384, hidden_size=128, num_hidden_layers=3, ffn_dim=256, num_attention_heads=2,
                    max_position_embeddings=256, word_embed_proj_dim=128, dropout=0.0)
    product = OPTForCausalLM(cfg)
    reference = copy.deepcopy(product).float()
    for m in (product, reference):
        for p in m.parameters():
            p.requires_grad = False
    product.cuda().eval()
    reference.cuda().eval()
    n_pre, b, s, heads, d = 20, 2, 150, 2, 64
    table = (torch.randn(n_pre, cfg.num_hidden_layers, 2, cfg.hidden_size, generator=gen) * 0.5).to(BF16).float().cuda()
    emb = (torch.randn(b, s, cfg.hidden_size, generator=gen) * 0.5).cuda()
    am = torch.ones(b, s, dtype=torch.long)
    am[0, 100:118] = 0
    am[1, 140:] = 0
    labels = torch.randint(1, cfg.vocab_size, (b, s), generator=gen)
    labels[:, 118:] = -100
    am, labels = am.cuda(), labels.cuda()

    w = table.clone().requires_grad_(True)
    x = emb.to(BF16).requires_grad_(True)
    out = L.opt_forward(product, inputs_embeds=x, attention_mask=am, labels=labels, prefix_kv=w)
    out.loss.backward()

    wr = table.clone().requires_grad_(True)
    xr = emb.to(BF16).float().requires_grad_(True)
    cache = DynamicCache(config=cfg)
    for l in range(cfg.num_hidden_layers):
        k = wr[:, l, 0].view(n_pre, heads, d).permute(1, 0, 2)[None].expand(b, -1, -1, -1)
        v = wr[:, l, 1].view(n_pre, heads, d).permute(1, 0, 2)[None].expand(b, -1, -1, -1)
        cache.update(k, v, l)
    full = torch.cat((torch.ones(b, n_pre, dtype=am.dtype, device=am.device), am), dim=1)
    ref_out = reference(inputs_embeds=xr, attention_mask=full, past_key_values=cache, labels=labels)
    ref_out.loss.backward()
    rep = Report()
    rep.scalar("loss", out.loss, ref_out.loss, 0.0, 2e-2)
    rep.close("logits", out.logits, ref_out.logits, 2e-2)
    rep.close("d prefix table", w.grad, wr.grad, 6e-2)
    rep.close("d inputs_embeds", x.grad, xr.grad, 6e-2)
    rep.finish()


def test_self_attention_model_prefix_tuning_trains():
    from transformers import CLIPVisionConfig, OPTConfig, RobertaConfig
    from mmgl_b200 import synth
    from mmgl_b200.self_attention import SelfAttentionModel
    torch.manual_seed(0)
    lm_cfg = OPTConfig(vocab_size=512, hidden_size=128, num_hidden_layers=2, ffn_dim=256, num_attention_heads=2,
                       max_position_embeddings=512, word_embed_proj_dim=128, dropout=0.1)
    txt = RobertaConfig(vocab_size=512, hidden_size=128, num_hidden_layers=1, num_attention_heads=2,
                        intermediate_size=256, max_position_embeddings=80, pad_token_id=1)
    vis = CLIPVisionConfig(hidden_size=128, intermediate_size=256, num_hidden_layers=1, num_attention_heads=2,
                           image_size=32, patch_size=16)
    args = types.SimpleNamespace(context="all", decoder_only=True, neighbor_mode="embedding", position_type="none",
                                 n_text_tokens=2, n_visual_tokens=2, model_name_or_path=lm_cfg, peft_type="prefix",
                                 text_model=txt, visual_model=vis, max_output_length=16, freeze_lm=False,
                                 max_text_neighbors=3, max_image_neighbors=2, lora_r=8, lora_alpha=1, lora_dropout=0.0)
    model = SelfAttentionModel(args, tokenizer=None).cuda().train()
    keys = list(model.state_dict())
    assert "lm.prompt_encoder.default.embedding.weight" in keys
    assert tuple(model.lm.prompt_encoder["default"]["embedding"].weight.shape) == (20, 2 * 2 * 128)
    spec = synth.BatchSpec(batch=2, max_input_length=48, max_output_length=16, text_neighbors=3, image_neighbors=2,
                           vocab_size=512, neighbor_vocab_size=512, image_size=32, decoder_only=True)
    batch = synth.to_device(synth.make_batch(spec, seed=3), torch.device("cuda"))
    out = model(**batch)
    out.loss.backward()
    assert torch.isfinite(out.loss)
    g = model.lm.prompt_encoder["default"]["embedding"].weight.grad
    assert g is not None and float(g.abs().max()) > 0
    trainable = {n for n, p in model.named_parameters() if p.requires_grad}
    assert all(n.startswith(("lm.prompt_encoder", "text_", "visual_")) for n in trainable), trainable


def test_t5_prefix_tuning_matches_hf_past_key_values():
    """Prefix tuning on T5 (peft PrefixTuningConfig for SEQ_2_SEQ_LM, model/modelling_self_attention.py:88-92; peft absent
    -> semantics restated: every DECODER layer's self-attention sees 20 virtual K / V in front of its own keys; the
    cross-attention and the encoder are untouched -- the effective behaviour of peft with the reference-era transformers,
    see lm._t5_stack).  Oracle: HF T5 fp32 handed the same tensors as the self-attention half of an EncoderDecoderCache
    (empty cross-attention half) -- loss, logits, d prefix, d encoder inputs, with encoder padding."""
    from transformers import DynamicCache, EncoderDecoderCache, T5Config, T5ForConditionalGeneration
    from mmgl_b200 import lm as L
    torch.manual_seed(0)
    gen = torch.Generator().manual_seed(6)
    cfg = T5Config(vocab_size=384, d_model=128, d_kv=64, d_ff=256, num_layers=2, num_decoder_layers=3, num_heads=2,
                   decoder_start_token_id=0, dropout_rate=0.0)
    reference = T5ForConditionalGeneration(cfg)
    product = T5ForConditionalGeneration(cfg)
    product.load_state_dict(reference.state_dict())
    for m in (product, reference):
        for p in m.parameters():
            p.data = p.data.to(BF16).float()
            p.requires_grad = False
    product.cuda().eval()
    reference.cuda().eval()
    n_pre, b, s, sd, heads, d = 20, 2, 150, 40, 2, 64
    table = (torch.randn(n_pre, cfg.num_decoder_layers, 2, heads * d, generator=gen) * 0.5).to(BF16).float().cuda()
    emb = (torch.randn(b, s, cfg.d_model, generator=gen) * 0.5).cuda()
    labels = torch.randint(1, cfg.vocab_size, (b, sd), generator=gen)
    labels[1, 30:] = -100
    labels = labels.cuda()
    am = torch.ones(b, s, dtype=torch.long)
    am[1, 120:] = 0
    am = am.cuda()

    w = table.clone().requires_grad_(True)
    x = emb.to(BF16).requires_grad_(True)
    out = L.t5_forward(product, inputs_embeds=x, attention_mask=am, labels=labels, prefix_kv=w)
    out.loss.backward()

    wr = table.clone().requires_grad_(True)
    xr = emb.to(BF16).float().requires_grad_(True)
    sc, cc = DynamicCache(), DynamicCache()   # (config=cfg would size both for num_layers, not num_decoder_layers)
    for l in range(cfg.num_decoder_layers):
        k = wr[:, l, 0].view(n_pre, heads, d).permute(1, 0, 2)[None].expand(b, -1, -1, -1)
        v = wr[:, l, 1].view(n_pre, heads, d).permute(1, 0, 2)[None].expand(b, -1, -1, -1)
        sc.update(k, v, l)
    ref_out = reference(inputs_embeds=xr, attention_mask=am, labels=labels, past_key_values=EncoderDecoderCache(sc, cc),
                        decoder_attention_mask=torch.ones(b, n_pre + sd, dtype=torch.long, device="cuda"))
    ref_out.loss.backward()
    plain = L.t5_forward(product, inputs_embeds=x.detach(), attention_mask=am, labels=labels)
    assert abs(float(plain.loss) - float(out.loss)) > 1e-3, "the prefix must change the loss (path is live)"
    rep = Report()
    rep.scalar("loss", out.loss, ref_out.loss, 0.0, 2e-2)
    rep.close("logits", out.logits, ref_out.logits, 2e-2)
    rep.close("d prefix table", w.grad, wr.grad, 6e-2)
    rep.close("d inputs_embeds", x.grad, xr.grad, 6e-2)
    rep.finish()


def test_self_attention_model_t5_prefix_tuning_trains():
    """the wrapper with peft_type == 'prefix' on T5: peft's table shape (2 x 20 rows: num_transformer_submodules = 2, of
    which get_prompt reads the first 20), only the table and the neighbor projections train"""
    from transformers import CLIPVisionConfig, RobertaConfig, T5Config
    from mmgl_b200 import synth
    from mmgl_b200.self_attention import SelfAttentionModel
    torch.manual_seed(0)
    lm_cfg = T5Config(vocab_size=512, d_model=128, d_kv=64, d_ff=256, num_layers=2, num_decoder_layers=2, num_heads=2,
                      decoder_start_token_id=0, dropout_rate=0.1)
    txt = RobertaConfig(vocab_size=512, hidden_size=128, num_hidden_layers=1, num_attention_heads=2,
                        intermediate_size=256, max_position_embeddings=80, pad_token_id=1)
    vis = CLIPVisionConfig(hidden_size=128, intermediate_size=256, num_hidden_layers=1, num_attention_heads=2,
                           image_size=32, patch_size=16)
    args = types.SimpleNamespace(context="all", decoder_only=False, neighbor_mode="embedding", position_type="none",
                                 n_text_tokens=2, n_visual_tokens=2, model_name_or_path=lm_cfg, peft_type="prefix",
                                 text_model=txt, visual_model=vis, max_output_length=16, freeze_lm=False,
                                 max_text_neighbors=3, max_image_neighbors=2, lora_r=8, lora_alpha=1, lora_dropout=0.0)
    model = SelfAttentionModel(args, tokenizer=None).cuda().train()
    assert tuple(model.lm.prompt_encoder["default"]["embedding"].weight.shape) == (40, 2 * 2 * 128)
    spec = synth.BatchSpec(batch=2, max_input_length=48, max_output_length=16, text_neighbors=3, image_neighbors=2,
                           vocab_size=512, neighbor_vocab_size=512, image_size=32, decoder_only=False, pad_token_id=0)
    batch = synth.to_device(synth.make_batch(spec, seed=3), torch.device("cuda"))
    out = model(**batch)
    out.loss.backward()
    assert torch.isfinite(out.loss)
    g = model.lm.prompt_encoder["default"]["embedding"].weight.grad
    assert g is not None and float(g[:20].abs().max()) > 0 and float(g[20:].abs().max()) == 0.0
    trainable = {n for n, p in model.named_parameters() if p.requires_grad}
    assert all(n.startswith(("lm.prompt_encoder", "text_", "visual_")) for n in trainable), trainable
