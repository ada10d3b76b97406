"""Pin the CPU oracle (oracle/mmgl_oracle.py) against golden vectors produced by the
REAL reference modules (tests/golden/make_golden.py).  CPU only."""
import pytest
import torch

from oracle import mmgl_oracle as O

TOL = dict(rtol=1e-5, atol=2e-6)  # fp32 vs fp32, different op order only


@pytest.mark.parametrize("name", ["xattn_layer_preln", "xattn_layer_postln", "xattn_layer_d64_preln",
                                  "xattn_layer_d64_postln"])
def test_cross_layer_forward_backward(golden, name):
    g = golden(name)
    p = {k: v.clone().requires_grad_(True) for k, v in g["state"].items()}
    x = g["x"].clone().requires_grad_(True)
    bank = g["bank"].clone().requires_grad_(True)
    add = O.expand_mask(g["mask"], x.dtype, x.shape[1])
    y = O.mpt_decoder_layer(x, p, g["cfg"]["num_heads"], cross_attention=True, bank=bank, bank_add_mask=add,
                            do_layer_norm_before=g["cfg"]["do_layer_norm_before"])
    torch.testing.assert_close(y, g["y"], **TOL)
    (y * g["w"]).sum().backward()
    torch.testing.assert_close(x.grad, g["dx"], rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(bank.grad, g["dbank"], rtol=1e-4, atol=1e-5)
    for k, gr in g["grads"].items():
        torch.testing.assert_close(p[k].grad, gr, rtol=1e-4, atol=1e-5, msg=lambda m, k=k: f"{k}: {m}")


def test_xattn_core_matches_attention(golden):
    """xattn_core (the kernel's slice) reproduces the attention inside the layer."""
    g = golden("xattn_layer_preln")
    p, nh = g["state"], g["cfg"]["num_heads"]
    x = torch.nn.functional.layer_norm(g["x"], (64,), p["self_attn_layer_norm.weight"], p["self_attn_layer_norm.bias"])
    ref = O.mpt_attention(x, g["bank"], O.expand_mask(g["mask"], x.dtype, x.shape[1]), p, nh)
    q = torch.nn.functional.linear(x, p["self_attn.q_proj.weight"], p["self_attn.q_proj.bias"]) * (16 ** -0.5)
    k = torch.nn.functional.linear(g["bank"], p["self_attn.k_proj.weight"], p["self_attn.k_proj.bias"])
    v = torch.nn.functional.linear(g["bank"], p["self_attn.v_proj.weight"], p["self_attn.v_proj.bias"])
    o, lse = O.xattn_core(q, k, v, g["mask"], nh)
    out = torch.nn.functional.linear(o, p["self_attn.out_proj.weight"], p["self_attn.out_proj.bias"])
    torch.testing.assert_close(out, ref, **TOL)
    assert lse.shape == (2, nh, 24)


def test_mpt_causal_lm(golden):
    g = golden("mpt_lm")
    loss, logits = O.mpt_causal_lm(g["state"], g["cfg"], g["input_ids"], g["attention_mask"], g["input_ids"],
                                   g["bank"], g["bank_mask"])
    torch.testing.assert_close(logits, g["logits"], rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(loss, g["loss"], rtol=1e-5, atol=1e-6)
    # invariant I1: gates == 0 -> bank has no influence
    st = {k: (torch.zeros_like(v) if "gating" in k else v) for k, v in g["state"].items()}
    loss0, logits0 = O.mpt_causal_lm(st, g["cfg"], g["input_ids"], g["attention_mask"], g["input_ids"],
                                     g["bank"], g["bank_mask"])
    torch.testing.assert_close(logits0, g["logits_gates0"], rtol=1e-4, atol=1e-5)
    _, logits_nb = O.mpt_causal_lm(st, g["cfg"], g["input_ids"], g["attention_mask"], None, None, None)
    torch.testing.assert_close(logits_nb, logits0, rtol=0, atol=0)
    # I3: with live gates the output differs
    assert (logits - logits0).abs().max() > 1e-3


def test_masked_neighbors_have_zero_influence(golden):
    """Invariant I2: bank rows whose mask is False never change the logits."""
    g = golden("mpt_lm")
    bank2 = g["bank"].clone()
    bank2[~g["bank_mask"]] = 123.0
    _, a = O.mpt_causal_lm(g["state"], g["cfg"], g["input_ids"], g["attention_mask"], None, g["bank"], g["bank_mask"])
    _, b = O.mpt_causal_lm(g["state"], g["cfg"], g["input_ids"], g["attention_mask"], None, bank2, g["bank_mask"])
    assert torch.equal(a, b)


def test_cross_wrapper(golden):
    g = golden("wrapper_cross")
    b = g["batch"]
    p = g["state"]
    n_tok = g["cfg"]["n_tokens"]
    te = O.neighbor_projection(g["text_pooled"], p["text_embeddings.weight"], p["text_embeddings.bias"],
                               p["text_position_embeddings.weight"], b["neighbor_pos_ids"], n_tok)
    ve = O.neighbor_projection(g["visual_pooled"], p["visual_embeddings.weight"], p["visual_embeddings.bias"],
                               p["visual_position_embeddings.weight"], b["neighbor_images_pos_ids"], n_tok)
    bank, mask = O.pack_bank(te, b["neighbor_pos_ids"], b["text_locations"], ve, b["neighbor_images_pos_ids"],
                             b["image_locations"])
    torch.testing.assert_close(bank, g["bank"], **TOL)
    assert torch.equal(mask, g["bank_mask"])
    loss, logits = O.cross_attention_model_from_pooled(p, g["cfg"], b, g["text_pooled"], g["visual_pooled"])
    torch.testing.assert_close(logits, g["logits"], rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(loss, g["loss"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("lm", ["t5", "opt"])
@pytest.mark.parametrize("pt", ["none", "laplacian", "gnn"])
def test_self_wrapper_concat(golden, lm, pt):
    g = golden(f"wrapper_self_{lm}_{pt}")
    b, p, n_tok = g["batch"], g["state"], g["cfg"]["n_tokens"]
    has_pos = pt != "none"
    te = O.neighbor_projection(g["text_pooled"], p["text_embeddings.weight"], p["text_embeddings.bias"],
                               p.get("text_position_embeddings.weight") if has_pos else None, b["neighbor_pos_ids"], n_tok)
    ve = O.neighbor_projection(g["visual_pooled"], p["visual_embeddings.weight"], p["visual_embeddings.bias"],
                               p.get("visual_position_embeddings.weight") if has_pos else None,
                               b["neighbor_images_pos_ids"], n_tok)
    bank, mask = O.pack_bank(te, b["neighbor_pos_ids"], b["text_locations"], ve, b["neighbor_images_pos_ids"],
                             b["image_locations"])
    if pt == "laplacian":
        bank = O.lpe_add(bank, b["lpe"], p["lpe_embeddings.weight"], p["lpe_embeddings.bias"], n_tok)
    if pt == "gnn":
        bank = O.gnn_add(bank, b["graph"], p["gnn.w1.weight"], p["gnn.w2.weight"], n_tok)
    emb = torch.nn.functional.embedding(b["input_ids"], p["input_embeddings.weight"])
    embs, am, labels = O.concat_neighbors(emb, b["attention_mask"], b["labels"], bank, mask, g["cfg"]["decoder_only"])
    torch.testing.assert_close(embs, g["inputs_embeds"], rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(am.float(), g["attention_mask"].float(), rtol=0, atol=0)
    assert torch.equal(labels, g["labels"])


def test_gcn(golden):
    g = golden("gcn")
    x = g["x"].clone().requires_grad_(True)
    w1 = g["state"]["w1.weight"].clone().requires_grad_(True)
    w2 = g["state"]["w2.weight"].clone().requires_grad_(True)
    y = O.gcn_forward(x, g["adj"], w1, w2)
    torch.testing.assert_close(y, g["y"], **TOL)
    (y * g["w"]).sum().backward()
    torch.testing.assert_close(x.grad, g["dx"], rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(w1.grad, g["grads"]["w1.weight"], rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(w2.grad, g["grads"]["w2.weight"], rtol=1e-4, atol=1e-5)
    # invariant I6: adj == 0 -> agg == 0
    y0 = O.gcn_forward(g["x"], torch.zeros_like(g["adj"]), g["state"]["w1.weight"], g["state"]["w2.weight"])
    xr = torch.cat((torch.zeros(3, 1, 48), g["x"]), 1)
    h = torch.relu(torch.nn.functional.linear(torch.cat((xr, torch.zeros_like(xr)), -1), g["state"]["w1.weight"]))
    exp = torch.nn.functional.linear(torch.cat((h, torch.zeros_like(h)), -1), g["state"]["w2.weight"])[:, 1:]
    torch.testing.assert_close(y0, exp, **TOL)


def test_lora_identity_at_init():
    """Invariant I5: LoRA with B == 0 is the base linear (parity otherwise unpinned: peft absent)."""
    gen = torch.Generator().manual_seed(0)
    x = torch.randn(5, 32, generator=gen)
    w = torch.randn(16, 32, generator=gen)
    a, b = O.lora_init(8, 32, 16, gen)
    assert torch.equal(O.lora_linear(x, w, None, a, b, 1.0, 8), torch.nn.functional.linear(x, w))
    b = torch.randn(16, 8, generator=gen)
    y = O.lora_linear(x, w, None, a, b, 2.0, 8)
    torch.testing.assert_close(y, x @ (w + (2.0 / 8) * b @ a).T, rtol=1e-5, atol=1e-5)


def test_t5_relative_position_bias_matches_hf():
    """The compact per-distance bias vector the attention kernel takes ([heads, Sq + Sk - 1]) and the oracle's dense
    restatement both equal HF T5Attention.compute_bias (encoder: bidirectional buckets; decoder: causal buckets)."""
    import types
    from transformers import T5Config
    from transformers.models.t5.modeling_t5 import T5Attention
    from mmgl_b200 import lm as L
    torch.manual_seed(0)
    for is_decoder in (False, True):
        cfg = T5Config(d_model=64, d_kv=16, num_heads=4, is_decoder=is_decoder, relative_attention_num_buckets=32,
                       relative_attention_max_distance=128)
        attn = T5Attention(cfg, has_relative_attention_bias=True, layer_idx=0)
        with torch.no_grad():
            attn.relative_attention_bias.weight.normal_()
        for sq, sk in ((7, 7), (130, 130), (40, 300)):
            dense = attn.compute_bias(sq, sk)                                   # [1, nh, sq, sk]
            vec = L.t5_rel_bias(attn, sq, sk)                                   # [nh, sq + sk - 1]
            idx = torch.arange(sk)[None, :] - torch.arange(sq)[:, None] + sq - 1
            assert torch.equal(vec[:, idx][None], dense.detach()), (is_decoder, sq, sk)
            ref = O.t5_position_bias(attn.relative_attention_bias.weight.detach(), sq, sk, not is_decoder)
            assert torch.equal(ref, dense.detach())


def _additive_mask(key_mask, sq, causal):
    """[B,1,Sq,Sk] additive fp32 mask the way HF builds it: finfo.min where a key is padding or (causal) in the future."""
    b, sk = key_mask.shape
    allowed = key_mask[:, None, None, :].bool().expand(b, 1, sq, sk)
    if causal:
        allowed = allowed & torch.tril(torch.ones(sq, sk, dtype=torch.bool))[None, None]
    return torch.zeros(b, 1, sq, sk).masked_fill(~allowed, torch.finfo(torch.float32).min)


@pytest.mark.parametrize("decoder", [False, True])
def test_attention_core_matches_hf_t5_attention(decoder):
    """oracle.attention_core (the checker of the self-attention kernels) against the HF module it restates, run here in
    fp32: T5Attention.forward (HF models/t5/modeling_t5.py:253-345) -- unscaled scores + bucketed position bias + mask,
    encoder (bidirectional) and decoder (causal) self-attention with key padding, forward and input gradient."""
    from transformers import T5Config
    from transformers.models.t5.modeling_t5 import T5Attention
    torch.manual_seed(1 + decoder)
    cfg = T5Config(d_model=64, d_kv=16, num_heads=4, is_decoder=decoder, dropout_rate=0.0)
    attn = T5Attention(cfg, has_relative_attention_bias=True, layer_idx=0).eval()
    with torch.no_grad():
        attn.relative_attention_bias.weight.normal_()
    b, s = 2, 50
    x = torch.randn(b, s, 64, requires_grad=True)
    key_mask = torch.ones(b, s, dtype=torch.long)
    key_mask[0, 41:] = 0
    key_mask[1, 20:27] = 0
    d_o = torch.randn(b, s, 64)
    ref = attn(x, mask=_additive_mask(key_mask, s, decoder))[0]
    ref.backward(d_o)
    want_dx = x.grad.clone()
    x.grad = None
    bias = O.t5_position_bias(attn.relative_attention_bias.weight.detach(), s, s, not decoder)
    q, k, v = (lin(x) for lin in (attn.q, attn.k, attn.v))
    got = attn.o(O.attention_core(q, k, v, 4, 1.0, key_mask.bool(), decoder, bias))
    got.backward(d_o)
    torch.testing.assert_close(got, ref, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(x.grad, want_dx, rtol=1e-4, atol=1e-5)


def test_attention_core_matches_hf_opt_attention():
    """... and OPTAttention.forward (HF models/opt/modeling_opt.py:74-182; the self branch of MPTAttention,
    model/modelling_cross_attention.py:201-275, is its copy): q scaled by d^-1/2, causal + key-padding additive mask."""
    from transformers import OPTConfig
    from transformers.models.opt.modeling_opt import OPTAttention
    torch.manual_seed(3)
    cfg = OPTConfig(hidden_size=64, num_attention_heads=4, ffn_dim=128, num_hidden_layers=1, dropout=0.0,
                    attention_dropout=0.0)
    cfg._attn_implementation = "eager"
    attn = OPTAttention(cfg, layer_idx=0).eval()
    b, s = 2, 45
    x = torch.randn(b, s, 64, requires_grad=True)
    key_mask = torch.ones(b, s, dtype=torch.long)
    key_mask[0, :6] = 0                                  # left padding: the first rows attend nothing but padding
    key_mask[1, 30:] = 0
    d_o = torch.randn(b, s, 64)
    ref = attn(x, attention_mask=_additive_mask(key_mask, s, True))[0]
    # rows whose every allowed key is padding are uniform over ALL keys in both (finfo.min ties); they carry no loss in
    # the model, compare the rest
    rows = torch.ones(b, s, dtype=torch.bool)
    rows[0, :6] = False
    (ref * rows[..., None]).backward(d_o)
    want_dx = x.grad.clone()
    x.grad = None
    q, k, v = attn.q_proj(x), attn.k_proj(x), attn.v_proj(x)
    got = attn.out_proj(O.attention_core(q, k, v, 4, 16 ** -0.5, key_mask.bool(), True))
    (got * rows[..., None]).backward(d_o)
    torch.testing.assert_close(got[rows], ref[rows], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(x.grad, want_dx, rtol=1e-4, atol=1e-5)


def test_compiled_reference_in_oracle_ref_is_the_reference(golden):
    """oracle/_ref (the reference's model files byte-compiled by oracle/build_ref.py; bench.py's `--impl reference` /
    `--impl eager` arms run it) reproduces a golden vector made from /root/reference BIT-exactly: it IS the reference."""
    from oracle import ref_loader as R
    if not R.available():
        pytest.skip("oracle/_ref not built (python oracle/build_ref.py needs /root/reference)")
    import types
    from transformers import OPTConfig
    xa = R.cross_attention_module()
    g = golden("xattn_layer_preln")
    cfg = OPTConfig(vocab_size=512, hidden_size=64, num_hidden_layers=4, ffn_dim=128, num_attention_heads=4,
                    max_position_embeddings=200, word_embed_proj_dim=64, do_layer_norm_before=True)
    args = types.SimpleNamespace(neighbor_layer_wise=2, neighbor_mode="cross_attention", peft_type="flamingo", lora_r=64,
                                 lora_alpha=1, lora_dropout=0.0)
    layer = xa.MPTDecoderLayer(xa.MPTConfig(args, cfg), cross_attention=True)
    layer.load_state_dict(g["state"])
    layer.eval()
    x = g["x"].clone().requires_grad_(True)
    y = layer(x, neighbor_embeds=g["bank"], neighbor_attention_mask=xa._expand_mask(g["mask"], x.dtype, tgt_len=x.shape[1]))[0]
    assert torch.equal(y, g["y"])
    (y * g["w"]).sum().backward()
    assert torch.equal(x.grad, g["dx"])
