"""SelfAttentionModel (concatenated-embedding path) on the GPU against fixtures captured from the REAL reference
wrapper (tests/golden/wrapper_self_{t5,opt}_{none,laplacian,gnn}.pt): the tensors handed to the HF language model --
inputs_embeds (token embeddings ++ packed neighbor bank with the position / Laplacian-PE / GCN terms), the
concatenated attention mask and the padded labels.  The LM itself is HF library code on both sides.

Tolerances: inputs_embeds 6e-3 rel-L2 (bf16 bank vs the fp32 reference; the gnn case chains two bf16 GEMMs: 1.2e-2);
attention mask and labels bit-exact."""
import types

import pytest
import torch

from util import BF16, Report

pytestmark = pytest.mark.gpu


def _args(lm, pt, dec_only):
    from transformers import CLIPVisionConfig, OPTConfig, RobertaConfig, T5Config
    if lm == "t5":
        lm_cfg = T5Config(vocab_size=512, d_model=64, d_kv=16, d_ff=128, num_layers=2, num_decoder_layers=2,
                          num_heads=4, decoder_start_token_id=0)
    else:
        lm_cfg = OPTConfig(vocab_size=512, hidden_size=64, num_hidden_layers=4, ffn_dim=128, num_attention_heads=4,
                           max_position_embeddings=200, word_embed_proj_dim=64, dropout=0.0)
    txt = RobertaConfig(vocab_size=512, hidden_size=32, num_hidden_layers=2, num_attention_heads=2,
                        intermediate_size=64, max_position_embeddings=40, pad_token_id=1)
    vis = CLIPVisionConfig(hidden_size=32, intermediate_size=64, num_hidden_layers=2, num_attention_heads=2,
                           image_size=32, patch_size=16)
    return types.SimpleNamespace(context="all", decoder_only=dec_only, neighbor_mode="embedding", position_type=pt,
                                 n_text_tokens=2, n_visual_tokens=2, model_name_or_path=lm_cfg, peft_type="none",
                                 text_model=txt, visual_model=vis, max_output_length=16, freeze_lm=False,
                                 max_text_neighbors=3, max_image_neighbors=2, lora_r=8, lora_alpha=1, lora_dropout=0.0)


@pytest.mark.parametrize("lm", ["t5", "opt"])
@pytest.mark.parametrize("pt", ["none", "laplacian", "gnn"])
def test_concat_path_inputs_match_reference(golden, lm, pt):
    from mmgl_b200.self_attention import SelfAttentionModel
    g = golden(f"wrapper_self_{lm}_{pt}")
    model = SelfAttentionModel(_args(lm, pt, g["cfg"]["decoder_only"]), tokenizer=None)
    state = dict(g["state"])
    emb_w = state.pop("input_embeddings.weight")
    missing, unexpected = model.load_state_dict(state, strict=False)
    assert not unexpected, unexpected
    assert all(k.startswith(("lm.", "text_model.", "visual_model.", "input_embeddings.")) for k in missing), missing
    model.input_embeddings.weight.data.copy_(emb_w)
    model.cuda().eval()
    model.skip_padding_neighbors = False   # the fixture's pooled features cover every neighbor slot
    tp, vp = g["text_pooled"].cuda(), g["visual_pooled"].cuda()
    model.encode_text = lambda ids, am: tp.reshape(-1, tp.shape[-1]).to(BF16)
    model.encode_images = lambda px: vp.reshape(-1, vp.shape[-1]).to(BF16)
    cap = {}
    model._run_lm = lambda **kw: cap.update(kw) or types.SimpleNamespace(loss=None, logits=None)
    batch = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in g["batch"].items()}
    model(**batch)
    rep = Report()
    rep.close("inputs_embeds", cap["inputs_embeds"], g["inputs_embeds"], 1.2e-2 if pt == "gnn" else 6e-3)
    rep.finish()
    assert torch.equal(cap["attention_mask"].float().cpu(), g["attention_mask"].float()), "attention mask differs"
    assert torch.equal(cap["labels"].cpu(), g["labels"]), "labels differ"


@pytest.mark.parametrize("lm", ["t5", "opt"])
def test_lora_model_trains_and_is_identity_at_init(lm):
    """Invariant I5: with B = 0 (init) the LoRA-adapted LM equals the base LM; after one backward the LoRA
    matrices (and only adapter / projection parameters) have gradients."""
    from transformers import OPTConfig, T5Config
    from mmgl_b200.self_attention import LoRALinear, SelfAttentionModel
    torch.manual_seed(0)
    a = _args(lm, "none", lm == "opt")
    a.peft_type = "lora"
    # head_dim 64: the LM runs on the package's kernels (the fixtures' head_dim-16 models are rejected, no HF fallback)
    a.model_name_or_path = (T5Config(vocab_size=512, d_model=128, d_kv=64, d_ff=256, num_layers=2, num_decoder_layers=2,
                                     num_heads=2, decoder_start_token_id=0) if lm == "t5" else
                            OPTConfig(vocab_size=512, hidden_size=128, num_hidden_layers=4, ffn_dim=256, num_attention_heads=2,
                                      max_position_embeddings=200, word_embed_proj_dim=128, dropout=0.0))
    a.context = "text_only"      # no frozen encoders needed: the test isolates the adapters
    model = SelfAttentionModel(a, tokenizer=None).cuda()
    n_lora = sum(isinstance(m, LoRALinear) for m in model.modules())
    assert n_lora == (2 * (2 + 2 * 2) if lm == "t5" else 2 * 4), n_lora   # q and v of every attention block
    gen = torch.Generator().manual_seed(1)
    b, s = 2, 12
    ids = torch.randint(4, 512, (b, s), generator=gen).cuda()
    am = torch.ones(b, s, dtype=torch.long).cuda()
    labels = torch.randint(2, 512, (b, 8 if lm == "t5" else s), generator=gen).cuda()
    model.eval()
    model.neighbor_mode, model.context = "raw", "text_only"      # LM only: isolates the adapters
    out = model(input_ids=ids, attention_mask=am, labels=labels)
    base = model.lm.base_model.model
    for m in base.modules():
        if isinstance(m, LoRALinear):
            assert float(m.lora_B["default"].weight.abs().max()) == 0.0
    model.train()
    out = model(input_ids=ids, attention_mask=am, labels=labels)
    out.loss.backward()
    got = {n for n, p in model.named_parameters() if p.grad is not None and float(p.grad.abs().max()) > 0}
    assert any("lora_B" in n for n in got), "LoRA B received no gradient"
    assert all(("lora_" in n) or ("lm_head" in n) or n.startswith(("text_", "visual_")) for n in got), got
    keys = list(model.state_dict())
    assert any(".lora_A.default.weight" in k and k.startswith("lm.base_model.model.") for k in keys)
