"""Generate the golden vectors under tests/golden/ from the REAL reference.

Run in the build container only (needs /root/reference; the GPU box has none):

    python tests/golden/make_golden.py

It imports the reference modules by file path (recipe: SURVEY.md Appendix A),
drives them with seeded tiny inputs and random-init weights, and stores inputs,
state dicts and outputs (and gradients) as small ``.pt`` files.  The oracle
(oracle/mmgl_oracle.py) is pinned against these in tests/test_oracle_golden.py;
the CUDA kernels are compared with them in the ``-m gpu`` tests.
"""
import importlib.util
import os
import sys
import tempfile
import types

import torch

REF = "/root/reference"
OUT = os.environ.get("MMGL_GOLDEN_OUT", os.path.dirname(os.path.abspath(__file__)))


def load_ref(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    sys.modules[name] = m  # HF PreTrainedModel.__init__ looks the module up in sys.modules
    spec.loader.exec_module(m)
    return m


def save(name, obj):
    path = os.path.join(OUT, name + ".pt")
    torch.save(obj, path)
    print(f"wrote {path} ({os.path.getsize(path)/1024:.1f} KiB)")


def sd(module):
    return {k: v.detach().clone() for k, v in module.state_dict().items()}


def tiny_opt_config(pre_ln=True, hidden=64, layers=4, heads=4, ffn=128, vocab=512):
    from transformers import OPTConfig
    return OPTConfig(vocab_size=vocab, hidden_size=hidden, num_hidden_layers=layers, ffn_dim=ffn,
                     num_attention_heads=heads, max_position_embeddings=200, word_embed_proj_dim=hidden,
                     do_layer_norm_before=pre_ln, dropout=0.0, attention_dropout=0.0)


def mpt_args(**kw):
    base = dict(neighbor_layer_wise=2, neighbor_mode="cross_attention", peft_type="flamingo",
                lora_r=64, lora_alpha=1, lora_dropout=0.0)
    base.update(kw)
    return types.SimpleNamespace(**base)


def layer_case(xa, name, pre_ln, seed, B, S, NK, H, heads, ffn):
    """one gated cross-attention layer of the REAL reference: output + all gradients"""
    g = torch.Generator().manual_seed(seed)
    cfg = xa.MPTConfig(mpt_args(), tiny_opt_config(pre_ln, hidden=H, heads=heads, ffn=ffn))
    layer = xa.MPTDecoderLayer(cfg, cross_attention=True)
    for prm in layer.parameters():
        prm.data.normal_(0, 0.08, generator=g)
    for n, prm in layer.named_parameters():
        if "layer_norm.weight" in n:
            prm.data.add_(1.0)
    layer.gating1.data.fill_(0.7)
    layer.gating2.data.fill_(-0.4)
    layer.eval()
    # bf16-representable weights and inputs: the fp32 reference then sees exactly what the bf16 kernels see
    for prm in layer.parameters():
        if prm.dim() == 2:
            prm.data.copy_(prm.data.bfloat16().float())
    r16 = lambda t: t.bfloat16().float()
    x = r16(torch.randn(B, S, H, generator=g)).requires_grad_(True)
    bank = r16(torch.randn(B, NK, H, generator=g)).requires_grad_(True)
    mask = torch.rand(B, NK, generator=g) > 0.3
    mask[0, :] = True
    mask[1, -5:] = False
    add = xa._expand_mask(mask, x.dtype, tgt_len=S)
    y = layer(x, neighbor_embeds=bank, neighbor_attention_mask=add)[0]
    w = r16(torch.randn(B, S, H, generator=g))
    (y * w).sum().backward()
    save(name, dict(
        cfg=dict(num_heads=heads, do_layer_norm_before=pre_ln), state=sd(layer), x=x.detach(), bank=bank.detach(),
        mask=mask, y=y.detach(), w=w, dx=x.grad.clone(), dbank=bank.grad.clone(),
        grads={k: v.grad.clone() for k, v in layer.named_parameters()}))


def wrapper_cross_d64(xa):
    """full reference CrossAttentionModel at head_dim 64 with bf16-representable weights: loss, logits, bank and the
    gradients of every trainable parameter (the CUDA model-level parity test compares against these)."""
    from transformers import (CLIPVisionConfig, CLIPVisionModel, OPTForCausalLM, RobertaConfig, RobertaModel)
    tmp = tempfile.mkdtemp(prefix="mmgl_golden_d64_")
    d_lm, d_txt, d_vis = (os.path.join(tmp, n) for n in ("opt", "roberta", "clipv"))
    torch.manual_seed(71)
    lm_cfg = tiny_opt_config(True, hidden=128, layers=4, heads=2, ffn=256, vocab=512)
    txt_cfg = RobertaConfig(vocab_size=512, hidden_size=32, num_hidden_layers=2, num_attention_heads=2,
                            intermediate_size=64, max_position_embeddings=40, pad_token_id=1)
    vis_cfg = CLIPVisionConfig(hidden_size=32, intermediate_size=64, num_hidden_layers=2, num_attention_heads=2,
                               image_size=32, patch_size=16)
    OPTForCausalLM(lm_cfg).save_pretrained(d_lm)
    RobertaModel(txt_cfg).save_pretrained(d_txt)
    CLIPVisionModel(vis_cfg).save_pretrained(d_vis)
    args = mpt_args(context="all", n_text_tokens=2, n_visual_tokens=2, model_name_or_path=d_lm, text_model=d_txt,
                    visual_model=d_vis, max_output_length=16, freeze_lm=False)
    model = xa.CrossAttentionModel(args, tokenizer=None)
    g = torch.Generator().manual_seed(72)
    for n, prm in model.named_parameters():
        if "gating" in n:
            prm.data.fill_(0.5 if "gating1" in n else -0.6)
        elif "neighbor_layers" in n and prm.dim() == 2:
            prm.data.normal_(0, 0.05, generator=g)     # init_std 0.02 gives a nearly dead branch at this size
        if prm.is_floating_point() and not n.startswith(("text_model.", "visual_model.")):
            prm.data.copy_(prm.data.bfloat16().float())
    model.lm.lm_head.weight = model.lm.model.decoder.embed_tokens.weight   # tie
    model.eval()
    B, S, T, I, L = 2, 24, 3, 2, 12
    ids = torch.randint(4, 512, (B, S), generator=g)
    am = torch.ones(B, S, dtype=torch.long)
    am[0, -3:] = 0
    ids[0, -3:] = 1
    am[1, 10:13] = 0          # padding in the middle (input segment right-padded before the output segment)
    ids[1, 10:13] = 1
    nids = torch.randint(4, 512, (B, T, L), generator=g)
    nam = torch.ones(B, T, L, dtype=torch.long)
    nam[:, :, -4:] = 0
    batch = dict(
        input_ids=ids, attention_mask=am, labels=ids.clone(),
        neighbor_input_ids=nids, neighbor_attention_mask=nam,
        neighbor_pos_ids=torch.tensor([[1, 2, 3], [1, 2, 0]]),
        text_locations=torch.tensor([[0, 2, 4], [0, 2, 3]]),
        neighbor_images=torch.randn(B, I, 3, 32, 32, generator=g),
        neighbor_images_pos_ids=torch.tensor([[1, 2], [1, 0]]),
        image_locations=torch.tensor([[1, 3], [1, 4]]),
    )
    cap = {}

    def pool_hook(m, i, o):
        o2 = o.detach().bfloat16().float()           # bf16-representable pooled features, fed on to the projections
        cap["text_pooled"] = o2
        return o2 + (o - o.detach())                 # keep the autograd edge to the pooler

    def vis_hook(m, i, o):
        o.pooler_output = o.pooler_output.detach().bfloat16().float()
        cap["visual_pooled"] = o.pooler_output
        return o
    h1 = model.text_pooler.register_forward_hook(pool_hook)
    h2 = model.visual_model.register_forward_hook(vis_hook)
    h3 = model.lm.register_forward_pre_hook(lambda m, a, kw: cap.__setitem__("lm_kwargs", {k: v.detach() for k, v in kw.items() if v is not None}), with_kwargs=True)
    out = model(**batch)
    out.loss.backward()
    for h in (h1, h2, h3):
        h.remove()
    state = {k: v for k, v in sd(model).items() if not k.startswith(("text_model.", "visual_model."))}
    grads = {n: prm.grad.clone() for n, prm in model.named_parameters()
             if prm.grad is not None and not n.startswith("text_pooler.") and n != "lm.lm_head.weight"}
    save("wrapper_cross_d64", dict(
        cfg=dict(num_heads=2, num_layers=4, neighbor_layer_wise=2, do_layer_norm_before=True, n_tokens=2),
        lm_config=lm_cfg.to_dict(), text_config=txt_cfg.to_dict(), visual_config=vis_cfg.to_dict(),
        state=state, batch=batch, text_pooled=cap["text_pooled"].reshape(B, T, -1),
        visual_pooled=cap["visual_pooled"].reshape(B, I, -1), bank=cap["lm_kwargs"]["neighbor_embeds"],
        bank_mask=cap["lm_kwargs"]["neighbor_attention_mask"], loss=out.loss.detach(), logits=out.logits.detach(),
        grads=grads))


def main():
    torch.manual_seed(0)
    xa = load_ref("ref_xattn", f"{REF}/model/modelling_cross_attention.py")
    wrapper_cross_d64(xa)

    # ---- case 1b: the same layer at head_dim 64 (the CUDA kernels support d in {64,128}) -------
    for pre_ln in (True, False):
        layer_case(xa, f"xattn_layer_d64_{'pre' if pre_ln else 'post'}ln", pre_ln, 61 + int(pre_ln),
                   B=2, S=40, NK=24, H=128, heads=2, ffn=256)
    if "--only-d64" in sys.argv:
        return

    # ---- case 1: one gated cross-attention layer (pre-LN and post-LN) -----------------------
    for pre_ln in (True, False):
        g = torch.Generator().manual_seed(11 + int(pre_ln))
        cfg = xa.MPTConfig(mpt_args(), tiny_opt_config(pre_ln))
        layer = xa.MPTDecoderLayer(cfg, cross_attention=True)
        for prm in layer.parameters():
            prm.data.normal_(0, 0.08, generator=g)
        layer.gating1.data.fill_(0.7)
        layer.gating2.data.fill_(-0.4)
        layer.eval()
        B, S, NK, H = 2, 24, 16, 64
        x = torch.randn(B, S, H, generator=g, requires_grad=True)
        bank = torch.randn(B, NK, H, generator=g, requires_grad=True)
        mask = torch.rand(B, NK, generator=g) > 0.3
        mask[0, :] = True
        mask[1, -5:] = False
        add = xa._expand_mask(mask, x.dtype, tgt_len=S)
        y = layer(x, neighbor_embeds=bank, neighbor_attention_mask=add)[0]
        w = torch.randn(B, S, H, generator=g)
        (y * w).sum().backward()
        save(f"xattn_layer_{'pre' if pre_ln else 'post'}ln", dict(
            cfg=dict(num_heads=4, do_layer_norm_before=pre_ln), state=sd(layer), x=x.detach(), bank=bank.detach(),
            mask=mask, y=y.detach(), w=w, dx=x.grad.clone(), dbank=bank.grad.clone(),
            grads={k: v.grad.clone() for k, v in layer.named_parameters()}))

    # ---- case 2: MPTForCausalLM tiny (interleave loop + loss) --------------------------------
    g = torch.Generator().manual_seed(21)
    torch.manual_seed(20)   # the weights come from the global RNG: seed it here, not from whatever case 1 left behind
    cfg = xa.MPTConfig(mpt_args(), tiny_opt_config(True))
    lm = xa.MPTForCausalLM(cfg)
    for n, prm in lm.named_parameters():
        if "gating" in n:
            prm.data.fill_(0.5)
    lm.eval()
    B, S, NK = 2, 20, 12
    ids = torch.randint(4, 512, (B, S), generator=g)
    am = torch.ones(B, S, dtype=torch.long)
    am[1, -6:] = 0
    ids[1, -6:] = 1
    bank = torch.randn(B, NK, 64, generator=g)
    bmask = torch.ones(B, NK, dtype=torch.bool)
    bmask[0, -4:] = False
    out = lm(input_ids=ids, attention_mask=am, labels=ids, neighbor_embeds=bank, neighbor_attention_mask=bmask)
    # invariant I1: with the gates at their init value 0.0 the bank has no influence (== plain OPT)
    gates = {n: prm.data.clone() for n, prm in lm.named_parameters() if "gating" in n}
    for n, prm in lm.named_parameters():
        if "gating" in n:
            prm.data.zero_()
    out_nb = lm(input_ids=ids, attention_mask=am, labels=ids, neighbor_embeds=bank, neighbor_attention_mask=bmask)
    for n, prm in lm.named_parameters():
        if "gating" in n:
            prm.data.copy_(gates[n])
    save("mpt_lm", dict(cfg=dict(num_heads=4, num_layers=4, neighbor_layer_wise=2, do_layer_norm_before=True),
                        state=sd(lm), input_ids=ids, attention_mask=am, bank=bank, bank_mask=bmask,
                        loss=out.loss.detach(), logits=out.logits.detach(),
                        loss_gates0=out_nb.loss.detach(), logits_gates0=out_nb.logits.detach()))

    # ---- case 3: full CrossAttentionModel wrapper with tiny local encoders --------------------
    from transformers import (CLIPVisionConfig, CLIPVisionModel, OPTForCausalLM, RobertaConfig, RobertaModel)
    tmp = tempfile.mkdtemp(prefix="mmgl_golden_")
    d_lm, d_txt, d_vis = (os.path.join(tmp, n) for n in ("opt", "roberta", "clipv"))
    torch.manual_seed(31)
    OPTForCausalLM(tiny_opt_config(True)).save_pretrained(d_lm)
    RobertaModel(RobertaConfig(vocab_size=512, hidden_size=32, num_hidden_layers=2, num_attention_heads=2,
                               intermediate_size=64, max_position_embeddings=40, pad_token_id=1)).save_pretrained(d_txt)
    CLIPVisionModel(CLIPVisionConfig(hidden_size=32, intermediate_size=64, num_hidden_layers=2,
                                     num_attention_heads=2, image_size=32, patch_size=16)).save_pretrained(d_vis)
    args = mpt_args(context="all", n_text_tokens=2, n_visual_tokens=2, model_name_or_path=d_lm, text_model=d_txt,
                    visual_model=d_vis, max_output_length=16, freeze_lm=False)
    model = xa.CrossAttentionModel(args, tokenizer=None)
    for n, prm in model.named_parameters():
        if "gating" in n:
            prm.data.fill_(0.5)
    model.eval()  # returns None in the reference
    g = torch.Generator().manual_seed(32)
    B, S, T, I, L = 2, 20, 3, 2, 12
    ids = torch.randint(4, 512, (B, S), generator=g)
    am = torch.ones(B, S, dtype=torch.long)
    am[0, -3:] = 0
    ids[0, -3:] = 1
    nids = torch.randint(4, 512, (B, T, L), generator=g)
    nam = torch.ones(B, T, L, dtype=torch.long)
    nam[:, :, -4:] = 0
    batch = dict(
        input_ids=ids, attention_mask=am, labels=ids.clone(),
        neighbor_input_ids=nids, neighbor_attention_mask=nam,
        neighbor_pos_ids=torch.tensor([[1, 2, 3], [1, 2, 0]]),
        text_locations=torch.tensor([[0, 2, 4], [0, 2, 3]]),
        neighbor_images=torch.randn(B, I, 3, 32, 32, generator=g),
        neighbor_images_pos_ids=torch.tensor([[1, 2], [1, 0]]),
        image_locations=torch.tensor([[1, 3], [1, 4]]),
    )
    cap = {}
    h1 = model.text_pooler.register_forward_hook(lambda m, i, o: cap.__setitem__("text_pooled", o.detach()))
    h2 = model.visual_model.register_forward_hook(lambda m, i, o: cap.__setitem__("visual_pooled", o.pooler_output.detach()))
    h3 = model.lm.register_forward_pre_hook(lambda m, a, kw: cap.__setitem__("lm_kwargs", {k: v.detach() for k, v in kw.items() if v is not None}), with_kwargs=True)
    out = model(**batch)
    for h in (h1, h2, h3):
        h.remove()
    state = {k: v for k, v in sd(model).items() if not k.startswith(("text_model.", "visual_model."))}
    save("wrapper_cross", dict(
        cfg=dict(num_heads=4, num_layers=4, neighbor_layer_wise=2, do_layer_norm_before=True, n_tokens=2),
        state=state, batch=batch, text_pooled=cap["text_pooled"].reshape(B, T, -1),
        visual_pooled=cap["visual_pooled"].reshape(B, I, -1), bank=cap["lm_kwargs"]["neighbor_embeds"],
        bank_mask=cap["lm_kwargs"]["neighbor_attention_mask"], loss=out.loss.detach(), logits=out.logits.detach()))

    # ---- case 4: SelfAttentionModel concat path (peft stubbed for import; peft_type=none) ------
    peft = types.ModuleType("peft")
    for n in ["LoraConfig", "PrefixTuningConfig", "PromptTuningInit", "PromptTuningConfig", "TaskType", "get_peft_model"]:
        setattr(peft, n, types.SimpleNamespace(SEQ_2_SEQ_LM="s2s", CAUSAL_LM="clm", RANDOM="r"))
    sys.modules["peft"] = peft
    pkg = types.ModuleType("refmodel")
    pkg.__path__ = [f"{REF}/model"]
    sys.modules["refmodel"] = pkg
    gr = load_ref("refmodel.graph", f"{REF}/model/graph.py")
    sa = load_ref("refmodel.modelling_self_attention", f"{REF}/model/modelling_self_attention.py")
    del sys.modules["peft"]

    from transformers import T5Config, T5ForConditionalGeneration
    d_t5 = os.path.join(tmp, "t5")
    torch.manual_seed(41)
    T5ForConditionalGeneration(T5Config(vocab_size=512, d_model=64, d_kv=16, d_ff=128, num_layers=2,
                                        num_decoder_layers=2, num_heads=4, decoder_start_token_id=0)).save_pretrained(d_t5)
    N = T + I
    for lm_kind, d_model_dir, dec_only in (("t5", d_t5, False), ("opt", d_lm, True)):
        for pt in ("none", "laplacian", "gnn"):
            torch.manual_seed(42)
            a = types.SimpleNamespace(context="all", decoder_only=dec_only, neighbor_mode="embedding", position_type=pt,
                                      n_text_tokens=2, n_visual_tokens=2, model_name_or_path=d_model_dir, peft_type="none",
                                      text_model=d_txt, visual_model=d_vis, max_output_length=16, freeze_lm=False,
                                      max_text_neighbors=T, max_image_neighbors=I, lora_r=8, lora_alpha=1, lora_dropout=0.0)
            m = sa.SelfAttentionModel(a, tokenizer=None)
            m.eval()
            g = torch.Generator().manual_seed(43)
            b2 = dict(batch)
            if not dec_only:
                lab = torch.randint(2, 512, (B, 8), generator=g)
                lab[1, -2:] = -100
                b2["labels"] = lab
            else:
                b2["labels"] = ids.clone()
            k = 1 + T + I - 5
            if pt == "laplacian":
                b2["lpe"] = torch.randn(B, N + 1, k, generator=g)
            if pt == "gnn":
                adj = (torch.rand(B, N + 1, N + 1, generator=g) > 0.5).float()
                adj = adj + adj.transpose(1, 2) + torch.eye(N + 1)
                b2["graph"] = adj / adj.sum(-1, keepdim=True)
            cap = {}
            h1 = m.text_pooler.register_forward_hook(lambda mm, i, o: cap.__setitem__("text_pooled", o.detach()))
            h2 = m.visual_model.register_forward_hook(lambda mm, i, o: cap.__setitem__("visual_pooled", o.pooler_output.detach()))
            h3 = m.lm.register_forward_pre_hook(lambda mm, aa, kw: cap.__setitem__("lm_kwargs", {kk: v.detach().clone() for kk, v in kw.items() if v is not None}), with_kwargs=True)
            out = m(**{kk: (v.clone() if torch.is_tensor(v) else v) for kk, v in b2.items()})
            for h in (h1, h2, h3):
                h.remove()
            state = {kk: v for kk, v in sd(m).items() if not kk.startswith(("text_model.", "visual_model.", "lm."))}
            state["input_embeddings.weight"] = m.input_embeddings.weight.detach().clone()
            save(f"wrapper_self_{lm_kind}_{pt}", dict(
                cfg=dict(n_tokens=2, decoder_only=dec_only, position_type=pt), state=state, batch=b2,
                text_pooled=cap["text_pooled"].reshape(B, T, -1), visual_pooled=cap["visual_pooled"].reshape(B, I, -1),
                inputs_embeds=cap["lm_kwargs"]["inputs_embeds"], attention_mask=cap["lm_kwargs"]["attention_mask"],
                labels=cap["lm_kwargs"]["labels"], loss=out.loss.detach(), logits_shape=tuple(out.logits.shape)))

    # ---- case 5: GCN alone with gradients --------------------------------------------------
    g = torch.Generator().manual_seed(51)
    gcn = gr.GCN(input_dim=48, output_dim=48, hidden_dim=24)
    B, N = 3, 6
    X = torch.randn(B, N, 48, generator=g, requires_grad=True)
    adj = (torch.rand(B, N + 1, N + 1, generator=g) > 0.6).float() + torch.eye(N + 1)
    adj = adj / adj.sum(-1, keepdim=True)
    Y = gcn(X, adj)
    w = torch.randn(B, N, 48, generator=g)
    (Y * w).sum().backward()
    save("gcn", dict(state=sd(gcn), x=X.detach(), adj=adj, y=Y.detach(), w=w, dx=X.grad.clone(),
                     grads={k: v.grad.clone() for k, v in gcn.named_parameters()}))


if __name__ == "__main__":
    main()
