"""Model-level parity on the GPU: mmgl_b200.CrossAttentionModel (CUDA kernels) against outputs and gradients of the
REAL reference CrossAttentionModel (tests/golden/wrapper_cross_d64.pt, made by tests/golden/make_golden.py from
/root/reference), downstream of the frozen encoders (their pooled features are part of the fixture).

Tolerances (bf16 compute vs the fp32 reference on bf16-representable weights), set from the errors MEASURED on B200
(profiles/r02_parity_measured.txt): bank 1.9e-3 measured -> 4e-3; byte mask bit-exact; logits 7.0e-3 -> 1.5e-2; loss 1.5e-4
absolute (of 6.27) -> 1e-3; trainable-parameter gradients 4.4e-2 .. 9.1e-2 -> 1.5e-1.  The gradient figure is a property of
the TOY width, not of the kernels: the fixture model is 128 wide with 256 FFN units, its gradients pass through 4 frozen + 2
gated ReLU layers, and one unit whose pre-activation sits within a bf16 ulp of zero takes the other branch -- 1/256 of a
layer each.  At the benchmarked widths the same comparison gives 1.6e-2 (tests/test_gpu_layer.py::
test_gated_cross_layer_at_benchmarked_sizes), which is the tight statement."""
import types

import pytest
import torch

from util import BF16, Report

pytestmark = pytest.mark.gpu


def _build(g, train=False):
    from transformers import CLIPVisionConfig, OPTConfig, RobertaConfig
    from mmgl_b200 import modules as M
    args = types.SimpleNamespace(
        context="all", neighbor_mode="embedding", peft_type="flamingo", n_text_tokens=2, n_visual_tokens=2,
        model_name_or_path=OPTConfig(**g["lm_config"]), text_model=RobertaConfig(**g["text_config"]),
        visual_model=CLIPVisionConfig(**g["visual_config"]), max_output_length=16, freeze_lm=False,
        neighbor_layer_wise=2, lora_r=64, lora_alpha=1, lora_dropout=0.0)
    model = M.CrossAttentionModel(args, tokenizer=None)
    missing, unexpected = model.load_state_dict(g["state"], strict=False)
    assert not unexpected, f"state-dict keys the drop-in does not know: {unexpected}"
    assert all(k.startswith(("text_model.", "visual_model.")) for k in missing), f"missing keys: {missing}"
    M.prepare_for_training(model, "cuda")
    model.train(train)
    model.skip_padding_neighbors = False   # the fixture's pooled features cover every neighbor slot
    # the fixture supplies the frozen encoders' pooled outputs (tiny random encoders are not part of the path)
    tp, vp = g["text_pooled"].cuda(), g["visual_pooled"].cuda()
    model.encode_images = lambda px: vp.reshape(-1, vp.shape[-1]).to(BF16)
    model.encode_text = lambda ids, am: tp.reshape(-1, tp.shape[-1]).to(BF16)
    return model


def test_state_dict_keys_match_reference(golden):
    g = golden("wrapper_cross_d64")
    model = _build(g)
    ours = {k for k in model.state_dict() if not k.startswith(("text_model.", "visual_model."))}
    assert ours == set(g["state"]), (sorted(ours - set(g["state"])), sorted(set(g["state"]) - ours))
    trainable = {n for n, p in model.named_parameters() if p.requires_grad}
    assert all(("neighbor_layers" in n) or n.startswith(("text_embeddings", "visual_embeddings", "text_position",
                                                         "visual_position", "text_pooler")) for n in trainable)
    assert any("gating1" in n for n in trainable)


def test_cross_attention_model_forward_backward_vs_reference(golden):
    g = golden("wrapper_cross_d64")
    model = _build(g)
    batch = {k: v.cuda() for k, v in g["batch"].items()}
    cap = {}
    hook = model.lm.register_forward_pre_hook(lambda m, a, kw: cap.update(kw), with_kwargs=True)
    out = model(**batch)
    hook.remove()
    rep = Report()
    assert torch.equal(cap["neighbor_attention_mask"].bool().cpu(), g["bank_mask"]), "bank mask must be bit-exact"
    rep.close("bank", cap["neighbor_embeds"], g["bank"], 4e-3)
    rep.close("logits", out.logits, g["logits"], 1.5e-2)
    rep.scalar("loss", out.loss, g["loss"], 0.0, 1e-3)
    out.loss.backward()
    params = dict(model.named_parameters())
    for k, gr in g["grads"].items():
        assert params[k].grad is not None, f"{k} received no gradient"
        if gr.numel() == 1:
            rep.scalar("d " + k, params[k].grad, gr, 1.5e-1, 1e-3)
        elif k.endswith("k_proj.bias"):
            rep.absolute("d " + k, params[k].grad, gr, 1e-4)      # analytically zero
        else:
            rep.close("d " + k, params[k].grad, gr, 1.5e-1)
    frozen_with_grad = [n for n, p in params.items() if not p.requires_grad and p.grad is not None]
    assert not frozen_with_grad
    rep.finish()


def test_gates_at_zero_equal_plain_opt(golden):
    """Invariant I1: with gating == 0 (the reference's init) the bank has no influence."""
    g = golden("wrapper_cross_d64")
    model = _build(g)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if "gating" in n:
                p.zero_()
    batch = {k: v.cuda() for k, v in g["batch"].items()}
    with torch.no_grad():
        a = model(**batch).logits
        model.neighbor_mode = "raw"          # the reference's sanity-check branch: pure OPT
        b = model(**batch).logits
    assert torch.equal(a, b)


def test_gradients_at_the_reference_init(golden):
    """Invariant I4: at the reference's init (gating1 = gating2 = 0) the loss gradient reaches the
    gates, while every weight inside the gated branches gets an exactly-zero gradient (tanh(0) multiplies the branch)."""
    g = golden("wrapper_cross_d64")
    model = _build(g)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if "gating" in n:
                p.zero_()
    batch = {k: v.cuda() for k, v in g["batch"].items()}
    model(**batch).loss.backward()
    params = dict(model.named_parameters())
    gates = [n for n in params if "gating" in n]
    assert gates and all(params[n].grad is not None and float(params[n].grad.abs()) > 0 for n in gates)
    inside = [n for n in params if "neighbor_layers" in n and "gating" not in n and "layer_norm" not in n]
    assert inside
    for n in inside:
        assert params[n].grad is not None and float(params[n].grad.abs().max()) == 0.0, n


def test_training_mode_runs_with_dropout(golden):
    g = golden("wrapper_cross_d64")
    model = _build(g, train=True)
    for m in model.modules():
        if hasattr(m, "dropout") and isinstance(m.dropout, float):
            m.dropout = 0.1
    batch = {k: v.cuda() for k, v in g["batch"].items()}
    torch.manual_seed(0)
    l1 = model(**batch).loss
    torch.manual_seed(0)
    model.zero_grad()
    out = model(**batch)
    out.loss.backward()
    assert torch.isfinite(out.loss) and abs(float(out.loss) - float(g["loss"])) < 0.5
    assert float(l1) != float(out.loss) or True   # masks are counter-based per call; just exercise the path
    assert all(torch.isfinite(p.grad).all() for p in model.parameters() if p.grad is not None)


def test_skipping_padding_neighbors_changes_nothing(golden):
    """Row f2: the frozen encoders are not run on padding neighbors (pos_id == 0).  Loss and every gradient must be
    the same as with the full (reference) computation; tolerance 2e-3 rel-L2 only because the encoder GEMMs see a
    different batch size (library kernel selection), the masked rows themselves contribute exactly zero."""
    g = golden("wrapper_cross_d64")
    from transformers import CLIPVisionConfig, OPTConfig, RobertaConfig
    from mmgl_b200 import modules as M
    # the frozen encoders run on the package's kernels (head_dim 64; the fixture's own encoders have head_dim 16, which the
    # product rejects), so this test builds its own small towers and uses the fixture for the LM weights and the batch only
    txt = RobertaConfig(vocab_size=512, hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256,
                        max_position_embeddings=40, pad_token_id=1)
    vis = CLIPVisionConfig(hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=2,
                           image_size=32, patch_size=16)
    args = types.SimpleNamespace(
        context="all", neighbor_mode="embedding", peft_type="flamingo", n_text_tokens=2, n_visual_tokens=2,
        model_name_or_path=OPTConfig(**g["lm_config"]), text_model=txt, visual_model=vis, max_output_length=16,
        freeze_lm=False, neighbor_layer_wise=2, lora_r=64, lora_alpha=1, lora_dropout=0.0)
    torch.manual_seed(5)
    model = M.CrossAttentionModel(args, tokenizer=None)
    model.load_state_dict({k: v for k, v in g["state"].items() if k.startswith("lm.")}, strict=False)
    with torch.no_grad():
        for n, prm in model.named_parameters():
            if "gating" in n:
                prm.fill_(0.5)
    M.prepare_for_training(model, "cuda").eval()
    batch = {k: v.cuda() for k, v in g["batch"].items()}
    batch["neighbor_pos_ids"] = torch.tensor([[1, 2, 0], [1, 0, 0]]).cuda()         # padding neighbors present
    batch["neighbor_images_pos_ids"] = torch.tensor([[1, 0], [0, 0]]).cuda()
    res = {}
    for skip in (False, True):
        model.skip_padding_neighbors = skip
        model.zero_grad()
        out = model(**batch)
        out.loss.backward()
        res[skip] = (out.loss.detach().clone(), out.logits.detach().clone(),
                     {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None})
    rep = Report()
    rep.scalar("loss", res[True][0], res[False][0], 0.0, 1e-4)
    rep.close("logits", res[True][1], res[False][1], 2e-3)
    assert res[True][2].keys() == res[False][2].keys()
    for n in res[True][2]:
        if n.endswith("k_proj.bias"):
            continue
        rep.close("d " + n, res[True][2][n], res[False][2][n], 2e-3)
    rep.finish()


@pytest.mark.parametrize("position_type", ["laplacian", "gnn"])
def test_cross_attention_model_with_graph_position_encodings_vs_oracle(golden, position_type):
    """BASELINE configs[3] (cfg4: flamingo + graph positional encodings) is an EXTENSION of the reference (its
    CrossAttentionModel.forward has no lpe / graph kwargs): the bank gets the same ``lpe_embeddings`` / ``GCN`` add the
    self-attention wrapper applies (model/modelling_self_attention.py:311-320) before the gated cross-attention layers.
    Oracle: oracle.cross_attention_model_from_pooled with the lpe / graph terms (lpe_add / gnn_add are pinned against the
    reference's SelfAttentionModel by the wrapper_self_*_{laplacian,gnn} fixtures), on the wrapper_cross_d64 weights."""
    from transformers import CLIPVisionConfig, OPTConfig, RobertaConfig
    from mmgl_b200 import modules as M
    from oracle import mmgl_oracle as O
    g = golden("wrapper_cross_d64")
    t, i = g["batch"]["neighbor_pos_ids"].shape[1], g["batch"]["neighbor_images_pos_ids"].shape[1]
    n, b = t + i, g["batch"]["input_ids"].shape[0]
    args = types.SimpleNamespace(
        context="all", neighbor_mode="embedding", peft_type="flamingo", n_text_tokens=2, n_visual_tokens=2,
        model_name_or_path=OPTConfig(**g["lm_config"]), text_model=RobertaConfig(**g["text_config"]),
        visual_model=CLIPVisionConfig(**g["visual_config"]), max_output_length=16, freeze_lm=False,
        neighbor_layer_wise=2, lora_r=64, lora_alpha=1, lora_dropout=0.0, position_type=position_type,
        max_text_neighbors=t, max_image_neighbors=i)
    model = M.CrossAttentionModel(args, tokenizer=None)
    missing, unexpected = model.load_state_dict(g["state"], strict=False)
    assert not unexpected
    gen = torch.Generator().manual_seed(123)
    extra = {}
    with torch.no_grad():
        for name, prm in model.named_parameters():
            if name.startswith(("lpe_embeddings.", "gnn.")):
                prm.copy_((torch.randn(prm.shape, generator=gen) * 0.05).to(BF16).float())
                extra[name] = prm.detach().clone()
    assert extra, "the graph-PE parameters were not created"
    M.prepare_for_training(model, "cuda")
    model.eval()
    model.skip_padding_neighbors = False
    tp, vp = g["text_pooled"].cuda(), g["visual_pooled"].cuda()
    model.encode_images = lambda px: vp.reshape(-1, vp.shape[-1]).to(BF16)
    model.encode_text = lambda ids, am: tp.reshape(-1, tp.shape[-1]).to(BF16)
    batch = dict(g["batch"])
    if position_type == "laplacian":
        lpe = torch.randn((b, n + 1, n - 4), generator=gen)
        batch["lpe"] = lpe / lpe.norm(dim=1, keepdim=True).clamp_min(1e-6)
    else:
        a = (torch.rand((b, n + 1, n + 1), generator=gen) < 0.5).float()
        a = ((a + a.transpose(1, 2)) > 0).float() + torch.eye(n + 1)
        batch["graph"] = a / a.sum(-1, keepdim=True)
    out = model(**{k: v.cuda() for k, v in batch.items()})
    out.loss.backward()

    p = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in {**g["state"], **extra}.items()}
    cfg = dict(g["cfg"], flamingo=True)
    loss, logits = O.cross_attention_model_from_pooled(p, cfg, batch, g["text_pooled"], g["visual_pooled"])
    loss.backward()
    assert abs(float(loss) - float(g["loss"])) > 1e-4, "the graph PE term must change the loss (path is live)"
    rep = Report()
    rep.close("logits", out.logits, logits, 2e-2)
    rep.scalar("loss", out.loss, loss, 0.0, 2e-2)
    params = dict(model.named_parameters())
    for k in list(extra) + ["text_embeddings.weight", "visual_embeddings.weight",
                            "lm.model.decoder.neighbor_layers.0.self_attn.v_proj.weight",
                            "lm.model.decoder.neighbor_layers.1.fc2.weight"]:
        assert params[k].grad is not None, f"{k} received no gradient"
        rep.close("d " + k, params[k].grad, p[k].grad, 8e-2)
    rep.finish()


def test_host_plan_changes_nothing(golden):
    """mmgl_b200.plan (the data pipeline's ragged-neighbor bookkeeping, shipped with the batch) removes the per-step host
    syncs; loss, logits and gradients must be IDENTICAL to the run that reads the sizes back from the device."""
    from transformers import CLIPVisionConfig, OPTConfig, RobertaConfig
    from mmgl_b200 import modules as M, plan, synth
    g = golden("wrapper_cross_d64")
    txt = RobertaConfig(vocab_size=512, hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256,
                        max_position_embeddings=80, pad_token_id=1)
    vis = CLIPVisionConfig(hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=2,
                           image_size=32, patch_size=16)
    args = types.SimpleNamespace(
        context="all", neighbor_mode="embedding", peft_type="flamingo", n_text_tokens=2, n_visual_tokens=2,
        model_name_or_path=OPTConfig(**g["lm_config"]), text_model=txt, visual_model=vis, max_output_length=16,
        freeze_lm=False, neighbor_layer_wise=2, lora_r=64, lora_alpha=1, lora_dropout=0.0)
    torch.manual_seed(7)
    model = M.CrossAttentionModel(args, tokenizer=None)
    with torch.no_grad():
        for n, prm in model.named_parameters():
            if "gating" in n:
                prm.fill_(0.5)
    M.prepare_for_training(model, "cuda").eval()
    spec = synth.BatchSpec(batch=4, max_input_length=48, max_output_length=16, text_neighbors=5, image_neighbors=3,
                           vocab_size=512, neighbor_vocab_size=512, image_size=32)
    host = synth.make_batch(spec, seed=11)
    res = []
    for with_plan in (False, True):
        batch = plan.attach_plan(host) if with_plan else host
        batch = {k: v.cuda() for k, v in batch.items()}
        model.zero_grad()
        out = model(**batch)
        out.loss.backward()
        res.append((out.loss.detach().clone(), out.logits.detach().clone(),
                    {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}))
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])
    assert res[0][2].keys() == res[1][2].keys() and all(torch.equal(res[0][2][k], res[1][2][k]) for k in res[0][2])
