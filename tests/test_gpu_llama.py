"""Llama-2 + gated cross-attention (BASELINE configs[4]) on the GPU.

There is no reference Llama wrapper (language_modelling/run_generation.py:286-301 dispatches t5 / opt / mpt only), so the
oracle is assembled as SURVEY 8f defines it: HF ``LlamaForCausalLM`` in fp32 (library code present on both boxes), run layer
by layer, with the ORACLE's restatement of the reference's gated layer (oracle.mpt_decoder_layer, pinned to
model/modelling_cross_attention.py:304-375 by the golden fixtures) inserted after every ``neighbor_layer_wise``-th layer.

Tolerances (bf16 kernels vs fp32 oracle, toy width 256; measured values in profiles/r02_parity_measured.txt): logits 1.5e-2,
loss 5e-3 absolute, gradients 8e-2 (ReLU flips of a 512-unit gated FFN, see tests/test_gpu_model.py)."""
import types

import pytest
import torch

from util import BF16, Report, randn

pytestmark = pytest.mark.gpu


def _tiny_llama():
    from transformers import LlamaConfig, LlamaForCausalLM
    cfg = LlamaConfig(vocab_size=384, hidden_size=256, intermediate_size=512, num_hidden_layers=4, num_attention_heads=2,
                      num_key_value_heads=2, max_position_embeddings=256, rms_norm_eps=1e-5, tie_word_embeddings=False,
                      pad_token_id=0, attn_implementation="eager")
    torch.manual_seed(0)
    return LlamaForCausalLM(cfg)


def test_rope_and_swiglu_match_hf_functions():
    """the two elementwise kernels against HF's own apply_rotary_pos_emb and LlamaMLP arithmetic (fp32 on bf16 inputs)"""
    from transformers.models.llama.modeling_llama import apply_rotary_pos_emb
    from mmgl_b200 import llama as L, ops
    lm = _tiny_llama().cuda()
    gen = torch.Generator().manual_seed(1)
    b, s, heads, d = 2, 70, 2, 128
    h = heads * d
    qkv = randn(gen, b, s, 3 * h).to(BF16)
    x = qkv.clone().requires_grad_(True)
    cs = L.rope_table(lm, s, x.device)
    y = ops.rope_qk(x, cs, heads)
    w = randn(gen, b, s, 3 * h).to(BF16)
    y.backward(w)
    xr = qkv.float().requires_grad_(True)
    q, k, v = (t.reshape(b, s, heads, d).transpose(1, 2) for t in xr.split(h, dim=-1))
    cos, sin = lm.model.rotary_emb(torch.zeros(1, device="cuda"), position_ids=torch.arange(s, device="cuda")[None])
    qe, ke = apply_rotary_pos_emb(q, k, cos.float(), sin.float())
    yr = torch.cat([t.transpose(1, 2).reshape(b, s, h) for t in (qe, ke, v)], dim=-1)
    yr.backward(w.float())
    rep = Report()
    rep.close("rope(q|k), v untouched", y, yr, 4e-3)
    rep.close("d qkv", x.grad, xr.grad, 4e-3)
    assert torch.equal(y[..., 2 * h:], qkv[..., 2 * h:]), "the V third must pass through bit-exactly"
    gu = randn(gen, 50, 2 * 512).to(BF16)
    g = gu.clone().requires_grad_(True)
    hh = ops.swiglu(g)
    dh = randn(gen, 50, 512).to(BF16)
    hh.backward(dh)
    gr = gu.float().requires_grad_(True)
    hr = torch.nn.functional.silu(gr[:, :512]) * gr[:, 512:]
    hr.backward(dh.float())
    rep.close("swiglu", hh, hr, 4e-3)
    rep.close("d gu", g.grad, gr.grad, 4e-3)
    rep.finish()


def _model(nlw=2):
    from mmgl_b200 import llama as L
    args = types.SimpleNamespace(neighbor_layer_wise=nlw, peft_type="flamingo", neighbor_dropout=0.0)
    lm = L.GatedLlamaForCausalLM(_tiny_llama(), args).cuda()
    for p in lm.model.parameters():
        p.data = p.data.to(BF16).float()          # bf16-representable frozen weights: both sides see the same values
    gen = torch.Generator().manual_seed(5)
    with torch.no_grad():
        for n, p in lm.neighbor_layers.named_parameters():
            if "gating" not in n:
                p.copy_((torch.randn(p.shape, generator=gen) * (0.05 if p.dim() == 2 else 0.02)).to(BF16).float())
    return lm


def _inputs():
    gen = torch.Generator().manual_seed(9)
    b, s, nk = 2, 80, 24
    ids = torch.randint(4, 384, (b, s), generator=gen)
    am = torch.ones(b, s, dtype=torch.long)
    am[1, 60:] = 0
    ids[1, 60:] = 0
    bank = (torch.randn(b, nk, 256, generator=gen) * 0.5).to(BF16).float()
    bmask = torch.rand(b, nk, generator=gen) > 0.3
    bmask[:, 0] = True
    return ids, am, bank, bmask


def test_gates_at_zero_equal_hf_llama():
    """Invariant I1 for the Llama wrapper: at the reference init (gating = 0) the gated layers vanish and logits / loss are
    HF LlamaForCausalLM's own (fp32 forward)."""
    lm = _model()
    ids, am, bank, bmask = _inputs()
    lm.eval()
    out = lm(input_ids=ids.cuda(), attention_mask=am.cuda(), labels=ids.cuda(), neighbor_embeds=bank.cuda().to(BF16),
             neighbor_attention_mask=bmask.cuda())
    with torch.no_grad():
        ref = lm.model(input_ids=ids.cuda(), attention_mask=am.cuda(), labels=ids.cuda())
    rep = Report()
    valid = am.bool().cuda()
    rep.close("logits (real positions)", out.logits[valid], ref.logits[valid], 1.5e-2)
    rep.finish()


def test_gated_llama_forward_backward_vs_hf_layers_plus_oracle_gated_layer():
    from oracle import mmgl_oracle as O
    from transformers.masking_utils import create_causal_mask
    lm = _model()
    with torch.no_grad():
        for n, p in lm.named_parameters():
            if "gating" in n:
                p.fill_(0.5)
    ids, am, bank, bmask = _inputs()
    lm.eval()
    bank_g = bank.cuda().to(BF16).requires_grad_(True)
    out = lm(input_ids=ids.cuda(), attention_mask=am.cuda(), labels=ids.cuda(), neighbor_embeds=bank_g,
             neighbor_attention_mask=bmask.cuda())
    out.loss.backward()

    # oracle: HF layers in fp32 + oracle gated layer, on the CPU
    hf = lm.model.float().cpu()
    p = {n: v.detach().float().cpu().clone().requires_grad_(True) for n, v in lm.neighbor_layers.named_parameters()}
    bank_c = bank.clone().requires_grad_(True)
    x = hf.model.embed_tokens(ids)
    pos = torch.arange(ids.shape[1])[None]
    mask = create_causal_mask(config=hf.config, inputs_embeds=x, attention_mask=am, past_key_values=None, position_ids=pos)
    pe = hf.model.rotary_emb(x, position_ids=pos)
    add = O.expand_mask(bmask, torch.float32, ids.shape[1])
    for i, layer in enumerate(hf.model.layers):
        x = layer(x, attention_mask=mask, position_embeddings=pe, position_ids=pos)
        x = x[0] if isinstance(x, tuple) else x
        if (i + 1) % 2 == 0:
            k = (i + 1) // 2 - 1
            x = O.mpt_decoder_layer(x, O.sub(p, f"{k}."), hf.config.num_attention_heads, cross_attention=True, bank=bank_c,
                                    bank_add_mask=add, do_layer_norm_before=True)
    logits = hf.lm_head(hf.model.norm(x))
    loss = torch.nn.functional.cross_entropy(logits[:, :-1].reshape(-1, logits.shape[-1]), ids[:, 1:].reshape(-1))
    loss.backward()
    rep = Report()
    valid = am.bool()
    rep.close("logits (real positions)", out.logits.cpu()[valid], logits[valid], 1.5e-2)
    rep.scalar("loss", out.loss, loss, 0.0, 5e-3)
    rep.close("d bank", bank_g.grad, bank_c.grad, 8e-2)
    gp = dict(lm.neighbor_layers.named_parameters())
    for n in ("0.gating1", "1.gating2"):
        rep.scalar("d " + n, gp[n].grad, p[n].grad, 8e-2, 1e-3)
    for n in ("0.self_attn.q_proj.weight", "0.self_attn.v_proj.weight", "0.fc2.weight", "1.self_attn.out_proj.weight", "1.fc1.weight"):
        rep.close("d " + n, gp[n].grad, p[n].grad, 8e-2)
    assert all(q.grad is None for q in lm.model.parameters()), "the Llama itself must stay frozen"
    rep.finish()
