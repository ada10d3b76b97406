"""Self-attention kernels (mmgl_attn_fwd / mmgl_attn_bwd) against the CPU
oracle's mpt_attention core: the reference's MPTAttention self branch with the additive causal + padding mask
(model/modelling_cross_attention.py:201-275, :455-476), fp32 on the bf16-rounded inputs.

Tolerances (rel-L2): O 4e-3 (bf16 output + bf16 P operand), gradients 6e-3 (measured on B200: O 2.2e-3, dQ / dK / dV 2.4e-3 .. 2.8e-3; profiles/r02_parity_measured.txt).  Rows that are padding QUERIES are compared
too (the reference's loss covers them)."""
import pytest
import torch

from oracle import mmgl_oracle as O
from util import BF16, Report, randn

pytestmark = pytest.mark.gpu


def _oracle(qkv, key_mask, heads, causal, scale, d_o=None):
    b, s, h3 = qkv.shape
    h = h3 // 3
    d = h // heads
    x = qkv.float().cpu().requires_grad_(True)
    q, k, v = x[..., :h], x[..., h:2 * h], x[..., 2 * h:]
    qh, kh, vh = (t.reshape(b, s, heads, d).transpose(1, 2) for t in (q, k, v))
    w = (qh * scale) @ kh.transpose(-1, -2)
    add = torch.zeros(b, 1, s, s)
    if key_mask is not None:
        add = add + O.expand_mask(key_mask.cpu(), torch.float32, s)
    if causal:
        add = add + O.causal_mask(b, s, torch.float32)
    w = torch.max(w + add, torch.tensor(torch.finfo(torch.float32).min))
    o = (torch.softmax(w, -1) @ vh).transpose(1, 2).reshape(b, s, h)
    g = None
    if d_o is not None:
        (g,) = torch.autograd.grad(o, x, d_o.float().cpu())
    return o, g


@pytest.mark.parametrize("b,s,heads,d,causal,pad", [
    (2, 128, 2, 64, True, "none"),
    (2, 200, 2, 64, True, "segments"),     # seq tail + padding in the middle and at the end (input / output segments)
    (1, 640, 4, 64, True, "segments"),     # cfg2 sequence length
    (2, 300, 2, 64, False, "right"),       # bidirectional with key padding (encoder-style)
    (2, 130, 1, 128, True, "segments"),    # head_dim 128
    (1, 96, 2, 64, True, "none"),          # shorter than one block
    (6, 640, 32, 64, True, "segments"),    # cfg2 layer shape, 192 (sample, head) items > 148 SMs: CTAs walk several items
    (3, 1152, 32, 128, True, "segments"),  # cfg5 layer shape (head_dim 128, 9 query tiles)
    (5, 300, 40, 64, False, "right"),      # 200 items, bidirectional, ragged tail tile
])
def test_self_attention_forward_backward(b, s, heads, d, causal, pad):
    from mmgl_b200 import ops
    gen = torch.Generator().manual_seed(b * 100 + s + d)
    h = heads * d
    qkv = randn(gen, b, s, 3 * h).to(BF16)
    key_mask = torch.ones(b, s, dtype=torch.bool)
    if pad == "right":
        key_mask[0, int(s * 0.7):] = False
    elif pad == "segments":
        cut = int(s * 0.8)
        key_mask[:, int(cut * 0.6):cut] = False      # right padding of the input segment
        key_mask[0, int(s * 0.93):] = False          # right padding of the output segment
    d_o = randn(gen, b, s, h).to(BF16)
    x = qkv.clone().requires_grad_(True)
    o = ops.self_attention(x, key_mask.cuda() if pad != "none" else None, heads, causal=causal, scale=d ** -0.5)
    o.backward(d_o)
    o_ref, g_ref = _oracle(qkv, key_mask if pad != "none" else None, heads, causal, d ** -0.5, d_o)
    rep = Report()
    rep.close("O", o, o_ref, 4e-3)
    rep.close("dQ", x.grad[..., :h], g_ref[..., :h], 6e-3)
    rep.close("dK", x.grad[..., h:2 * h], g_ref[..., h:2 * h], 6e-3)
    rep.close("dV", x.grad[..., 2 * h:], g_ref[..., 2 * h:], 6e-3)
    rep.finish()


def test_backward_is_run_to_run_deterministic():
    """dQ is summed over key blocks in a per-CTA scratch in a fixed order (csrc/sattn_bwd_sm100.cu): bit-identical reruns."""
    from mmgl_b200 import ops
    gen = torch.Generator().manual_seed(11)
    b, s, heads, d = 5, 640, 32, 64
    h = heads * d
    qkv = randn(gen, b, s, 3 * h).to(BF16)
    d_o = randn(gen, b, s, h).to(BF16)
    grads = []
    for _ in range(3):
        x = qkv.clone().requires_grad_(True)
        ops.self_attention(x, None, heads, causal=True, scale=d ** -0.5).backward(d_o)
        grads.append(x.grad.clone())
    assert torch.equal(grads[0], grads[1]) and torch.equal(grads[0], grads[2])


def test_padding_keys_have_zero_influence():
    from mmgl_b200 import ops
    gen = torch.Generator().manual_seed(3)
    b, s, heads, d = 2, 256, 2, 64
    h = heads * d
    qkv = randn(gen, b, s, 3 * h).to(BF16)
    key_mask = torch.ones(b, s, dtype=torch.bool)
    key_mask[:, 100:140] = False
    o1 = ops.self_attention(qkv, key_mask.cuda(), heads)
    qkv2 = qkv.clone()
    qkv2[:, 100:140, h:] = 55.0       # K and V of padding keys
    o2 = ops.self_attention(qkv2, key_mask.cuda(), heads)
    assert torch.equal(o1, o2)


# ------------------------------------------------------------------------------------------------ general form
def _rel_vec(gen, heads, sq, sk):
    """a relative-position bias given as its per-distance vector [heads, sq + sk - 1] and as the dense [1,nh,sq,sk]"""
    vec = (torch.randn(heads, sq + sk - 1, generator=gen) * 1.5).requires_grad_(True)
    idx = torch.arange(sk)[None, :] - torch.arange(sq)[:, None] + sq - 1
    return vec, vec[:, idx][None]


@pytest.mark.parametrize("b,sq,sk,heads,d,causal,bias,pad", [
    (2, 128, 576, 2, 64, False, False, True),    # T5 decoder cross-attention: 128 queries over the encoder output
    (2, 100, 333, 2, 64, False, False, True),
    (1, 300, 300, 2, 64, True, True, False),     # T5 decoder self-attention: causal + relative bias
    (2, 576, 576, 2, 64, False, True, True),     # T5 encoder self-attention: bidirectional bias + key padding
    (2, 140, 140, 3, 64, True, True, True),
    (1, 640, 640, 2, 64, True, False, True),     # 5 query tiles: pairs (4,3) (2,1) (0,-)
    (1, 384, 384, 2, 64, False, False, False),   # 3 query tiles, all blocks interior
    (1, 200, 130, 1, 128, False, True, True),    # head_dim 128
])
def test_general_attention_forward_backward(b, sq, sk, heads, d, causal, bias, pad):
    from mmgl_b200 import ops
    gen = torch.Generator().manual_seed(sq * 7 + sk + d)
    h = heads * d
    scale = 1.0 if bias else d ** -0.5
    mag = 0.35 if bias else 1.0           # T5 scores are unscaled: keep q.k in a softmax-friendly range
    q, k, v = (randn(gen, b, n, h, scale=mag).to(BF16) for n in (sq, sk, sk))
    d_o = randn(gen, b, sq, h).to(BF16)
    key_mask = None
    if pad:
        key_mask = torch.ones(b, sk, dtype=torch.bool)
        key_mask[0, int(sk * 0.55):int(sk * 0.7)] = False
        key_mask[-1, int(sk * 0.9):] = False
    vec = dense = dvec = None
    if bias:
        vec, dense = _rel_vec(gen, heads, sq, sk)
        dvec = vec.detach().cuda().requires_grad_(True)      # a trainable bias (T5 with peft "none")
    xs = [t.clone().requires_grad_(True) for t in (q, k, v)]
    o = ops.attention(*xs, key_mask=None if key_mask is None else key_mask.cuda(),
                      rel_bias=dvec, heads=heads, causal=causal, scale=scale)
    o.backward(d_o)
    rs = [t.float().cpu().requires_grad_(True) for t in (q, k, v)]
    o_ref = O.attention_core(*rs, heads, scale, key_mask, causal, dense)
    o_ref.backward(d_o.float().cpu())
    rep = Report()
    rep.close("O", o, o_ref, 4e-3)
    for name, x, r in zip(("dQ", "dK", "dV"), xs, rs):
        rep.close(name, x.grad, r.grad, 6e-3)
    if bias:
        rep.close("d rel_bias", dvec.grad, vec.grad, 6e-3)
    rep.finish()


@pytest.mark.parametrize("b,sq,sk,heads,causal,p", [(2, 200, 200, 2, True, 0.1), (1, 128, 300, 2, False, 0.25)])
def test_attention_probability_dropout(b, sq, sk, heads, causal, p):
    """Dropout on the probabilities (HF T5Attention): the kernel's counter-based keep mask is restated in integers by
    oracle.dropout_multiplier over the [B*nh*Sq, Sk] probability matrix; given that mask the result must equal the
    reference semantics softmax(.) * keep / (1 - p) @ V, forward and backward."""
    from mmgl_b200 import ops
    gen = torch.Generator().manual_seed(11)
    d = 64
    h = heads * d
    q, k, v = (randn(gen, b, n, h).to(BF16) for n in (sq, sk, sk))
    d_o = randn(gen, b, sq, h).to(BF16)
    key_mask = torch.ones(b, sk, dtype=torch.bool)
    key_mask[0, sk - 37:] = False
    seed = ops.peek_dropout_seeds(1)[0]
    xs = [t.clone().requires_grad_(True) for t in (q, k, v)]
    o = ops.attention(*xs, key_mask=key_mask.cuda(), heads=heads, causal=causal, scale=d ** -0.5, dropout_p=p)
    o.backward(d_o)
    mult = O.dropout_multiplier(seed, p, b * heads * sq, sk).reshape(b, heads, sq, sk)
    frac = float((mult == 0).float().mean())
    assert abs(frac - p) < 0.01, frac
    rs = [t.float().cpu().requires_grad_(True) for t in (q, k, v)]
    o_ref = O.attention_core(*rs, heads, d ** -0.5, key_mask, causal, None, mult)
    o_ref.backward(d_o.float().cpu())
    rep = Report()
    rep.close("O", o, o_ref, 5e-3)
    for name, x, r in zip(("dQ", "dK", "dV"), xs, rs):
        rep.close(name, x.grad, r.grad, 7e-3)
    rep.finish()


def test_rows_without_any_attended_key_are_uniform():
    """A sample whose keys are ALL padding: the reference's clamp makes every score finfo.min -> uniform attention
    (non-causal: over all keys, exactly the reference)."""
    from mmgl_b200 import ops
    gen = torch.Generator().manual_seed(5)
    b, sq, sk, heads, d = 2, 130, 200, 2, 64
    h = heads * d
    q, k, v = (randn(gen, b, n, h).to(BF16) for n in (sq, sk, sk))
    key_mask = torch.ones(b, sk, dtype=torch.bool)
    key_mask[1] = False
    o = ops.attention(q, k, v, key_mask=key_mask.cuda(), heads=heads, causal=False, scale=d ** -0.5)
    o_ref = O.attention_core(q.float().cpu(), k.float().cpu(), v.float().cpu(), heads, d ** -0.5, key_mask, False)
    rep = Report()
    rep.close("O", o, o_ref, 4e-3)
    rep.finish()


@pytest.mark.parametrize("b,sq,prefix,heads,pad", [(2, 200, 20, 2, True), (1, 640, 20, 2, True), (2, 128, 128, 1, False),
                                                   (1, 300, 7, 2, False)])
def test_causal_attention_over_prefix_keys(b, sq, prefix, heads, pad):
    """Causal attention whose keys are `prefix` virtual tokens followed by the queries' own positions (peft prefix
    tuning, model/modelling_self_attention.py:88-92): key j is visible to query i iff j <= i + prefix."""
    from mmgl_b200 import ops
    gen = torch.Generator().manual_seed(sq + prefix)
    d = 64
    h = heads * d
    sk = sq + prefix
    q, k, v = (randn(gen, b, n, h).to(BF16) for n in (sq, sk, sk))
    d_o = randn(gen, b, sq, h).to(BF16)
    key_mask = torch.ones(b, sk, dtype=torch.bool)
    if pad:
        key_mask[0, prefix + int(sq * 0.5):prefix + int(sq * 0.7)] = False
    xs = [t.clone().requires_grad_(True) for t in (q, k, v)]
    o = ops.attention(*xs, key_mask=key_mask.cuda() if pad else None, heads=heads, causal=True, scale=d ** -0.5)
    o.backward(d_o)
    # dense additive mask of the bottom-right aligned causal pattern
    allowed = (torch.arange(sk)[None, :] <= torch.arange(sq)[:, None] + prefix)
    bias = torch.zeros(sq, sk).masked_fill(~allowed, torch.finfo(torch.float32).min)[None, None]
    rs = [t.float().cpu().requires_grad_(True) for t in (q, k, v)]
    o_ref = O.attention_core(*rs, heads, d ** -0.5, key_mask if pad else None, False, bias)
    o_ref.backward(d_o.float().cpu())
    rep = Report()
    rep.close("O", o, o_ref, 4e-3)
    for name, x, r in zip(("dQ", "dK", "dV"), xs, rs):
        rep.close(name, x.grad, r.grad, 6e-3)
    rep.finish()
