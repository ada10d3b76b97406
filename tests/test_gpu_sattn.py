"""Causal / key-padded self-attention kernels (mmgl_sattn_fwd / mmgl_sattn_bwd) against the CPU
oracle's mpt_attention core: the reference's MPTAttention self branch with the additive causal + padding mask
(model/modelling_cross_attention.py:201-275, :455-476), fp32 on the bf16-rounded inputs.

Tolerances (rel-L2): O 4e-3 (bf16 output + bf16 P operand), gradients 1e-2.  Rows that are padding QUERIES are compared
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
    rep.close("dQ", x.grad[..., :h], g_ref[..., :h], 1e-2)
    rep.close("dK", x.grad[..., h:2 * h], g_ref[..., h:2 * h], 1e-2)
    rep.close("dV", x.grad[..., 2 * h:], g_ref[..., 2 * h:], 1e-2)
    rep.finish()


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
