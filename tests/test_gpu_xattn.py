"""Fused cross-attention core (mmgl_xattn_fwd / mmgl_xattn_bwd) against the CPU oracle's xattn_core
(oracle/mmgl_oracle.py, pinned to the reference's MPTAttention, model/modelling_cross_attention.py:206-271).

The oracle runs in fp32 on the bf16-rounded inputs.  Tolerances (rel-L2 over the tensor): O 4e-3 (bf16 output
rounding + bf16 P in the second contraction), log-sum-exp 1e-5 abs-ish (fp32), gradients 1e-2 (bf16 P/dS operands).
"""
import pytest
import torch

from oracle import mmgl_oracle as O
from util import BF16, assert_close, randn

pytestmark = pytest.mark.gpu

TOL_O = 4e-3
TOL_GRAD = 6e-3   # measured on B200: 2.4e-3 .. 2.8e-3 (profiles/r02_parity_measured.txt)


def _case(seed, b, s, nk, heads, d, mask_kind="ragged", scale=1.0):
    gen = torch.Generator().manual_seed(seed)
    h = heads * d
    q = (randn(gen, b, s, h) * scale * d ** -0.5).to(BF16)
    k = randn(gen, b, nk, h).to(BF16)
    v = randn(gen, b, nk, h).to(BF16)
    if mask_kind == "full":
        mask = torch.ones(b, nk, dtype=torch.bool)
    elif mask_kind == "ragged":  # per-sample valid count, like the packed neighbor bank
        valid = torch.randint(1, nk + 1, (b,), generator=gen)
        mask = torch.arange(nk)[None, :] < valid[:, None]
    elif mask_kind == "one_empty":  # a sample whose neighbors are all padding (softmax over finfo.min -> uniform)
        mask = torch.rand(b, nk, generator=gen) > 0.4
        mask[0, :] = False
    else:
        raise ValueError(mask_kind)
    return q, k, v, mask.cuda()


def _run_kernel(q, k, v, mask, heads, d_o=None):
    from mmgl_b200 import ops
    qg, kg, vg = (t.clone().requires_grad_(True) for t in (q, k, v))
    o = ops.xattn_core(qg, kg, vg, mask, heads)
    grads = None
    if d_o is not None:
        grads = torch.autograd.grad(o, (qg, kg, vg), d_o)
    return o, grads


def _run_oracle(q, k, v, mask, heads, d_o=None):
    qc, kc, vc = (t.float().cpu().requires_grad_(True) for t in (q, k, v))
    o, lse = O.xattn_core(qc, kc, vc, mask.cpu(), heads)
    grads = None
    if d_o is not None:
        grads = torch.autograd.grad(o, (qc, kc, vc), d_o.float().cpu())
    return o, lse, grads


@pytest.mark.parametrize("b,s,nk,heads,d,mask_kind", [
    (2, 64, 64, 2, 64, "full"),
    (2, 100, 64, 3, 64, "ragged"),      # seq tail
    (3, 70, 40, 2, 64, "ragged"),       # Nk not a multiple of 16
    (2, 130, 128, 2, 64, "one_empty"),
    (2, 96, 128, 2, 128, "ragged"),     # Llama-style head_dim 128
    (2, 33, 9, 1, 128, "full"),
    (1, 640, 64, 32, 64, "ragged"),     # cfg2 shape (OPT-1.3B, <=16 neighbors x 4 tokens)
])
def test_xattn_forward_backward(b, s, nk, heads, d, mask_kind):
    q, k, v, mask = _case(b * 1000 + s + nk, b, s, nk, heads, d, mask_kind)
    gen = torch.Generator().manual_seed(99)
    d_o = randn(gen, b, s, heads * d).to(BF16)
    o, (dq, dk, dv) = _run_kernel(q, k, v, mask, heads, d_o)
    o_ref, _, (dq_r, dk_r, dv_r) = _run_oracle(q, k, v, mask, heads, d_o)
    assert_close("O", o, o_ref, TOL_O)
    assert_close("dQ", dq, dq_r, TOL_GRAD)
    assert_close("dK", dk, dk_r, TOL_GRAD)
    assert_close("dV", dv, dv_r, TOL_GRAD)


def test_xattn_forward_maximum_bank():
    """Nk = 256 (forward-only limit; training shapes stop at 128 = 32 neighbors x 4 tokens)."""
    q, k, v, mask = _case(11, 1, 64, 256, 2, 64, "ragged")
    o, _ = _run_kernel(q, k, v, mask, 2)
    o_ref, _, _ = _run_oracle(q, k, v, mask, 2)
    assert_close("O", o, o_ref, TOL_O)


def test_xattn_stats_are_logsumexp():
    from mmgl_b200 import _capi as K
    b, s, nk, heads, d = 2, 80, 48, 2, 64
    q, k, v, mask = _case(3, b, s, nk, heads, d, "ragged", scale=3.0)
    h = heads * d
    o = torch.empty((b * s, h), dtype=BF16, device="cuda")
    stats = torch.empty((b, heads, s, 2), dtype=torch.float32, device="cuda")
    K.xattn_fwd(q.reshape(-1, h), k.reshape(-1, h), v.reshape(-1, h), mask.to(torch.uint8), o, stats, b, s, nk, heads, d)
    _, lse, _ = _run_oracle(q, k, v, mask, heads)
    got = stats[..., 0] - torch.log(stats[..., 1])
    assert float((got.cpu() - lse).abs().max()) < 2e-4


def test_masked_neighbors_have_zero_influence():
    """Invariant I2: bank rows with mask == 0 never change O -- bit-exact."""
    b, s, nk, heads, d = 2, 64, 64, 2, 64
    q, k, v, mask = _case(4, b, s, nk, heads, d, "ragged")
    mask[:, 0] = True
    o1, _ = _run_kernel(q, k, v, mask, heads)
    k2, v2 = k.clone(), v.clone()
    k2[~mask] = 123.0
    v2[~mask] = -77.0
    o2, _ = _run_kernel(q, k2, v2, mask, heads)
    assert torch.equal(o1, o2)


def test_xattn_fused_kv_buffer_views():
    """K and V as the two halves of one [B*Nk, 2H] projection output (ldk = ldv = 2H)."""
    from mmgl_b200 import _capi as K
    b, s, nk, heads, d = 2, 64, 32, 2, 64
    h = heads * d
    q, k, v, mask = _case(5, b, s, nk, heads, d, "full")
    kv = torch.cat((k.reshape(-1, h), v.reshape(-1, h)), dim=1).contiguous()
    o = torch.empty((b * s, h), dtype=BF16, device="cuda")
    stats = torch.empty((b, heads, s, 2), dtype=torch.float32, device="cuda")
    K.xattn_fwd(q.reshape(-1, h), kv[:, :h], kv[:, h:], mask.to(torch.uint8), o, stats, b, s, nk, heads, d)
    o_ref, _, _ = _run_oracle(q, k, v, mask, heads)
    assert_close("O", o.reshape(b, s, h), o_ref, TOL_O)


def test_xattn_rejects_unsupported():
    from mmgl_b200 import _capi as K
    q = torch.zeros((64, 96), dtype=BF16, device="cuda")
    m = torch.ones((1, 8), dtype=torch.uint8, device="cuda")
    st = torch.empty((1, 2, 64, 2), dtype=torch.float32, device="cuda")
    with pytest.raises(RuntimeError, match="head_dim"):
        K.xattn_fwd(q, q[:8], q[:8], m, q.clone(), st, 1, 64, 8, 2, 48)
