"""Autograd operators of mmgl_b200.ops (forward AND backward through the C ABI) against the CPU oracle.

The oracle (oracle/mmgl_oracle.py, pinned to the reference by tests/test_oracle_golden.py) runs in fp32 with
autograd on the bf16-rounded parameters and inputs.  Tolerances (rel-L2 per tensor, bf16 compute / fp32 accumulate):
  block output 5e-3; gradients 1.5e-2 when the FFN's ReLU is kept away from its kink ("smooth" cases: fc1 bias
  shifted so every pre-activation is positive -- this checks all the algebra tightly); in the natural cases the
  LayerNorm output feeding fc1 is stored in bf16, which flips the ReLU mask of the ~0.3% of pre-activations that
  are within 2^-9 of zero -- a discontinuity, so every gradient downstream of the FFN carries a rel-L2 error of
  about sqrt(flipped fraction): tolerance 7e-2 for fc1 / final_layer_norm grads and 4e-2 for the rest.  (The ReLU
  mask logic itself is checked exactly in test_linear_and_mlp, where both sides see identical pre-activations.)
"""
import contextlib

import pytest
import torch
import torch.nn.functional as F

from oracle import mmgl_oracle as O
from util import BF16, Report, assert_close, randn

pytestmark = pytest.mark.gpu

TOL_Y = 5e-3
TOL_G = 1.5e-2        # smooth cases / operators without a kink
TOL_G_RELU = 4e-2     # natural cases, downstream of bf16-induced ReLU mask flips
TOL_G_FFN = 7e-2      # fc1 / final_layer_norm gradients in the natural cases
# test_gated_cross_layer_at_benchmarked_sizes (H = 2048 / 4096): MEASURED on B200 (profiles/r02_parity_measured.txt) and set to
# about 2x the measurement.  fp32 oracle: y 2.6e-3; attention-side gradients 1.6e-2; fc1 / final_layer_norm 3.9e-2.
# Oracle with the kernels' bf16 storage points emulated (oracle.bf16_storage): y 1.9e-3; 1.1e-2; 2.5e-2 -- what remains is
# the bf16 rounding of the backward intermediates (dO, dS, P: 2^-9 each), not branch flips.
TOL_BIG = dict(fp32=(5e-3, 3.2e-2, 8e-2), storage=(4e-3, 2.2e-2, 5e-2))


def _leaf(t, dtype=None):
    t = t.detach().clone()
    if dtype is not None:
        t = t.to(dtype)
    return t.requires_grad_(True)


def _cpu32(t):
    return t.detach().float().cpu().requires_grad_(True)


@pytest.mark.parametrize("param_dtype", [torch.float32, BF16])
def test_linear_and_mlp(param_dtype):
    from mmgl_b200 import ops
    gen = torch.Generator().manual_seed(1)
    b, s, h, f = 2, 50, 128, 384
    x = randn(gen, b, s, h).to(BF16)
    res = randn(gen, b, s, h).to(BF16)
    w1, b1 = randn(gen, f, h, scale=0.1).to(BF16), randn(gen, f, scale=0.1)
    w2, b2 = randn(gen, h, f, scale=0.1).to(BF16), randn(gen, h, scale=0.1)
    dy = randn(gen, b, s, h).to(BF16)

    xg, rg = _leaf(x), _leaf(res)
    pw1, pb1, pw2, pb2 = (_leaf(t, param_dtype) for t in (w1, b1, w2, b2))
    y = ops.mlp(xg, pw1, pb1, pw2, pb2, residual=rg)
    y.backward(dy)

    xc, rc = _cpu32(x), _cpu32(res)
    cw1, cb1, cw2, cb2 = (_cpu32(t.to(param_dtype)) for t in (w1, b1, w2, b2))
    yr = rc + F.linear(F.relu(F.linear(xc, cw1, cb1)), cw2, cb2)
    yr.backward(dy.float().cpu())
    assert_close("y", y, yr, TOL_Y)
    assert_close("dx", xg.grad, xc.grad, TOL_G)
    assert_close("dres", rg.grad, rc.grad, 1e-6)
    for name, p, c in (("w1", pw1, cw1), ("b1", pb1, cb1), ("w2", pw2, cw2), ("b2", pb2, cb2)):
        assert p.grad.dtype == param_dtype, f"{name}: gradient dtype {p.grad.dtype} != parameter dtype"
        assert_close("d" + name, p.grad, c.grad, TOL_G)

    # plain linear with alpha
    xg2, pw = _leaf(x), _leaf(w1, param_dtype)
    y2 = ops.linear(xg2, pw, pb1.detach().requires_grad_(True), alpha=0.25)
    y2.backward(torch.ones_like(y2))
    xc2, cw = _cpu32(x), _cpu32(w1.to(param_dtype))
    yr2 = 0.25 * F.linear(xc2, cw, cb1.detach())
    yr2.sum().backward()
    assert_close("linear y", y2, yr2, TOL_Y)
    assert_close("linear dx", xg2.grad, xc2.grad, TOL_G)
    assert_close("linear dw", pw.grad, cw.grad, TOL_G)


def test_lora_linear():
    """x W^T + b + (alpha/r)(x A^T) B^T -- oracle lora_linear (peft restated; parity unpinned, see oracle header)."""
    from mmgl_b200 import ops
    gen = torch.Generator().manual_seed(2)
    m, kin, nout, r, alpha = 150, 768, 768, 64, 1.0
    x = randn(gen, 3, m // 3, kin).to(BF16)
    w, bias = randn(gen, nout, kin, scale=0.05).to(BF16), randn(gen, nout, scale=0.1)
    a, bb = randn(gen, r, kin, scale=0.05).to(BF16), randn(gen, nout, r, scale=0.05).to(BF16)
    dy = randn(gen, 3, m // 3, nout).to(BF16)
    xg, pw, pb, pa, pbb = _leaf(x), _leaf(w), _leaf(bias), _leaf(a), _leaf(bb)
    y = ops.lora_linear(xg, pw, pb, pa, pbb, alpha / r)
    y.backward(dy)
    xc, cw, cb, ca, cbb = (_cpu32(t) for t in (x, w, bias, a, bb))
    yr = O.lora_linear(xc, cw, cb, ca, cbb, alpha, r)
    yr.backward(dy.float().cpu())
    assert_close("y", y, yr, TOL_Y)
    assert_close("dx", xg.grad, xc.grad, TOL_G)
    assert_close("dW", pw.grad, cw.grad, TOL_G)
    assert_close("dA", pa.grad, ca.grad, TOL_G)
    assert_close("dB", pbb.grad, cbb.grad, TOL_G)
    assert_close("dbias", pb.grad, cb.grad, TOL_G)
    # invariant I5: B == 0 -> exactly the base linear
    y0 = ops.lora_linear(x, w, bias, a, torch.zeros_like(bb), alpha / r)
    y1 = ops.linear(x, w, bias)
    assert torch.equal(y0, y1)


def _layer_params(gen, h, f, scale=0.08):
    p = {}
    for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
        p[f"self_attn.{n}.weight"] = randn(gen, h, h, scale=scale)
        p[f"self_attn.{n}.bias"] = randn(gen, h, scale=scale)
    p["fc1.weight"], p["fc1.bias"] = randn(gen, f, h, scale=scale), randn(gen, f, scale=scale)
    p["fc2.weight"], p["fc2.bias"] = randn(gen, h, f, scale=scale), randn(gen, h, scale=scale)
    for n in ("self_attn_layer_norm", "final_layer_norm"):
        p[n + ".weight"] = 1 + randn(gen, h, scale=0.1)
        p[n + ".bias"] = randn(gen, h, scale=0.1)
    p["gating1"] = torch.tensor([0.7], device="cuda")
    p["gating2"] = torch.tensor([-0.4], device="cuda")
    return p


def _round_big(p):
    """bf16-round the matrices (what the kernels consume); small fp32 params stay fp32."""
    return {k: (v.to(BF16).float() if v.dim() == 2 else v) for k, v in p.items()}


@pytest.mark.parametrize("smooth", [True, False])
@pytest.mark.parametrize("pre_ln", [True, False])
@pytest.mark.parametrize("b,s,nk,heads,h,f", [(2, 40, 24, 2, 128, 256), (1, 130, 64, 4, 256, 1024)])
def test_gated_cross_layer_vs_oracle(pre_ln, b, s, nk, heads, h, f, smooth):
    from mmgl_b200 import ops
    gen = torch.Generator().manual_seed(17 + int(pre_ln) + h)
    p = _round_big(_layer_params(gen, h, f))
    if smooth:
        p["fc1.bias"] = p["fc1.bias"] + 5.0
    x = randn(gen, b, s, h).to(BF16)
    bank = randn(gen, b, nk, h).to(BF16)
    mask = torch.rand(b, nk, generator=gen) > 0.3
    mask[0, :] = True
    dy = randn(gen, b, s, h).to(BF16)

    gp = {k: _leaf(v) for k, v in p.items()}
    xg, bg = _leaf(x), _leaf(bank)
    y = ops.gated_cross_layer(
        xg, bg, mask.cuda(), gp["self_attn_layer_norm.weight"], gp["self_attn_layer_norm.bias"],
        gp["self_attn.q_proj.weight"], gp["self_attn.q_proj.bias"], gp["self_attn.k_proj.weight"],
        gp["self_attn.k_proj.bias"], gp["self_attn.v_proj.weight"], gp["self_attn.v_proj.bias"],
        gp["self_attn.out_proj.weight"], gp["self_attn.out_proj.bias"], gp["gating1"],
        gp["final_layer_norm.weight"], gp["final_layer_norm.bias"], gp["fc1.weight"], gp["fc1.bias"],
        gp["fc2.weight"], gp["fc2.bias"], gp["gating2"], heads, 1e-5, pre_ln)
    y.backward(dy)

    cp = {k: _cpu32(v) for k, v in p.items()}
    xc, bc = _cpu32(x), _cpu32(bank)
    yr = O.mpt_decoder_layer(xc, cp, heads, cross_attention=True, bank=bc,
                             bank_add_mask=O.expand_mask(mask, torch.float32, s), do_layer_norm_before=pre_ln)
    yr.backward(dy.float().cpu())
    rep = Report()
    tol = TOL_G if smooth else TOL_G_RELU
    rep.close("y", y, yr, TOL_Y)
    rep.close("dx", xg.grad, xc.grad, tol)
    rep.close("dbank", bg.grad, bc.grad, tol)
    _compare_param_grads(rep, gp, {k: v.grad for k, v in cp.items()}, tol, smooth)
    rep.finish()


@pytest.mark.parametrize("b,s,nk,heads,h,f", [
    (2, 640, 64, 32, 2048, 8192),      # cfg2: OPT-1.3B width, seq 512+128, 16 neighbors x 4 tokens (the benchmarked shape)
    (2, 640, 128, 32, 2048, 8192),     # cfg4: 32 neighbors
    (1, 1152, 128, 32, 4096, 11008),   # cfg5: Llama-2-7B width, head_dim 128, seq 1024+128
])
def test_gated_cross_layer_at_benchmarked_sizes(b, s, nk, heads, h, f):
    """the gated block against the oracle at the dims bench.py times (pair-kernel GEMM tiles, 5 / 9
    query tiles per head, Nk = 64 / 128), natural ReLU (no smoothing), ragged masks, init_std-scaled weights."""
    from mmgl_b200 import ops
    gen = torch.Generator().manual_seed(97 + nk + h)
    p = _round_big(_layer_params(gen, h, f, scale=0.02))
    x = randn(gen, b, s, h).to(BF16)
    bank = randn(gen, b, nk, h).to(BF16)
    mask = torch.rand(b, nk, generator=gen) > 0.3
    mask[0, :] = True
    dy = randn(gen, b, s, h).to(BF16)
    gp = {k: _leaf(v) for k, v in p.items()}
    xg, bg = _leaf(x), _leaf(bank)
    y = ops.gated_cross_layer(
        xg, bg, mask.cuda(), gp["self_attn_layer_norm.weight"], gp["self_attn_layer_norm.bias"],
        gp["self_attn.q_proj.weight"], gp["self_attn.q_proj.bias"], gp["self_attn.k_proj.weight"],
        gp["self_attn.k_proj.bias"], gp["self_attn.v_proj.weight"], gp["self_attn.v_proj.bias"],
        gp["self_attn.out_proj.weight"], gp["self_attn.out_proj.bias"], gp["gating1"],
        gp["final_layer_norm.weight"], gp["final_layer_norm.bias"], gp["fc1.weight"], gp["fc1.bias"],
        gp["fc2.weight"], gp["fc2.bias"], gp["gating2"], heads, 1e-5, True)
    y.backward(dy)
    rep = Report()
    # (1) accuracy: the plain fp32 oracle; (2) algebra: the oracle with the kernels' bf16 storage points emulated, so both
    # sides take the same ReLU branches (oracle.bf16_storage) -- tight tolerance
    for label, ctx, (tol_y, tol_g, tol_ffn) in (("fp32", contextlib.nullcontext(), TOL_BIG["fp32"]),
                                                ("bf16-storage", O.bf16_storage(), TOL_BIG["storage"])):
        cp = {k: _cpu32(v) for k, v in p.items()}
        xc, bc = _cpu32(x), _cpu32(bank)
        with ctx:
            yr = O.mpt_decoder_layer(xc, cp, heads, cross_attention=True, bank=bc,
                                     bank_add_mask=O.expand_mask(mask, torch.float32, s), do_layer_norm_before=True)
        yr.backward(dy.float().cpu())
        rep.close(f"[{label}] y", y, yr, tol_y)
        rep.close(f"[{label}] dx", xg.grad, xc.grad, tol_g)
        rep.close(f"[{label}] dbank", bg.grad, bc.grad, tol_g)
        _compare_param_grads(rep, gp, {k: v.grad for k, v in cp.items()}, tol_g, False, label=f"[{label}] ", tol_ffn=tol_ffn)
    rep.finish()


@pytest.mark.parametrize("pre_ln", [True, False])
def test_gated_cross_layer_with_dropout(pre_ln):
    """Training mode (dropout p = 0.1 after out_proj and after fc2, model/modelling_cross_attention.py:332, :356):
    the oracle applies the SAME keep masks (counter-based RNG restated in oracle.dropout_multiplier)."""
    from mmgl_b200 import ops
    b, s, nk, heads, h, f, pdrop = 2, 72, 32, 2, 128, 512, 0.1
    gen = torch.Generator().manual_seed(31 + int(pre_ln))
    p = _round_big(_layer_params(gen, h, f))
    x, bank = randn(gen, b, s, h).to(BF16), randn(gen, b, nk, h).to(BF16)
    mask = torch.rand(b, nk, generator=gen) > 0.3
    mask[:, 0] = True
    dy = randn(gen, b, s, h).to(BF16)
    gp = {k: _leaf(v) for k, v in p.items()}
    xg, bg = _leaf(x), _leaf(bank)
    seed1, seed2 = ops.peek_dropout_seeds(2)
    y = ops.gated_cross_layer(
        xg, bg, mask.cuda(), gp["self_attn_layer_norm.weight"], gp["self_attn_layer_norm.bias"],
        gp["self_attn.q_proj.weight"], gp["self_attn.q_proj.bias"], gp["self_attn.k_proj.weight"],
        gp["self_attn.k_proj.bias"], gp["self_attn.v_proj.weight"], gp["self_attn.v_proj.bias"],
        gp["self_attn.out_proj.weight"], gp["self_attn.out_proj.bias"], gp["gating1"],
        gp["final_layer_norm.weight"], gp["final_layer_norm.bias"], gp["fc1.weight"], gp["fc1.bias"],
        gp["fc2.weight"], gp["fc2.bias"], gp["gating2"], heads, 1e-5, pre_ln, pdrop)
    y.backward(dy)
    cp = {k: _cpu32(v) for k, v in p.items()}
    xc, bc = _cpu32(x), _cpu32(bank)
    yr = O.mpt_decoder_layer(xc, cp, heads, cross_attention=True, bank=bc,
                             bank_add_mask=O.expand_mask(mask, torch.float32, s), do_layer_norm_before=pre_ln,
                             drop1=O.dropout_multiplier(seed1, pdrop, b * s, h),
                             drop2=O.dropout_multiplier(seed2, pdrop, b * s, h))
    yr.backward(dy.float().cpu())
    rep = Report()
    rep.close("y", y, yr, TOL_Y)
    rep.close("dx", xg.grad, xc.grad, TOL_G_RELU)
    rep.close("dbank", bg.grad, bc.grad, TOL_G_RELU)
    _compare_param_grads(rep, gp, {k: v.grad for k, v in cp.items()}, TOL_G_RELU, False)
    rep.finish()


def _compare_param_grads(rep, gp, ref_grads, tol, smooth, label="", tol_ffn=TOL_G_FFN):
    """k_proj.bias has an analytically ZERO gradient (a constant shift of all scores leaves the softmax unchanged), so
    it is compared absolutely against the scale of the v_proj.bias gradient; a scalar gate gradient is a sum of ~1e4
    signed bf16 products driven by a RANDOM cotangent (it nearly cancels), so its bf16 rounding noise is set by the
    size of the terms, not of the sum: rtol 3e-2 + atol 5e-2 * max|d out_proj.bias| (same terms, summed per column)."""
    scale = float(ref_grads["self_attn.v_proj.bias"].abs().max())
    for k, gr in ref_grads.items():
        if k.startswith("gating"):
            rep.scalar(label + "d " + k, gp[k].grad, gr, 3e-2, 5e-2 * float(ref_grads["self_attn.out_proj.bias"].abs().max()) + 1e-3)
        elif k == "self_attn.k_proj.bias":
            rep.absolute(label + "d " + k, gp[k].grad, gr, 4e-3 * scale)
        elif not smooth and k.startswith(("fc1.", "final_layer_norm.")):
            rep.close(label + "d " + k, gp[k].grad, gr, tol_ffn)
        else:
            rep.close(label + "d " + k, gp[k].grad, gr, tol)


@pytest.mark.parametrize("name", ["xattn_layer_d64_preln", "xattn_layer_d64_postln"])
def test_gated_cross_layer_vs_reference_golden(golden, name):
    """Directly against outputs + gradients of the REAL reference MPTDecoderLayer (tests/golden/make_golden.py)."""
    from mmgl_b200 import ops
    g = golden(name)
    st = g["state"]
    heads, pre_ln = g["cfg"]["num_heads"], g["cfg"]["do_layer_norm_before"]
    gp = {k: _leaf(v.cuda()) for k, v in st.items()}           # fp32 master parameters, bf16 shadows inside
    xg, bg = _leaf(g["x"].cuda()), _leaf(g["bank"].cuda())
    y = ops.gated_cross_layer(
        xg, bg, g["mask"].cuda(), gp["self_attn_layer_norm.weight"], gp["self_attn_layer_norm.bias"],
        gp["self_attn.q_proj.weight"], gp["self_attn.q_proj.bias"], gp["self_attn.k_proj.weight"],
        gp["self_attn.k_proj.bias"], gp["self_attn.v_proj.weight"], gp["self_attn.v_proj.bias"],
        gp["self_attn.out_proj.weight"], gp["self_attn.out_proj.bias"], gp["gating1"],
        gp["final_layer_norm.weight"], gp["final_layer_norm.bias"], gp["fc1.weight"], gp["fc1.bias"],
        gp["fc2.weight"], gp["fc2.bias"], gp["gating2"], heads, 1e-5, pre_ln)
    (y.float() * g["w"].cuda()).sum().backward()
    # fp32 reference on bf16-representable inputs/weights vs bf16 kernels (natural case, see module docstring)
    rep = Report()
    rep.close("y", y, g["y"], TOL_Y)
    rep.close("dx", xg.grad, g["dx"], TOL_G_RELU)
    rep.close("dbank", bg.grad, g["dbank"], TOL_G_RELU)
    _compare_param_grads(rep, gp, g["grads"], TOL_G_RELU, False)
    rep.finish()


def test_gcn_vs_oracle():
    from mmgl_b200 import ops
    gen = torch.Generator().manual_seed(23)
    b, n, din, dh = 2, 16, 512, 96
    x = randn(gen, b, n, din).to(BF16)
    adj = (torch.rand(b, n + 1, n + 1, generator=gen) > 0.6).float() + torch.eye(n + 1)
    adj = adj / adj.sum(-1, keepdim=True)
    w1 = randn(gen, dh, 2 * din, scale=0.05).to(BF16)
    w2 = randn(gen, din, 2 * dh, scale=0.05).to(BF16)
    dy = randn(gen, b, n, din).to(BF16)
    xg, p1, p2 = _leaf(x), _leaf(w1), _leaf(w2)
    y = ops.gcn(xg, adj.cuda(), p1, p2)
    y.backward(dy)
    xc, c1, c2 = _cpu32(x), _cpu32(w1), _cpu32(w2)
    yr = O.gcn_forward(xc, adj, c1, c2)
    yr.backward(dy.float().cpu())
    assert_close("y", y, yr, 8e-3)   # two chained bf16 GEMMs + bf16 aggregates
    assert_close("dx", xg.grad, xc.grad, 2e-2)
    assert_close("dw1", p1.grad, c1.grad, 2e-2)
    assert_close("dw2", p2.grad, c2.grad, 2e-2)


def test_no_cpu_fallback():
    from mmgl_b200 import ops
    x = torch.zeros(2, 8, 64, dtype=BF16)
    w = torch.zeros(64, 64, dtype=BF16)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.linear(x, w)
