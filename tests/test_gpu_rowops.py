"""HBM-bound row kernels (LayerNorm, column / scalar reductions, neighbor-bank packing, GCN aggregate) against the
CPU oracle / plain torch fp32 on the same bf16-rounded inputs.

Tolerances: bf16 outputs 4e-3 rel-L2 (one output rounding), fp32 reductions 1e-5; the byte mask and the
placement of bank rows are bit-exact (integer / index work).
"""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

from oracle import mmgl_oracle as O
from util import BF16, assert_close, randn

pytestmark = pytest.mark.gpu

TOL_BF16 = 4e-3
TOL_F32 = 1e-5


def _K():
    from mmgl_b200 import _capi
    return _capi


# ------------------------------------------------------------------------------------------------ LayerNorm
@pytest.mark.parametrize("rows,hidden", [(5, 64), (640, 768), (1280, 2048), (33, 4096), (9, 1000)])
def test_layernorm_fwd_bwd(rows, hidden):
    gen = torch.Generator().manual_seed(rows + hidden)
    x = (randn(gen, rows, hidden) * 2 + 0.5).to(BF16)
    gamma = randn(gen, hidden) * 0.5 + 1
    beta = randn(gen, hidden) * 0.1
    dy = randn(gen, rows, hidden).to(BF16)
    d_res = randn(gen, rows, hidden).to(BF16)
    K = _K()
    y = torch.empty_like(x)
    mean = torch.empty(rows, dtype=torch.float32, device="cuda")
    rstd = torch.empty_like(mean)
    K.layernorm_fwd(x, gamma, beta, y, mean, rstd, 1e-5)

    xr = x.float().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    y_ref = F.layer_norm(xr, (hidden,), gr, br, 1e-5)
    assert_close("y", y, y_ref, TOL_BF16)
    assert_close("mean", mean, xr.mean(-1), 1e-4)
    assert_close("rstd", rstd, (xr.var(-1, unbiased=False) + 1e-5).rsqrt(), 1e-4)

    y_ref.backward(dy.float())
    dx = torch.empty_like(x)
    dg = torch.empty(hidden, dtype=torch.float32, device="cuda")
    db = torch.empty_like(dg)
    K.layernorm_bwd(dy, x, gamma, mean, rstd, d_res, dx, dg, db)
    assert_close("dx(+res)", dx, xr.grad + d_res.float(), TOL_BF16)
    assert_close("dgamma", dg, gr.grad, 1e-4)
    assert_close("dbeta", db, br.grad, 1e-4)
    # frozen affine: dx only, no residual
    dx2 = torch.empty_like(x)
    K.layernorm_bwd(dy, x, gamma, mean, rstd, None, dx2)
    assert_close("dx", dx2, xr.grad, TOL_BF16)


# ------------------------------------------------------------------------------------------------ reductions
@pytest.mark.parametrize("m,n", [(1, 8), (640, 2048), (2563, 8192), (77, 100)])
def test_colsum_and_gate_grad(m, n):
    gen = torch.Generator().manual_seed(m + n)
    K = _K()
    ld = (n + 7) // 8 * 8
    buf = randn(gen, m, ld).to(BF16)
    x = buf[:, :n]
    gate = torch.tensor([-0.3], device="cuda")
    out = torch.empty(n, dtype=torch.float32, device="cuda")
    K.colsum(x, out, scale=0.5, gate=gate)
    assert_close("colsum", out, 0.5 * torch.tanh(gate) * x.float().sum(0), TOL_F32)
    base = out.clone()
    K.colsum(x, out, accumulate=True)
    assert_close("colsum +=", out, base + x.float().sum(0), TOL_F32)
    if n % 8 == 0:
        a = randn(gen, m, n).to(BF16)
        g = torch.empty(1, dtype=torch.float32, device="cuda")
        K.gate_grad(x, a, gate, g)
        want = (1 - torch.tanh(gate) ** 2) * (x.float() * a.float()).sum()
        assert abs(float(g) - float(want)) <= 1e-4 * max(1.0, float((x.float() * a.float()).abs().sum()) ** 0.5)


# ------------------------------------------------------------------------------------------------ dropout
@pytest.mark.parametrize("m,n,p", [(64, 256, 0.1), (130, 2048, 0.1), (33, 100, 0.5), (5, 8, 0.0)])
def test_dropout_mask_is_bit_exact(m, n, p):
    """The counter-based keep mask of the kernels == its integer restatement in the oracle (bit-exact), both from
    the standalone kernel and from the GEMM epilogue."""
    gen = torch.Generator().manual_seed(m + n)
    K = _K()
    seed = 0x1234_5678_9ABC_DEF0 + m
    ld = (n + 7) // 8 * 8
    x = (randn(gen, m, ld).abs() + 0.5).to(BF16)[:, :n]
    out = torch.empty((m, ld), dtype=BF16, device="cuda")[:, :n]
    K.dropout_apply(x, out, p, seed)
    mult = O.dropout_multiplier(seed, p, m, n)
    assert torch.equal(out.float().cpu() != 0, mult != 0), "keep mask differs from the oracle restatement"
    assert_close("values", out, x.float().cpu() * mult, TOL_BF16)
    if p > 0:
        frac = float((mult == 0).float().mean())
        assert abs(frac - p) < 4 * (p * (1 - p) / (m * n)) ** 0.5 + 1e-3, f"drop fraction {frac} vs p {p}"
    # GEMM epilogue: identity weight so the product is x itself
    if n % 8 == 0:
        eye = torch.eye(n, device="cuda").to(BF16)
        y = torch.empty((m, n), dtype=BF16, device="cuda")
        K.gemm(x.contiguous(), eye, y, dropout_p=p, dropout_seed=seed)
        assert torch.equal(y.float().cpu() != 0, mult != 0), "GEMM epilogue mask differs"
        assert_close("gemm values", y, x.float().cpu() * mult, TOL_BF16)


# ------------------------------------------------------------------------------------------------ neighbor bank
def _bank_inputs(gen, b, t, i, n_tok, h, with_lpe):
    """WikiWeb2M-shaped ragged bank: per sample a random interleave of valid text/image neighbors first, pads last
    (wikiweb2m/data.py:349-454); pos ids 1..n for valid neighbors, 0 for padding."""
    n = t + i
    row = n_tok * h
    text_proj = randn(gen, b, t, row).to(BF16)
    image_proj = randn(gen, b, i, row).to(BF16)
    text_tab = randn(gen, t + 2, row).to(BF16)
    image_tab = randn(gen, i + 2, row).to(BF16)
    tpos = torch.zeros(b, t, dtype=torch.int64)
    ipos = torch.zeros(b, i, dtype=torch.int64)
    tloc = torch.zeros(b, t, dtype=torch.int64)
    iloc = torch.zeros(b, i, dtype=torch.int64)
    for s in range(b):
        nt = int(torch.randint(1, t + 1, (1,), generator=gen))
        ni = int(torch.randint(0, i + 1, (1,), generator=gen))
        tpos[s, :nt] = torch.arange(1, nt + 1)
        ipos[s, :ni] = torch.arange(1, ni + 1)
        order = torch.randperm(nt + ni, generator=gen)  # slots of the valid neighbors, interleaved
        rest = torch.arange(nt + ni, n)
        slots_valid_t, slots_valid_i = order[:nt], order[nt:]
        tloc[s] = torch.cat((slots_valid_t, rest[: t - nt]))
        iloc[s] = torch.cat((slots_valid_i, rest[t - nt:]))
    args = dict(text_proj=text_proj, text_tab=text_tab, tpos=tpos.cuda(), tloc=tloc.cuda(),
                image_proj=image_proj, image_tab=image_tab, ipos=ipos.cuda(), iloc=iloc.cuda())
    if with_lpe:
        k = n - 4
        args.update(lpe=randn(gen, b, n + 1, k), lpe_w=randn(gen, row, k, scale=0.2).to(BF16), lpe_b=randn(gen, row))
    return args


def _bank_oracle(a, n_tok, h):
    b, t, _ = a["text_proj"].shape
    i = a["image_proj"].shape[1]
    te = (a["text_proj"].float().cpu() + F.embedding(a["tpos"].cpu(), a["text_tab"].float().cpu())).reshape(b, t, n_tok, h)
    ve = (a["image_proj"].float().cpu() + F.embedding(a["ipos"].cpu(), a["image_tab"].float().cpu())).reshape(b, i, n_tok, h)
    bank, mask = O.pack_bank(te, a["tpos"].cpu(), a["tloc"].cpu(), ve, a["ipos"].cpu(), a["iloc"].cpu())
    if "lpe" in a:
        bank = O.lpe_add(bank, a["lpe"].cpu(), a["lpe_w"].float().cpu(), a["lpe_b"].cpu(), n_tok)
    return bank, mask


@pytest.mark.parametrize("b,t,i,n_tok,h,with_lpe", [
    (2, 3, 2, 2, 64, False), (4, 11, 5, 4, 2048, False), (3, 22, 10, 4, 768, True), (2, 11, 5, 4, 2048, True)])
def test_bank_pack_forward_backward(b, t, i, n_tok, h, with_lpe):
    from mmgl_b200 import ops
    gen = torch.Generator().manual_seed(b * 100 + t + i + h)
    a = _bank_inputs(gen, b, t, i, n_tok, h, with_lpe)
    leaves = {k: a[k].clone().requires_grad_(True) for k in ("text_proj", "text_tab", "image_proj", "image_tab")}
    lpe_w = a["lpe_w"].clone().requires_grad_(True) if with_lpe else None
    lpe_b = a["lpe_b"].clone().requires_grad_(True) if with_lpe else None
    bank, mask = ops.bank_pack(leaves["text_proj"], leaves["text_tab"], a["tpos"], a["tloc"],
                               leaves["image_proj"], leaves["image_tab"], a["ipos"], a["iloc"],
                               lpe=a.get("lpe"), lpe_weight=lpe_w, lpe_bias=lpe_b, n_tok=n_tok)
    # oracle (fp32, autograd)
    ao = {k: (v.float().cpu() if v.is_floating_point() else v.cpu()) for k, v in a.items()}
    for k in ("text_proj", "text_tab", "image_proj", "image_tab", "lpe_w", "lpe_b"):
        if k in ao:
            ao[k] = ao[k].requires_grad_(True)
    bank_ref, mask_ref = _bank_oracle(ao, n_tok, h)
    assert torch.equal(mask.cpu().bool(), mask_ref), "byte mask must be bit-exact"
    assert_close("bank", bank, bank_ref, TOL_BF16)
    if not with_lpe:  # pure placement + one add: padding slots are exact zeros or exact copies
        unused = ~mask_ref
        # rows never named by any location stay zero; here every slot is named, so compare placement exactly
        assert torch.equal(bank.float().cpu()[unused], bank_ref.detach().to(BF16).float()[unused])

    w = randn(gen, *bank.shape).to(BF16)
    bank.backward(w)
    (bank_ref * w.float().cpu()).sum().backward()
    assert_close("d text_proj", leaves["text_proj"].grad, ao["text_proj"].grad, TOL_BF16)
    assert_close("d image_proj", leaves["image_proj"].grad, ao["image_proj"].grad, TOL_BF16)
    assert_close("d text_tab", leaves["text_tab"].grad, ao["text_tab"].grad, TOL_BF16)
    assert_close("d image_tab", leaves["image_tab"].grad, ao["image_tab"].grad, TOL_BF16)
    if with_lpe:
        assert_close("d lpe_w", lpe_w.grad, ao["lpe_w"].grad, TOL_BF16)
        assert_close("d lpe_b", lpe_b.grad, ao["lpe_b"].grad, 1e-4)


# ------------------------------------------------------------------------------------------------ GCN helpers
@pytest.mark.parametrize("b,n,dim", [(3, 6, 48), (2, 16, 8192), (2, 32, 768)])
def test_gcn_concat_and_combine(b, n, dim):
    gen = torch.Generator().manual_seed(n + dim)
    K = _K()
    nodes = n + 1
    x = randn(gen, b, n, dim).to(BF16)
    adj = (torch.rand(b, nodes, nodes, generator=gen) > 0.5).float() + torch.eye(nodes)
    adj = (adj / adj.sum(-1, keepdim=True)).cuda()
    out = torch.empty((b * nodes, 2 * dim), dtype=BF16, device="cuda")
    K.gcn_concat_fwd(x, adj, out, b, nodes, dim, True)
    xr = torch.cat((torch.zeros(b, 1, dim, device="cuda"), x.float()), 1)
    want = torch.cat((xr, torch.bmm(adj, xr)), -1).reshape(b * nodes, 2 * dim)
    assert_close("concat", out, want, TOL_BF16)
    assert torch.equal(out[:, :dim].float(), xr.reshape(-1, dim)), "identity half must be an exact copy"

    dc = randn(gen, b * nodes, 2 * dim).to(BF16)
    relu_src = randn(gen, b * nodes, dim).to(BF16)
    dx = torch.empty((b * nodes, dim), dtype=BF16, device="cuda")
    K.gcn_combine_bwd(dc, adj, relu_src, dx, b, nodes, dim, False)
    d3 = dc.float().reshape(b, nodes, 2 * dim)
    want = (d3[..., :dim] + torch.bmm(adj.transpose(1, 2), d3[..., dim:])) * (relu_src.float().reshape(b, nodes, dim) > 0)
    assert_close("combine", dx.reshape(b, nodes, dim), want, TOL_BF16)
    dx2 = torch.empty((b, n, dim), dtype=BF16, device="cuda")
    K.gcn_combine_bwd(dc, adj, None, dx2, b, nodes, dim, True)
    want2 = (d3[..., :dim] + torch.bmm(adj.transpose(1, 2), d3[..., dim:]))[:, 1:]
    assert_close("combine drop root", dx2, want2, TOL_BF16)


# ------------------------------------------------------------------------------------------------ cross-entropy
@pytest.mark.parametrize("b,s,v", [(2, 20, 512), (3, 17, 50272), (1, 5, 1001)])
def test_fused_cross_entropy_matches_torch(b, s, v):
    """mmgl_ce_fwd/bwd vs nn.CrossEntropyLoss on the same bf16 logits in fp32 (model/modelling_cross_attention.py:
    828-836: shifted, mean over non-ignored positions).  loss 1e-5 relative, dlogits 4e-3 rel-L2 (bf16 output)."""
    from mmgl_b200 import ops
    gen = torch.Generator().manual_seed(b + s + v)
    ld = (v + 7) // 8 * 8
    logits = (randn(gen, b, s, ld) * 3).to(BF16)[..., :v].requires_grad_(True)
    labels = torch.randint(0, v, (b, s), generator=gen).cuda()
    labels[0, 3] = -100
    loss = ops.shifted_cross_entropy(logits, labels)
    loss.backward()
    ref_in = logits.detach().float().requires_grad_(True)
    ref = F.cross_entropy(ref_in[:, :-1].reshape(-1, v), labels[:, 1:].reshape(-1), ignore_index=-100)
    ref.backward()
    assert abs(float(loss) - float(ref)) <= 1e-5 * abs(float(ref)) + 1e-6, (float(loss), float(ref))
    assert_close("dlogits", logits.grad, ref_in.grad, TOL_BF16)
    assert float(logits.grad[:, -1].abs().max()) == 0.0, "last position must not receive gradient"


@pytest.mark.parametrize("rows,hidden", [(70, 768), (33, 128), (16, 2048)])
def test_rmsnorm_fwd_bwd(rows, hidden):
    """T5LayerNorm (HF models/t5/modeling_t5.py:46-70): y = w * x * rsqrt(mean(x^2) + eps), no mean, no shift.
    bf16 in/out, fp32 statistics: y 4e-3, dx 1e-2, dw 1e-2 rel-L2."""
    from mmgl_b200 import ops
    gen = torch.Generator().manual_seed(rows + hidden)
    x = (randn(gen, rows, hidden) * 2 + 0.5).to(BF16)
    w = (1 + 0.2 * randn(gen, hidden)).requires_grad_(True)
    dy = randn(gen, rows, hidden).to(BF16)
    xg = x.clone().requires_grad_(True)
    y = ops.rms_norm(xg, w, 1e-6)
    y.backward(dy)
    xr = x.float().cpu().requires_grad_(True)
    wr = w.detach().float().cpu().requires_grad_(True)
    yr = wr * (xr * torch.rsqrt(xr.pow(2).mean(-1, keepdim=True) + 1e-6))
    yr.backward(dy.float().cpu())
    assert_close("y", y, yr, 4e-3)
    assert_close("dx", xg.grad, xr.grad, 1e-2)
    assert_close("dw", w.grad, wr.grad, 1e-2)


@pytest.mark.parametrize("rows,hidden,kind", [(4741, 2048, "layer"), (4800, 768, "rms"), (4737, 4096, "layer"),
                                              (9000, 128, "rms")])
def test_staged_norm_backward_at_large_row_counts(rows, hidden, kind):
    """Above 4736 rows the norm backward runs as persistent CTAs fed by a cp.async.bulk shared-memory ring (ragged last
    tile, 2-6 stages by row size; hidden 4096 falls back to the register-resident kernel): dx must equal torch fp32 to
    the bf16 tolerances of the small-shape tests, with and without a residual gradient; the forward at these sizes too."""
    from mmgl_b200 import _capi as K
    gen = torch.Generator().manual_seed(rows + hidden)
    x = (randn(gen, rows, hidden) * 2 + 0.5).to(BF16)
    g = 1 + 0.2 * randn(gen, hidden)
    b = 0.1 * randn(gen, hidden)
    dy = randn(gen, rows, hidden).to(BF16)
    res = randn(gen, rows, hidden).to(BF16)
    y = torch.empty_like(x)
    mean = torch.empty(rows, device="cuda") if kind == "layer" else None
    rstd = torch.empty(rows, device="cuda")
    xr = x.float().requires_grad_(True)
    if kind == "layer":
        K.layernorm_fwd(x, g, b, y, mean, rstd, 1e-5)
        yr = F.layer_norm(xr, (hidden,), g, b, 1e-5)
        assert_close("mean", mean, xr.detach().mean(-1), 1e-3)
    else:
        K.rmsnorm_fwd(x, g, y, rstd, 1e-6)
        yr = g * (xr * torch.rsqrt(xr.pow(2).mean(-1, keepdim=True) + 1e-6))
    assert_close("y", y, yr, 4e-3)
    yr.backward(dy.float())
    for r in (None, res):
        dx = torch.full_like(x, float("nan"))
        if kind == "layer":
            K.layernorm_bwd(dy, x, g, mean, rstd, r, dx)
        else:
            K.rmsnorm_bwd(dy, x, g, rstd, r, dx)
        want = xr.grad if r is None else xr.grad + r.float()
        assert_close("dx" + ("" if r is None else " + d_res"), dx, want, 1e-2)
        assert bool(torch.isfinite(dx.float()).all())


@pytest.mark.parametrize("kind", ["layer", "rms"])
def test_norm_fork_fuses_the_residual_gradient(kind):
    """(norm(x), x) fork of a pre-norm residual block (modelling_cross_attention.py:318-337 pattern: h = x + f(LN(x))):
    the gradient of x must equal autograd's sum of the norm branch and the residual branch."""
    from mmgl_b200 import ops
    gen = torch.Generator().manual_seed(9)
    rows, hidden = 150, 768
    x = (randn(gen, 2, rows // 2, hidden) * 1.5 + 0.3).to(BF16)
    w = (1 + 0.2 * randn(gen, hidden)).requires_grad_(True)
    b = (0.1 * randn(gen, hidden)).requires_grad_(True)
    d1 = randn(gen, 2, rows // 2, hidden).to(BF16)
    d2 = randn(gen, 2, rows // 2, hidden).to(BF16)
    xg = x.clone().requires_grad_(True)
    if kind == "layer":
        y, r = ops.layer_norm_fork(xg, w, b, 1e-5)
    else:
        y, r = ops.rms_norm_fork(xg, w, 1e-6)
    assert r.data_ptr() == xg.data_ptr()
    torch.autograd.backward([y, r], [d1, d2])
    xr = x.float().cpu().requires_grad_(True)
    wr = w.detach().float().cpu().requires_grad_(True)
    if kind == "layer":
        yr = F.layer_norm(xr, (hidden,), wr, b.detach().float().cpu(), 1e-5)
    else:
        yr = wr * (xr * torch.rsqrt(xr.pow(2).mean(-1, keepdim=True) + 1e-6))
    torch.autograd.backward([yr, xr * 1.0], [d1.float().cpu(), d2.float().cpu()])
    assert_close("y", y, yr, 4e-3)
    assert_close("dx", xg.grad, xr.grad, 1e-2)
    assert_close("dw", w.grad, wr.grad, 1e-2)
    # residual branch unused: only the norm gradient
    xg2 = x.clone().requires_grad_(True)
    y2, _ = ops.layer_norm_fork(xg2, w, b, 1e-5) if kind == "layer" else ops.rms_norm_fork(xg2, w, 1e-6)
    y2.backward(d1)
    xr2 = x.float().cpu().requires_grad_(True)
    if kind == "layer":
        F.layer_norm(xr2, (hidden,), wr.detach(), b.detach().float().cpu(), 1e-5).backward(d1.float().cpu())
    else:
        (wr.detach() * (xr2 * torch.rsqrt(xr2.pow(2).mean(-1, keepdim=True) + 1e-6))).backward(d1.float().cpu())
    assert_close("dx (no residual)", xg2.grad, xr2.grad, 1e-2)


def test_mlp_with_hidden_dropout_matches_reference_semantics():
    """T5DenseActDense + T5LayerFF: y = r + drop2(wo(drop1(relu(wi(x))))) with both masks from the counter-based RNG
    (restated in oracle.dropout_multiplier); both dropouts and the ReLU live in GEMM epilogues."""
    from mmgl_b200 import ops
    gen = torch.Generator().manual_seed(9)
    m, h, f, p = 96, 128, 256, 0.1
    x = randn(gen, m, h).to(BF16)
    w1 = (randn(gen, f, h) * 0.1).to(BF16)
    w2 = (randn(gen, h, f) * 0.1).to(BF16)
    dy = randn(gen, m, h).to(BF16)
    seed_h, seed_o = ops.peek_dropout_seeds(2)
    xg, w1g, w2g = (t.clone().requires_grad_(True) for t in (x, w1, w2))
    y = ops.mlp(xg, w1g, None, w2g, None, residual=xg, dropout_p=p, hidden_dropout_p=p)
    y.backward(dy)
    xr, w1r, w2r = (t.float().cpu().requires_grad_(True) for t in (x, w1, w2))
    hid = torch.relu(xr @ w1r.t()).to(BF16).float() * O.dropout_multiplier(seed_h, p, m, f)
    yr = xr + (hid.to(BF16).float() @ w2r.t()) * O.dropout_multiplier(seed_o, p, m, h)
    yr.backward(dy.float().cpu())
    assert_close("y", y, yr, 5e-3)
    assert_close("dx", xg.grad, xr.grad, 1.5e-2)
    assert_close("dw1", w1g.grad, w1r.grad, 1.5e-2)
    assert_close("dw2", w2g.grad, w2r.grad, 1.5e-2)
