"""optim.FusedAdamW (csrc/optim.cu) against torch.optim.AdamW -- the optimizer the reference builds
(language_modelling/run_generation.py:329-333): same update, same state layout, bf16 shadows kept current."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _params(seed, dev):
    g = torch.Generator().manual_seed(seed)
    shapes = [(256, 384), (1000, 7), (513,), (1,), (64, 64, 3)]      # odd sizes: vector body + scalar tail, 1-D (no shadow), 3-D
    return [torch.nn.Parameter(torch.randn(*s, generator=g).to(dev)) for s in shapes]


def test_fused_adamw_matches_torch_adamw():
    from mmgl_b200 import ops
    from mmgl_b200.optim import FusedAdamW
    dev = "cuda"
    ours, ref = _params(0, dev), _params(0, dev)
    kw = dict(lr=3e-3, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.05)
    o1, o2 = FusedAdamW(ours, **kw), torch.optim.AdamW(ref, **kw)
    g = torch.Generator().manual_seed(1)
    for step in range(6):
        for a, b in zip(ours, ref):
            gr = torch.randn(a.shape, generator=g).to(dev) * (10.0 if step == 3 else 1.0)
            a.grad, b.grad = gr.clone(), gr.clone()
        v0 = [p._version for p in ours]
        o2.step(); o1.step()      # (torch's step first: the global post-step hook drops every trainable shadow, ours included)
        for p, v in zip(ours, v0):
            assert p._version > v, "the version counter must move: caches keyed on it would go stale"
        for a, b in zip(ours, ref):
            assert torch.allclose(a, b, rtol=2e-6, atol=2e-7), (step, a.shape, (a - b).abs().max())
    for a, b in zip(ours, ref):
        sa, sb = o1.state[a], o2.state[b]
        assert torch.allclose(sa["exp_avg"], sb["exp_avg"], rtol=1e-5, atol=1e-6)
        assert torch.allclose(sa["exp_avg_sq"], sb["exp_avg_sq"], rtol=1e-5, atol=1e-7)
        assert int(sa["step"]) == int(sb["step"]) == 6
        if a.dim() >= 2:       # the shadow IS what the kernels will read next step, and it equals a fresh conversion
            assert ops.w16(a).data_ptr() == sa["_shadow"].data_ptr()
            assert torch.equal(ops.w16(a), a.detach().to(torch.bfloat16))
    sd = o1.state_dict()
    assert all("_shadow" not in st for st in sd["state"].values())
    assert set(sd["state"][0].keys()) == set(o2.state_dict()["state"][0].keys())
    o3 = torch.optim.AdamW(_params(0, dev), **kw)
    o3.load_state_dict(sd)      # checkpoints are interchangeable with torch's AdamW


def test_fused_adamw_grad_scale_and_rejects():
    from mmgl_b200.optim import FusedAdamW
    dev = "cuda"
    a, b = _params(2, dev)[:1], _params(2, dev)[:1]
    o1, o2 = FusedAdamW(a, lr=1e-2), torch.optim.AdamW(b, lr=1e-2)
    gr = torch.randn_like(a[0])
    a[0].grad, b[0].grad = gr.clone(), gr * 0.25
    o1.step(grad_scale=0.25); o2.step()
    assert torch.allclose(a[0], b[0], rtol=2e-6, atol=2e-7)
    with pytest.raises(NotImplementedError):
        FusedAdamW(a, amsgrad=True)


@pytest.mark.parametrize("make", [lambda ps: torch.optim.AdamW(ps, lr=1e-2, fused=True),
                                  lambda ps: torch.optim.AdamW(ps, lr=1e-2),
                                  lambda ps: torch.optim.SGD(ps, lr=1e-1)], ids=["adamw_fused", "adamw_foreach", "sgd"])
def test_forward_sees_the_updated_weights_under_torch_optimizers(make):
    """torch.optim.AdamW(fused=True) updates parameters WITHOUT bumping their version counter; the bf16 shadow cache must not
    serve the old weights afterwards (global optimizer post-step hook in ops.py)."""
    from mmgl_b200 import ops
    w = torch.nn.Parameter(torch.randn(64, 32, device="cuda"))
    x = torch.randn(8, 32, device="cuda").to(torch.bfloat16)
    opt = make([w])
    y0 = ops.linear(x, w)
    y0.float().square().mean().backward()
    opt.step()
    opt.zero_grad(set_to_none=True)
    assert torch.equal(ops.w16(w), w.detach().to(torch.bfloat16)), "stale bf16 shadow after optimizer.step()"
    y1 = ops.linear(x, w)
    ref = (x.float() @ w.detach().to(torch.bfloat16).float().t())
    assert (y1.float() - ref).abs().max() <= 2e-2 * ref.abs().max()
    assert not torch.equal(y0, y1)
