"""Host-side logic of the product path that needs no GPU: index / bucket arithmetic that the CUDA kernels are driven by.
Integer work: bit-exact.  (The kernels themselves are covered by the `-m gpu` tests.)"""
import pytest
import torch


def test_t5_bucket_matches_hf_relative_position_bucket():
    """lm._t5_bucket restates T5Attention._relative_position_bucket (HF models/t5/modeling_t5.py:189-233)."""
    from transformers.models.t5.modeling_t5 import T5Attention
    from mmgl_b200 import lm as L
    rel = torch.arange(-700, 701)
    for bidirectional in (True, False):
        for nb, md in ((32, 128), (16, 64)):
            want = T5Attention._relative_position_bucket(rel, bidirectional=bidirectional, num_buckets=nb, max_distance=md)
            got = L._t5_bucket(rel, bidirectional, nb, md)
            assert torch.equal(got, want), (bidirectional, nb, md)


@pytest.mark.parametrize("decoder", [False, True])
def test_t5_rel_bias_vector_is_the_dense_bias_by_distance(decoder):
    """t5_rel_bias returns [heads, sq + sk - 1] with entry d + sq - 1 = bias at distance d = key - query; expanding it
    must give T5Attention.compute_bias (:236-251) exactly, and it stays differentiable for a trainable table."""
    from transformers import T5Config
    from transformers.models.t5.modeling_t5 import T5Attention
    from mmgl_b200 import lm as L
    torch.manual_seed(0)
    cfg = T5Config(d_model=64, d_kv=16, num_heads=4, is_decoder=decoder)
    cfg.is_decoder = decoder
    attn = T5Attention(cfg, has_relative_attention_bias=True, layer_idx=0)
    with torch.no_grad():
        attn.relative_attention_bias.weight.normal_()
    sq, sk = 37, 37
    dense = attn.compute_bias(sq, sk)[0]                      # [heads, sq, sk]
    vec = L.t5_rel_bias(attn, sq, sk)
    assert vec.shape == (4, sq + sk - 1) and vec.requires_grad
    idx = torch.arange(sk)[None, :] - torch.arange(sq)[:, None] + sq - 1
    assert torch.equal(vec[:, idx], dense)
    vec.sum().backward()
    assert attn.relative_attention_bias.weight.grad is not None
    attn.relative_attention_bias.weight.requires_grad_(False)
    assert not L.t5_rel_bias(attn, sq, sk).requires_grad


def test_pack_plan_drops_right_padding_only():
    """encoders._pack_plan (frozen RoBERTa on real tokens only; data.py:457 right-pads every neighbor): prefix-form masks
    give the flat indices of the real tokens and their cumulative lengths; anything else (left padding, holes, an empty
    row, < 10% padding) must refuse, so the dense path runs."""
    from mmgl_b200.encoders import _pack_plan
    lens = torch.tensor([5, 12, 1, 9])
    s = 12
    am = (torch.arange(s)[None, :] < lens[:, None]).long()
    idx, cu, total, longest = _pack_plan(am)
    assert total == int(lens.sum()) and longest == 12
    assert cu.dtype == torch.int32 and cu.tolist() == [0, 5, 17, 18, 27]
    assert torch.equal(idx, torch.nonzero(am.reshape(-1)).squeeze(1))
    hole = am.clone()
    hole[1, 3] = 0
    assert _pack_plan(hole) is None
    left = am.flip(1)
    assert _pack_plan(left) is None
    empty = am.clone()
    empty[2] = 0
    assert _pack_plan(empty) is None
    assert _pack_plan(torch.ones(4, s, dtype=torch.long)) is None          # nothing to save


def test_lm_output_is_indexable_like_hf_outputs():
    """run_generation.py reads outputs.loss / outputs[1] (:466-474)."""
    from mmgl_b200.lm import LMOutput
    loss, logits = torch.tensor(1.5), torch.zeros(2, 3)
    out = LMOutput(loss=loss, logits=logits)
    assert out.loss is loss and out["logits"] is logits and out[0] is loss and out[1] is logits
    assert LMOutput(logits=logits)[0] is logits
