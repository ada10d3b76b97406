"""Frozen neighbor encoders on the package's kernels against the HF modules' own forward on the
same weights (fp32 HF forward on the GPU, bf16-rounded weights).  Tolerance 1.5e-2 rel-L2 on the consumed outputs
([CLS] hidden state / pooler_output): two to four post-/pre-LN transformer layers in bf16."""
import pytest
import torch

from util import BF16, Report

pytestmark = pytest.mark.gpu


def _round_weights(model):
    with torch.no_grad():
        for p in model.parameters():
            if p.dim() >= 2:
                p.copy_(p.to(BF16).float())


@pytest.mark.parametrize("layers,seq", [(2, 40), (3, 130)])
def test_roberta_cls_matches_hf(layers, seq):
    from transformers import RobertaConfig, RobertaModel
    from mmgl_b200 import encoders
    torch.manual_seed(0)
    cfg = RobertaConfig(vocab_size=300, hidden_size=128, num_hidden_layers=layers, num_attention_heads=2,
                        intermediate_size=256, max_position_embeddings=seq + 4, pad_token_id=1, layer_norm_eps=1e-5,
                        hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    model = RobertaModel(cfg).cuda().eval()
    _round_weights(model)
    gen = torch.Generator().manual_seed(1)
    n = 5
    ids = torch.randint(4, 300, (n, seq), generator=gen)
    am = torch.ones(n, seq, dtype=torch.long)
    for r, ln in enumerate((seq, seq - 7, 9, seq // 2, 1)):
        am[r, ln:] = 0
        ids[r, ln:] = 1
    ids, am = ids.cuda(), am.cuda()
    with torch.no_grad():
        ref = model(input_ids=ids, attention_mask=am).last_hidden_state[:, 0]
    got = encoders.roberta_cls_hidden(model, ids, am)
    rep = Report()
    rep.close("cls hidden", got, ref, 1.5e-2)
    rep.finish()


def test_clip_pooler_matches_hf():
    from transformers import CLIPVisionConfig, CLIPVisionModel
    from mmgl_b200 import encoders
    torch.manual_seed(0)
    cfg = CLIPVisionConfig(hidden_size=128, intermediate_size=256, num_hidden_layers=3, num_attention_heads=2,
                           image_size=64, patch_size=16, hidden_act="quick_gelu", attention_dropout=0.0)
    model = CLIPVisionModel(cfg).cuda().eval()
    _round_weights(model)
    gen = torch.Generator().manual_seed(2)
    px = torch.randn(4, 3, 64, 64, generator=gen).to(BF16).float().cuda()
    with torch.no_grad():
        ref = model(px).pooler_output
    got = encoders.clip_pooler_output(model, px)
    rep = Report()
    rep.close("pooler_output", got, ref, 1.5e-2)
    rep.finish()


def test_clip_vit_b16_shape_runs():
    """ViT-B/16 geometry (224 px, 197 tokens = 2 key blocks with a ragged tail)."""
    from transformers import CLIPVisionModel
    from mmgl_b200 import configs, encoders
    torch.manual_seed(0)
    cfg = configs.visual_config("clip-vit-base-patch16")
    cfg.num_hidden_layers = 2
    model = CLIPVisionModel(cfg).cuda().eval()
    _round_weights(model)
    px = torch.randn(3, 3, 224, 224, generator=torch.Generator().manual_seed(3)).to(BF16).float().cuda()
    with torch.no_grad():
        ref = model(px).pooler_output
    got = encoders.clip_pooler_output(model, px)
    rep = Report()
    rep.close("pooler_output", got, ref, 1.5e-2)
    rep.finish()


def test_roberta_packing_changes_nothing():
    """Dropping the padding positions (variable-length attention over the real tokens only) must give the same [CLS]
    hidden states as running on the padded batch: padded keys are masked and padded queries are never consumed."""
    from transformers import RobertaConfig, RobertaModel
    from mmgl_b200 import encoders
    torch.manual_seed(0)
    cfg = RobertaConfig(vocab_size=512, hidden_size=128, num_hidden_layers=3, num_attention_heads=2, intermediate_size=256,
                        max_position_embeddings=400, pad_token_id=1)
    model = RobertaModel(cfg).cuda().eval()
    gen = torch.Generator().manual_seed(1)
    n, s = 9, 300
    lens = torch.tensor([300, 17, 128, 129, 1, 255, 64, 200, 33])
    am = (torch.arange(s)[None, :] < lens[:, None]).long()
    ids = torch.where(am.bool(), torch.randint(4, 512, (n, s), generator=gen), torch.tensor(1))
    a = encoders.roberta_cls_hidden(model, ids.cuda(), am.cuda(), pack_padding=True)
    b = encoders.roberta_cls_hidden(model, ids.cuda(), am.cuda(), pack_padding=False)
    ref = model(input_ids=ids.cuda(), attention_mask=am.cuda()).last_hidden_state[:, 0]
    rep = Report()
    rep.close("packed vs padded", a, b, 2e-3)
    rep.close("packed vs HF fp32", a, ref, 1.5e-2)
    rep.finish()
    # a mask that is not prefix-form falls back to the padded path
    am2 = am.clone(); am2[2, 5] = 0
    c = encoders.roberta_cls_hidden(model, ids.cuda(), am2.cuda())
    assert torch.isfinite(c.float()).all()
