"""Host-side plan of the ragged neighbor work of a batch (optional; removes the per-step host syncs).

The neighbor encoders run on the REAL tokens of the VALID neighbors only: which neighbors are valid
(``pos_id > 0``) and how long each text is decides tensor shapes and kernel grids, so the module needs those numbers on
the host.  Without a plan it reads them back from the device tensors it was handed (two small device->host reads per
step: ``_needed`` and ``encoders._pack_plan``), which drains the launch queue at the start of every step and de-phases
DDP ranks.  The data pipeline knows all of it before the batch ever leaves the host: ``attach_plan(batch)`` (call it in the
``collate_fn`` / right after ``WikiWeb2M.__getitem__`` batches are stacked, wikiweb2m/data.py:458-468) adds one extra entry,
``neighbor_plan``, that travels with the batch -- it has ``.cuda()`` / ``.to()`` / ``.pin_memory()`` like a tensor, so the
reference loop's ``{k: v.cuda(gpu, non_blocking=True) for k, v in batch.items()}`` (run_generation.py:464) keeps working --
and ``CrossAttentionModel.forward`` / ``SelfAttentionModel.forward`` accept it as a keyword.  Same results, no sync.
"""
from __future__ import annotations

import torch


class NeighborPlan:
    """index tensors (device-movable) + the python ints that size the kernels"""

    _TENSORS = ("text_idx", "tok_idx", "cu", "image_idx")

    def __init__(self, text_idx, tok_idx, cu, image_idx, total, max_len, pack):
        self.text_idx, self.tok_idx, self.cu, self.image_idx = text_idx, tok_idx, cu, image_idx
        self.total, self.max_len, self.pack = int(total), int(max_len), bool(pack)

    def _map(self, fn):
        return NeighborPlan(*(None if getattr(self, n) is None else fn(getattr(self, n)) for n in self._TENSORS),
                            self.total, self.max_len, self.pack)

    def to(self, *a, **kw):
        return self._map(lambda t: t.to(*a, **kw))

    def cuda(self, device=None, non_blocking=False):
        return self._map(lambda t: t.cuda(device, non_blocking=non_blocking))

    def pin_memory(self):
        return self._map(lambda t: t.pin_memory())

    def contiguous(self):
        return self

    def record_stream(self, stream):
        for n in self._TENSORS:
            t = getattr(self, n)
            if t is not None and t.is_cuda:
                t.record_stream(stream)

    def numel(self):
        return sum(getattr(self, n).numel() for n in self._TENSORS if getattr(self, n) is not None)

    def element_size(self):
        return 8


def _needed(pos_ids):
    """valid neighbors, plus every neighbor of a sample that has none (reference parity: such a sample attends uniformly
    over its masked rows) -- the rule of modules._NeighborEncoderMixin._needed"""
    valid = pos_ids > 0
    need = valid | ~valid.any(dim=1, keepdim=True)
    return None if bool(need.all()) else need.reshape(-1).nonzero(as_tuple=False).squeeze(1)


def make_plan(batch, skip_padding_neighbors=True, pack_padding=True) -> NeighborPlan:
    """batch: the HOST batch dict (wikiweb2m/data.py:458-468 keys).  Pure CPU integer work."""
    pos = batch["neighbor_pos_ids"]
    text_idx = _needed(pos) if skip_padding_neighbors else None
    am = batch["neighbor_attention_mask"]
    am2 = am.reshape(-1, am.shape[-1]) != 0
    if text_idx is not None:
        am2 = am2.index_select(0, text_idx)
    n, s = am2.shape
    lens = am2.sum(1)
    last = (am2 * torch.arange(1, s + 1)).amax(1)
    total, max_len = int(lens.sum()), int(lens.max()) if n else 0
    pack = bool(pack_padding and ((lens == last) & (lens > 0)).all() and total * 10 <= n * s * 9)
    tok_idx = cu = None
    if pack:
        tok_idx = am2.reshape(-1).nonzero(as_tuple=False).squeeze(1)
        cu = torch.zeros(n + 1, dtype=torch.int32)
        cu[1:] = torch.cumsum(lens, 0)
    image_idx = None
    if skip_padding_neighbors and batch.get("neighbor_images_pos_ids") is not None and batch["neighbor_images_pos_ids"].numel():
        image_idx = _needed(batch["neighbor_images_pos_ids"])
    if text_idx is None:
        text_idx = torch.arange(pos.numel())
    if image_idx is None and batch.get("neighbor_images_pos_ids") is not None:
        image_idx = torch.arange(batch["neighbor_images_pos_ids"].numel())
    return NeighborPlan(text_idx, tok_idx, cu, image_idx, total, max_len, pack)


def attach_plan(batch, **kw):
    out = dict(batch)
    out["neighbor_plan"] = make_plan(batch, **kw)
    return out
