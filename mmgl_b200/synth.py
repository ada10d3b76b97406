"""Seeded synthetic WikiWeb2M-shaped batches (the 600k-page dataset, its images and the tokenizers are not available
offline).  Shapes, dtypes and padding conventions mirror ``wikiweb2m/data.py:296-469`` (``get_embedding_item``):

  input_ids / attention_mask / labels   [B, S_in + S_out] int64 -- input segment and summary segment each right-padded
                                        with pad id 1 (:321-333); labels = input_ids for decoder-only models
  neighbor_input_ids / _attention_mask  [B, T, S_in] int64      -- every text neighbor padded to max_input_length (:457)
  neighbor_pos_ids                      [B, T] int64            -- 1..n_text for valid neighbors, 0 for padding
  neighbor_images                       [B, I, 3, 224, 224] f32 -- zeros for padding images (:451)
  neighbor_images_pos_ids               [B, I] int64
  text_locations / image_locations      [B, T] / [B, I] int64   -- together a permutation of 0..T+I-1: valid neighbors
                                        interleaved first, padding slots last (:349-454)
  lpe   [B, T+I+1, T+I-4] f32, graph [B, T+I+1, T+I+1] f32 (row-normalised, self loops)  -- optional

Tensors are created on the CPU (pinned on request): the host->device copy is part of the end-to-end measurement.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch


@dataclass
class BatchSpec:
    batch: int = 4
    max_input_length: int = 512
    max_output_length: int = 128
    text_neighbors: int = 11
    image_neighbors: int = 5
    vocab_size: int = 50272
    neighbor_vocab_size: int = 50265
    image_size: int = 224
    pad_token_id: int = 1
    with_lpe: bool = False
    with_graph: bool = False
    decoder_only: bool = True
    neighbor_length: int = 0      # padded length of every text neighbor; 0 = max_input_length (data.py:457).  RoBERTa's 514
                                  # learned positions cap it at 512 when the section itself is longer (cfg5: seq 1024)


def make_batch(spec: BatchSpec, seed: int, pin: bool = False) -> dict:
    g = torch.Generator().manual_seed(seed)
    b, s_in, s_out = spec.batch, spec.max_input_length, spec.max_output_length
    t, i = spec.text_neighbors, spec.image_neighbors
    n = t + i

    def rint(lo, hi, shape):
        return torch.randint(lo, hi, shape, generator=g)

    len_in = rint(min(64, s_in), s_in + 1, (b,))
    len_out = rint(min(8, s_out), s_out + 1, (b,))
    pos_in = torch.arange(s_in)[None, :]
    pos_out = torch.arange(s_out)[None, :]
    m_in = pos_in < len_in[:, None]
    m_out = pos_out < len_out[:, None]
    ids_in = torch.where(m_in, rint(4, spec.vocab_size, (b, s_in)), torch.tensor(spec.pad_token_id))
    ids_out = torch.where(m_out, rint(4, spec.vocab_size, (b, s_out)), torch.tensor(spec.pad_token_id))
    if spec.decoder_only:
        input_ids = torch.cat((ids_in, ids_out), 1)
        attention_mask = torch.cat((m_in, m_out), 1).long()
        labels = input_ids.clone()
    else:
        input_ids, attention_mask = ids_in, m_in.long()
        labels = torch.where(m_out, ids_out, torch.tensor(-100))

    n_text = rint(1, t + 1, (b,))
    n_img = rint(0, i + 1, (b,)) if i > 0 else torch.zeros(b, dtype=torch.long)
    tpos = torch.where(torch.arange(t)[None, :] < n_text[:, None], torch.arange(1, t + 1)[None, :], torch.tensor(0))
    ipos = torch.where(torch.arange(i)[None, :] < n_img[:, None], torch.arange(1, i + 1)[None, :], torch.tensor(0))
    s_nb = spec.neighbor_length or s_in
    nlen = rint(min(16, s_nb), s_nb + 1, (b, t))
    nmask = (torch.arange(s_nb)[None, None, :] < nlen[:, :, None]) & (tpos > 0)[:, :, None]
    nmask[:, :, 0] = True  # an empty-string neighbor still tokenises to <s> (data.py:444-457)
    nids = torch.where(nmask, rint(4, spec.neighbor_vocab_size, (b, t, s_nb)), torch.tensor(1))   # RoBERTa pad id
    images = torch.randn((b, i, 3, spec.image_size, spec.image_size), generator=g)
    images = images * (ipos > 0)[:, :, None, None, None]

    tloc = torch.zeros((b, t), dtype=torch.long)
    iloc = torch.zeros((b, i), dtype=torch.long)
    for r in range(b):
        nt, ni = int(n_text[r]), int(n_img[r])
        order = torch.randperm(nt + ni, generator=g)
        rest = torch.arange(nt + ni, n)
        tloc[r] = torch.cat((order[:nt], rest[: t - nt]))
        iloc[r] = torch.cat((order[nt:], rest[t - nt:]))

    batch = dict(input_ids=input_ids, attention_mask=attention_mask, labels=labels,
                 neighbor_input_ids=nids, neighbor_attention_mask=nmask.long(), neighbor_pos_ids=tpos,
                 text_locations=tloc, neighbor_images=images, neighbor_images_pos_ids=ipos, image_locations=iloc)
    if spec.with_lpe:
        lpe = torch.randn((b, n + 1, n - 4), generator=g)
        batch["lpe"] = lpe / lpe.norm(dim=1, keepdim=True).clamp_min(1e-6)
    if spec.with_graph:
        valid = torch.zeros((b, n + 1), dtype=torch.bool)
        valid[:, 0] = True
        for r in range(b):
            valid[r, 1 + tloc[r][tpos[r] > 0]] = True
            valid[r, 1 + iloc[r][ipos[r] > 0]] = True
        a = (torch.rand((b, n + 1, n + 1), generator=g) < 2.0 / n).float()
        a = ((a + a.transpose(1, 2)) > 0).float() + torch.eye(n + 1)
        a = a * valid[:, :, None] * valid[:, None, :]
        batch["graph"] = a / a.sum(-1, keepdim=True).clamp_min(1.0)
    if pin:
        batch = {k: v.contiguous().pin_memory() for k, v in batch.items()}
    return batch


def batch_nbytes(batch: dict) -> int:
    return int(sum(v.numel() * v.element_size() for v in batch.values()))


def to_device(batch: dict, device, non_blocking=True) -> dict:
    """The reference's ``batch = {k: v.cuda(gpu, non_blocking=True)}`` (language_modelling/run_generation.py:464)."""
    return {k: v.to(device, non_blocking=non_blocking) for k, v in batch.items()}
