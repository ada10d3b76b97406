"""Training-loop harness with the semantics of the reference's ``train_loop``
(language_modelling/run_generation.py:462-494), minus its defects:

  * under gradient accumulation the DDP all-reduce runs only on the LAST micro-step (``no_sync`` on the others; the
    reference all-reduces every micro-step),
  * gradient clipping happens BEFORE ``optimizer.step()`` (the reference clips after the step, and only if > 2).

The data path is sharded by rank with no collective: each rank draws its own sections; the only exchange step is
the gradient all-reduce DDP performs inside ``backward()`` (NCCL over NVLink on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

import contextlib
from typing import Callable, Iterable, Optional

import torch
import torch.distributed as dist


def world_info():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def rank_seed(base_seed: int, rank: int, step: int) -> int:
    """Seed of the micro-batch a rank draws at a step: disjoint streams per rank (DistributedSampler's role,
    run_generation.py:366-368)."""
    return base_seed + 1_000_003 * rank + step


def max_over_ranks(value: float, device) -> float:
    t = torch.tensor([float(value)], device=device, dtype=torch.float64 if str(device) == "cpu" else torch.float32)
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def optimizer_step(net, optimizer, micro_batches: Iterable, loss_fn: Callable, accum_steps: int = 1,
                   grad_clip: Optional[float] = None, scheduler=None) -> float:
    """One optimizer step over ``accum_steps`` micro-batches.  ``loss_fn(net, batch) -> scalar loss tensor``
    (e.g. ``lambda net, b: net(**b).loss``).  Returns the mean micro-batch loss (python float: one sync per step)."""
    micro_batches = list(micro_batches)
    assert len(micro_batches) == accum_steps
    total = None
    for i, batch in enumerate(micro_batches):
        last = i == accum_steps - 1
        ctx = net.no_sync() if (hasattr(net, "no_sync") and not last) else contextlib.nullcontext()
        with ctx:
            loss = loss_fn(net, batch) / accum_steps
            loss.backward()
        total = loss.detach() if total is None else total + loss.detach()
    if grad_clip is not None and grad_clip > 0:
        torch.nn.utils.clip_grad_norm_([p for p in net.parameters() if p.requires_grad], grad_clip)
    optimizer.step()
    if torch.cuda.is_available():
        from . import ops   # optimizers that write through param.data (HF Adafactor, T5's default at run_generation.py
        ops.invalidate_weight_cache(trainable_only=True)   # :329-333) do not bump _version: drop the trainable shadows
    if scheduler is not None:
        scheduler.step()
    optimizer.zero_grad(set_to_none=True)
    return float(total)
