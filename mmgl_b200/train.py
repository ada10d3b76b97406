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
import os
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
                   grad_clip: Optional[float] = None, scheduler=None, grad_sync=None) -> float:
    """One optimizer step over ``accum_steps`` micro-batches.  ``loss_fn(net, batch) -> scalar loss tensor``
    (e.g. ``lambda net, b: net(**b).loss``).  ``net`` is either a DDP wrapper (all-reduce inside the last backward) or the
    bare module with ``grad_sync=FlatGradSync(module)``.  Returns the mean micro-batch loss (python float: one sync per step)."""
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
    if grad_sync is not None:          # FlatGradSync instead of DDP: one all-reduce of the accumulated gradients
        grad_sync.all_reduce()
    if grad_clip is not None and grad_clip > 0:
        torch.nn.utils.clip_grad_norm_([p for p in net.parameters() if p.requires_grad], grad_clip)
    optimizer.step()
    if torch.cuda.is_available() and not getattr(optimizer, "keeps_shadows_current", False):
        from . import ops   # optimizers that write through param.data (HF Adafactor, T5's default at run_generation.py
        ops.invalidate_weight_cache(trainable_only=True)   # :329-333) do not bump _version: drop the trainable shadows
    if scheduler is not None:
        scheduler.step()
    optimizer.zero_grad(set_to_none=True)
    return float(total)


class FlatGradSync:
    """Data-parallel gradient exchange as ONE all-reduce after ``backward()`` instead of DDP's bucketed all-reduces
    overlapped with it.

    Why (measured on 8 x B200, profiles/r02e_timeline_n8_*.txt): the hot kernels of this package are persistent -- one CTA
    per SM with ~200 KB of shared memory and a static tile schedule.  An NCCL all-reduce kernel that runs concurrently
    holds its 24 SMs for as long as it waits for the slowest rank, a GEMM CTA cannot share an SM with it, so the GEMM's
    148 CTAs run as 124 + 24 and every GEMM launched under an in-flight bucket takes up to twice as long: +4 ms of compute
    per 72 ms step, plus 2-5 ms of all-reduce tail that is exposed anyway.  The NVSwitch all-reduce of all 867 MB of fp32
    gradients in one piece takes ~2.3 ms (NVLS), which is less than what the overlap costs.

    The weight gradients (everything with >= 2 dims: 99.9 % of the bytes) live in ONE flat fp32 buffer: each parameter
    gets a slot, ``ops._wgrad`` writes the wgrad GEMM's fp32 output straight into it and autograd adopts that view as
    ``param.grad`` -- no flatten / unflatten copies.  Small gradients (biases, norms, gates) are flattened per step.
    A gradient that did not land in its slot (parameter used twice, produced by a torch op, gradient accumulation) is
    copied in; an unused parameter's slot is zeroed.  Use ``optimizer.zero_grad(set_to_none=True)`` between steps.

        sync = FlatGradSync(model)            # after prepare_for_training; broadcasts rank 0's parameters once
        loss.backward(); sync.all_reduce(); optimizer.step(); optimizer.zero_grad(set_to_none=True)
    """

    def __init__(self, module, process_group=None, broadcast: bool = True):
        self.group = process_group
        self.wire_bf16 = os.environ.get("MMGL_FLAT_SYNC_BF16", "0") == "1"
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        params = [p for p in module.parameters() if p.requires_grad]
        self.big = [p for p in params if p.dim() >= 2 and p.dtype == torch.float32]
        self.small = [p for p in params if not (p.dim() >= 2 and p.dtype == torch.float32)]
        n = sum(p.numel() for p in self.big)
        dev = self.big[0].device if self.big else (params[0].device if params else "cpu")
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        o = 0
        for p in self.big:
            p._mmgl_grad_flat = (self.flat, o)     # ops._wgrad makes a fresh view of the slot per use (autograd adopts it)
            p._mmgl_grad_slot_used = False
            o += p.numel()
        if broadcast and self.world > 1:
            with torch.no_grad():
                for t in list(module.parameters()) + list(module.buffers()):
                    dist.broadcast(t.data, src=dist.get_global_rank(process_group, 0) if process_group is not None else 0,
                                   group=process_group)

    @staticmethod
    def slot(p):
        flat, o = p._mmgl_grad_flat
        return flat[o:o + p.numel()].view(p.shape)

    def all_reduce(self):
        """Average the gradients over the ranks; afterwards every ``param.grad`` holds the mean (big ones as views of the
        flat buffer)."""
        with torch.no_grad():
            self.last_copied = 0           # big gradients that did not land in their slot this step (diagnostic)
            for p in self.big:
                s = self.slot(p)
                if p.grad is None:
                    s.zero_()
                elif p.grad.data_ptr() != s.data_ptr():
                    s.copy_(p.grad)
                    p.grad = s
                    self.last_copied += 1
                p._mmgl_grad_slot_used = False
            if self.world == 1:
                return
            small = [p for p in self.small if p.grad is not None]
            nccl = dist.get_backend(self.group) == "nccl"
            avg = dist.ReduceOp.AVG if (nccl and os.environ.get("MMGL_FLAT_SYNC_OP", "avg") == "avg") else dist.ReduceOp.SUM
            if self.wire_bf16:
                # optional (not the default: the reference's DDP averages fp32 gradients): halve the bytes on the wire
                wire = self.flat.to(torch.bfloat16)
                dist.all_reduce(wire, op=avg, group=self.group)
                self.flat.copy_(wire)
            else:
                dist.all_reduce(self.flat, op=avg, group=self.group)
            sflat = None
            if small:
                sdt = small[0].grad.dtype if all(p.grad.dtype == small[0].grad.dtype for p in small) else torch.float32
                sflat = torch.cat([p.grad.reshape(-1).to(sdt) for p in small])
                dist.all_reduce(sflat, op=avg, group=self.group)
            if avg == dist.ReduceOp.SUM:
                self.flat.div_(self.world)
                if sflat is not None:
                    sflat.div_(self.world)
            o = 0
            for p in small:
                p.grad.copy_(sflat[o:o + p.numel()].view(p.shape))
                o += p.numel()
