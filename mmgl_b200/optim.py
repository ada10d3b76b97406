"""AdamW on the fp32 master weights with the kernels' bf16 operand copies ("shadows") written in the same pass.

Drop-in for the ``torch.optim.AdamW`` the reference builds (language_modelling/run_generation.py:329-333, stepped at :486):
same constructor arguments, same update (amsgrad / maximize / foreach / capturable are not supported and raise), same
``state_dict`` layout (``step``, ``exp_avg``, ``exp_avg_sq`` per parameter), so optimizer checkpoints are interchangeable.
Each fp32 CUDA parameter is updated by ``mmgl_adamw_step`` (csrc/optim.cu): one read of (p, g, m, v), one write of
(p, m, v, bf16 p).  The bf16 copy is registered with ``ops.w16``'s cache, so the next forward finds its operand ready and
the per-parameter fp32 -> bf16 conversion kernels disappear from the step.  Parameters that are not contiguous fp32 CUDA
tensors raise (no eager fallback).
"""
from __future__ import annotations

import torch

from . import _capi as K
from . import ops


class FusedAdamW(torch.optim.Optimizer):
    keeps_shadows_current = True      # train.optimizer_step: no need to drop the bf16 shadows after step()

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, amsgrad=False, *, maximize=False,
                 foreach=None, capturable=False, differentiable=False, fused=None):
        if amsgrad or maximize or capturable or differentiable:
            raise NotImplementedError("FusedAdamW: amsgrad / maximize / capturable / differentiable are not supported")
        if not 0.0 <= lr or not 0.0 <= eps or not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0 or not 0.0 <= weight_decay:
            raise ValueError("FusedAdamW: invalid hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))

    @torch.no_grad()
    def step(self, closure=None, grad_scale: float = 1.0):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            lr, (b1, b2), eps, wd = group["lr"], group["betas"], group["eps"], group["weight_decay"]
            lr = float(lr)
            for p in group["params"]:
                if p.grad is None:
                    continue
                if p.grad.is_sparse:
                    raise RuntimeError("FusedAdamW does not support sparse gradients")
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0, dtype=torch.float32)       # host scalar, as torch's non-capturable AdamW keeps it
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
                t = int(st["step"])
                g = p.grad
                fast = (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() and g.dtype == torch.float32
                        and st["exp_avg"].is_contiguous() and st["exp_avg_sq"].is_contiguous())
                if fast:
                    if not g.is_contiguous():
                        g = g.contiguous()
                    shadow = None
                    if p.dim() >= 2:                      # GEMM operands: keep a bf16 copy current (biases / norms are read as fp32)
                        shadow = st.get("_shadow")
                        if shadow is None or shadow.shape != p.shape or shadow.device != p.device:
                            shadow = st["_shadow"] = torch.empty_like(p, dtype=torch.bfloat16)
                    K.adamw_step(p, g, st["exp_avg"], st["exp_avg_sq"], shadow, lr, b1, b2, eps, wd, t, grad_scale)
                    torch.autograd.graph.increment_version(p)     # the kernel wrote through the raw pointer
                    if shadow is not None:
                        ops.register_shadow(p, shadow)
                else:
                    raise NotImplementedError(
                        "FusedAdamW updates contiguous fp32 CUDA parameters with fp32 gradients (the master weights "
                        f"prepare_for_training() leaves trainable); got {p.dtype} on {p.device} with a {g.dtype} gradient -- "
                        "there is no eager fallback, use torch.optim.AdamW for such parameters")
        return loss

    def state_dict(self):
        sd = super().state_dict()
        # the bf16 shadows are derived data: not part of a checkpoint (copy: the inner dicts are the live state)
        sd["state"] = {k: {n: v for n, v in st.items() if n != "_shadow"} for k, st in sd["state"].items()}
        return sd
