"""CPU oracle for the optimizer side of the KD step — TEST INFRASTRUCTURE ONLY (see oracle/kd_oracle.py's header).

Plain fp64/fp32 tensor arithmetic restating what the reference's training loop does after ``backward`` with library calls
(src/train_student_kd.py):
  :290  scaler.unscale_(optimizer)                      -> grads *= 1/scale, remember whether any is non-finite
  :293  clip_grad_norm_(student_model.parameters(), 1)  -> one global L2 norm over the student's gradients
  :296  clip_grad_norm_(projector.parameters(), 1)      -> a second, separate norm per projector
  :299  scaler.step(optimizer)                          -> AdamW (three LR groups, weight_decay 0.01; :230-234) unless inf/nan was found
  :300  scaler.update()                                 -> scale *= 0.5 after a skipped step, *= 2 after 2000 clean ones
  :303  scheduler.step(epoch + i/len)                   -> CosineAnnealingWarmRestarts(T_0=5, T_mult=2, eta_min=1e-6) (:236)

Parity pin: the algorithm lives in torch (a dependency of the reference, any 2.x): tests/test_optim_oracle.py runs the same
library calls the reference makes (torch.optim.AdamW, torch.nn.utils.clip_grad_norm_, torch.amp.GradScaler semantics,
CosineAnnealingWarmRestarts) on the CPU and checks every function below against them.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch

Tensor = torch.Tensor


def clip_coefficient(grads: Sequence[Tensor], max_norm: float) -> (float, float):
    """torch.nn.utils.clip_grad_norm_: total = ||(||g_i||_2)_i||_2, coef = min(1, max_norm / (total + 1e-6))."""
    total = math.sqrt(sum(float((g.double() ** 2).sum()) for g in grads))
    coef = max_norm / (total + 1e-6)
    return total, min(coef, 1.0)


def adamw_update(p: Tensor, g: Tensor, m: Tensor, v: Tensor, step: int, lr: float, beta1: float = 0.9, beta2: float = 0.999,
                 eps: float = 1e-8, weight_decay: float = 0.01):
    """One AdamW update of one tensor (torch.optim.AdamW, amsgrad=False, maximize=False); returns the new (p, m, v)."""
    p = p * (1.0 - lr * weight_decay)                    # decoupled weight decay comes first
    m = beta1 * m + (1.0 - beta1) * g
    v = beta2 * v + (1.0 - beta2) * g * g
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    denom = v.sqrt() / math.sqrt(bc2) + eps
    p = p - (lr / bc1) * (m / denom)
    return p, m, v


class OptimizerOracle:
    """State + step for a list of parameter groups ``{"params": [tensors], "lr", "weight_decay", "clip_group"}``.

    Works in fp64 internally (``dtype``) so the fp32 CUDA kernels can be checked against something strictly more accurate."""

    def __init__(self, groups: List[dict], betas=(0.9, 0.999), eps=1e-8, max_norm: Optional[float] = 1.0,
                 loss_scale: Optional[float] = None, growth_factor=2.0, backoff_factor=0.5, growth_interval=2000, dtype=torch.float64):
        self.groups = [{"params": [p.detach().to(dtype).clone() for p in g["params"]], "lr": g["lr"],
                        "weight_decay": g.get("weight_decay", 0.01), "clip_group": g.get("clip_group", 0)} for g in groups]
        for g in self.groups:
            g["m"] = [torch.zeros_like(p) for p in g["params"]]
            g["v"] = [torch.zeros_like(p) for p in g["params"]]
        self.betas, self.eps, self.max_norm, self.dtype = betas, eps, max_norm, dtype
        self.step_count = 0
        self.loss_scale, self.growth_tracker = loss_scale, 0
        self.growth_factor, self.backoff_factor, self.growth_interval = growth_factor, backoff_factor, growth_interval
        self.last_norms: Dict[int, float] = {}
        self.last_skipped = False

    def step(self, grads: List[List[Tensor]]) -> None:
        """``grads[gi][pi]`` is the (still loss-scaled) gradient of parameter pi of group gi."""
        inv = 1.0 if self.loss_scale is None else 1.0 / self.loss_scale
        grads = [[g.to(self.dtype) * inv for g in gg] for gg in grads]                       # :290 unscale_
        found_inf = any((not bool(torch.isfinite(g).all())) for gg in grads for g in gg)
        self.last_skipped = found_inf
        clip_groups = sorted({g["clip_group"] for g in self.groups if g["clip_group"] >= 0})
        coef = {}
        for cg in clip_groups:                                                                # :293-297 one norm per clip group
            members = [g for gi, grp in enumerate(self.groups) if grp["clip_group"] == cg for g in grads[gi]]
            if self.max_norm:
                self.last_norms[cg], coef[cg] = clip_coefficient(members, self.max_norm)
            else:
                self.last_norms[cg], coef[cg] = clip_coefficient(members, 1.0)[0], 1.0
        if not found_inf:                                                                     # :299 scaler.step
            self.step_count += 1
            for gi, grp in enumerate(self.groups):
                c = coef.get(grp["clip_group"], 1.0)
                for pi in range(len(grp["params"])):
                    grp["params"][pi], grp["m"][pi], grp["v"][pi] = adamw_update(
                        grp["params"][pi], grads[gi][pi] * c, grp["m"][pi], grp["v"][pi], self.step_count, grp["lr"],
                        self.betas[0], self.betas[1], self.eps, grp["weight_decay"])
        if self.loss_scale is not None:                                                       # :300 scaler.update
            if found_inf:
                self.loss_scale *= self.backoff_factor
                self.growth_tracker = 0
            else:
                self.growth_tracker += 1
                if self.growth_tracker == self.growth_interval:
                    self.loss_scale *= self.growth_factor
                    self.growth_tracker = 0


def cosine_warm_restarts_lr(epoch: float, base_lr: float, T_0: int = 5, T_mult: int = 2, eta_min: float = 1e-6) -> float:
    """Closed form of CosineAnnealingWarmRestarts at a fractional epoch: find the restart cycle by walking the cycle
    lengths T_0, T_0*T_mult, ... (a loop instead of the library's logarithm, so the two can disagree only by a bug)."""
    start, length = 0.0, float(T_0)
    while epoch >= start + length:
        start += length
        length *= T_mult
    return eta_min + (base_lr - eta_min) * (1.0 + math.cos(math.pi * (epoch - start) / length)) / 2.0
