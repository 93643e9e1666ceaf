"""Optimizer side of the KD step on flat buffers: unscale + per-group global-norm clip + AdamW in two native launches.

Mirrors the reference's optimizer block (src/train_student_kd.py):
  :219-234  ``optim.AdamW([{encoder, lr*0.1}, {decoder, lr}, {refinement + projectors, lr}], weight_decay=0.01)``
  :236      ``CosineAnnealingWarmRestarts(optimizer, T_0=5, T_mult=2, eta_min=1e-6)``
  :239      ``GradScaler('cuda')``
  :290-303  ``scaler.unscale_``; ``clip_grad_norm_(student_model.parameters(), 1.0)``; ``clip_grad_norm_(projector.parameters(), 1.0)``
            for every projector; ``scaler.step``; ``scaler.update``; ``optimizer.zero_grad``; ``scheduler.step(epoch + i/len)``

``FlatAdamW`` re-homes every parameter into ONE contiguous fp32 buffer (``p.data`` becomes a view), keeps both Adam
moments in equally laid out buffers and owns the flat gradient buffer (a ``FlatGradAllReducer``), so the whole update is
``b2c_optimizer_step`` on four arrays.  Learning rates, the step count and the loss scale live on the device: a captured
CUDA graph of the step stays valid when the scheduler changes the rates.

Create it BEFORE anything captures parameter addresses (CUDA graphs, ``_ops.set_grad_destinations``).
"""
from __future__ import annotations

import ctypes
import math
from typing import Dict, Iterable, List, Optional, Sequence

import torch

from . import _ops
from .ddp import FlatGradAllReducer


def reference_param_groups(student_model, projectors, learning_rate: float, active=("encoder",)):
    """The reference's three LR groups and two clip groups (train_student_kd.py:219-234, :293-297) as FlatAdamW groups.

    `active`: keys of the `projectors` dict whose parameters take part in the step.  The reference hands every projector to
    AdamW (:227-228), but only projectors['encoder'] is ever called in its loop (:281), so the 'hidden' projector's gradients
    stay None and torch.optim.AdamW skips it entirely (no moment update, no weight decay).  A flat-buffer optimizer cannot see
    "grad is None" -- an untouched gradient slot reads as zeros and would still be decayed -- so parameters that never receive
    a gradient are left out of the flat buffers here, which gives the same result."""
    other = []
    if getattr(student_model, "use_attention_refinement", False):
        other += list(student_model.attention_refinement.parameters())
    proj = []
    if isinstance(projectors, dict):
        for key, pr in projectors.items():
            if active is None or key in active:
                proj += list(pr.parameters())
    else:
        proj += list(projectors.parameters())
    return [
        {"params": list(student_model.encoder.parameters()), "lr": learning_rate * 0.1, "clip_group": 0},
        {"params": list(student_model.decoder.parameters()), "lr": learning_rate, "clip_group": 0},
        {"params": other, "lr": learning_rate, "clip_group": 0},
        {"params": proj, "lr": learning_rate, "lr_group": 2, "clip_group": 1},     # same LR group as `other`, its own clip norm
    ]


class FlatAdamW:
    """AdamW + clip_grad_norm_ + GradScaler semantics over flat fp32 buffers, executed by ``b2c_optimizer_step``.

    ``param_groups``: list of dicts with ``params`` and optional ``lr``, ``weight_decay``, ``clip_group`` (gradients of one
    clip group share one global L2 norm, default 0; -1 = never clipped) and ``lr_group`` (index into ``self.lr``; default:
    one rate per group).  Frozen parameters (requires_grad False) are skipped, empty groups are dropped.
    """

    def __init__(self, param_groups, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.01,
                 max_grad_norm: Optional[float] = 1.0, loss_scale: Optional[float] = None, growth_factor: float = 2.0,
                 backoff_factor: float = 0.5, growth_interval: int = 2000, process_group=None, single_process: bool = False):
        if not isinstance(param_groups, (list, tuple)) or (param_groups and not isinstance(param_groups[0], dict)):
            param_groups = [{"params": list(param_groups)}]
        groups = []
        for i, g in enumerate(param_groups):
            ps = [p for p in g["params"] if p.requires_grad]
            if ps:
                groups.append({"params": ps, "lr": float(g.get("lr", lr)), "weight_decay": float(g.get("weight_decay", weight_decay)),
                               "clip_group": int(g.get("clip_group", 0)), "lr_group": g.get("lr_group", None), "index": i})
        if not groups:
            raise ValueError("no trainable parameters")
        if len(groups) > _ops.B2C_OPT_MAX_SEG:
            raise ValueError(f"at most {_ops.B2C_OPT_MAX_SEG} parameter groups")
        lr_slots: Dict[int, int] = {}
        for g in groups:                                   # lr_group defaults to the group's own position in the caller's list
            key = g["index"] if g["lr_group"] is None else int(g["lr_group"])
            g["lr_index"] = lr_slots.setdefault(key, len(lr_slots))
        self.param_groups = groups
        self.params: List[torch.nn.Parameter] = [p for g in groups for p in g["params"]]
        if len({id(p) for p in self.params}) != len(self.params):
            raise ValueError("a parameter appears in more than one group")
        dev = self.params[0].device
        if dev.type != "cuda":
            raise RuntimeError("FlatAdamW runs on the GPU only (no CPU fallback); parameters must be CUDA tensors")
        if any(p.dtype != torch.float32 for p in self.params):
            raise TypeError("FlatAdamW keeps fp32 master parameters; cast the model to float32")
        # layout: groups back to back, every group start rounded up to 4 elements (16 bytes)
        starts = {}
        off = 0
        self.offsets: List[int] = []
        self.segments = (_ops.B2COptSegment * len(groups))()
        for gi, g in enumerate(groups):
            off = (off + 3) // 4 * 4
            begin = off
            for p in g["params"]:
                self.offsets.append(off)
                off += p.numel()
            self.segments[gi] = _ops.B2COptSegment(begin, off, g["lr_index"], g["clip_group"], g["weight_decay"])
            g["range"] = (begin, off)
        self.numel = off
        self.flat_param = torch.zeros(off, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for p, o in zip(self.params, self.offsets):
                view = self.flat_param[o:o + p.numel()].view_as(p)
                view.copy_(p.data)
                p.data = view
        self.exp_avg = torch.zeros_like(self.flat_param)
        self.exp_avg_sq = torch.zeros_like(self.flat_param)
        self.reducer = FlatGradAllReducer(self.params, group=process_group, offsets=self.offsets, numel=off, single_process=single_process)
        # grad_scale: the data-parallel average (1/world after the SUM all-reduce) is folded into the update kernel when the
        # caller asks for it (GraphedKDStep does); FlatGradAllReducer.finish() then must not be called.
        self.hyper = _ops.B2COptHyper(betas[0], betas[1], eps, float(max_grad_norm) if max_grad_norm else 0.0, 1.0,
                                      growth_factor, backoff_factor, growth_interval)
        self.lr = torch.zeros(len(lr_slots), dtype=torch.float32, device=dev)
        self._lr_host = torch.zeros(len(lr_slots), dtype=torch.float32).pin_memory()
        for g in groups:
            self._lr_host[g["lr_index"]] = g["lr"]
        self.base_lrs = self._lr_host.tolist()
        self.lr.copy_(self._lr_host)
        self.step_count = torch.zeros(1, dtype=torch.int32, device=dev)
        self.loss_scale = None if loss_scale is None else torch.full((1,), float(loss_scale), dtype=torch.float32, device=dev)
        self.growth_tracker = None if loss_scale is None else torch.zeros(1, dtype=torch.int32, device=dev)
        self.stats = torch.zeros(_ops.B2C_OPT_NSTATS, dtype=torch.float32, device=dev)
        self._scratch = torch.zeros(_ops.B2C_OPT_SCRATCH_BYTES, dtype=torch.uint8, device=dev)

    # ---- learning rates (host -> device array; the kernels read the device copy)
    def set_lrs(self, lrs: Sequence[float]) -> None:
        if len(lrs) != self.lr.numel():
            raise ValueError(f"{self.lr.numel()} learning-rate groups, got {len(lrs)}")
        for i, v in enumerate(lrs):
            self._lr_host[i] = float(v)
        self.lr.copy_(self._lr_host, non_blocking=True)

    def get_lrs(self) -> List[float]:
        return self._lr_host.tolist()

    def fold_average(self, world_size: int) -> None:
        """Gradients arrive as a SUM over `world_size` ranks: apply the 1/world inside the update kernel (no separate pass)."""
        self.hyper.grad_scale = 1.0 / float(world_size)

    # ---- the step
    def step(self) -> None:
        """unscale + clip + AdamW + scaler.update on the flat buffers (asynchronous; statistics land in ``self.stats``)."""
        lib = _ops.load_library()
        ls = self.loss_scale.data_ptr() if self.loss_scale is not None else None
        gt = self.growth_tracker.data_ptr() if self.growth_tracker is not None else None
        _ops._check(lib.b2c_optimizer_step(self.flat_param.data_ptr(), self.reducer.flat.data_ptr(), self.exp_avg.data_ptr(),
                                           self.exp_avg_sq.data_ptr(), self.segments, len(self.segments), ctypes.byref(self.hyper),
                                           self.lr.data_ptr(), self.lr.numel(), self.step_count.data_ptr(), ls, gt,
                                           self.stats.data_ptr(), self._scratch.data_ptr(), _ops._stream()), "b2c_optimizer_step")

    def zero_grad(self, set_to_none: bool = False) -> None:
        self.reducer.zero_grad()

    def grad_norms(self) -> List[float]:
        """Pre-clip gradient norm of every clip group at the last step (host sync)."""
        return self.stats[:_ops.B2C_OPT_MAX_CLIP].tolist()

    def last_step_skipped(self) -> bool:
        return bool(self.stats[_ops.B2C_OPT_MAX_CLIP].item() != 0.0)

    def state_dict(self) -> dict:
        return {"exp_avg": self.exp_avg.clone(), "exp_avg_sq": self.exp_avg_sq.clone(), "step": int(self.step_count.item()),
                "lr": self.get_lrs(), "loss_scale": None if self.loss_scale is None else float(self.loss_scale.item()),
                "growth_tracker": None if self.growth_tracker is None else int(self.growth_tracker.item())}

    def load_state_dict(self, sd: dict) -> None:
        self.exp_avg.copy_(sd["exp_avg"]); self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        self.step_count.fill_(int(sd["step"]))
        self.set_lrs(sd["lr"])
        if self.loss_scale is not None and sd.get("loss_scale") is not None:
            self.loss_scale.fill_(float(sd["loss_scale"]))
            self.growth_tracker.fill_(int(sd.get("growth_tracker") or 0))


def cosine_warm_restarts(epoch: float, base_lr: float, T_0: int = 5, T_mult: int = 2, eta_min: float = 1e-6) -> float:
    """Learning rate of ``CosineAnnealingWarmRestarts(T_0, T_mult, eta_min).step(epoch)`` for a (fractional) epoch
    (reference: src/train_student_kd.py:236 creates it, :303 steps it with ``epoch + batch_idx / len(train_loader)``).

    Cycle i lasts T_0 * T_mult**i epochs; inside a cycle the rate follows half a cosine from base_lr down to eta_min."""
    if epoch < 0:
        raise ValueError("epoch must be non-negative")
    if T_0 <= 0 or T_mult < 1:
        raise ValueError("T_0 must be positive and T_mult >= 1")
    if epoch < T_0:
        t_cur, t_i = float(epoch), float(T_0)
    elif T_mult == 1:
        t_cur, t_i = math.fmod(epoch, T_0), float(T_0)
    else:
        n = int(math.log(epoch / T_0 * (T_mult - 1) + 1, T_mult))       # index of the cycle that contains `epoch`
        t_cur = epoch - T_0 * (T_mult ** n - 1) / (T_mult - 1)
        t_i = float(T_0 * T_mult ** n)
    return eta_min + (base_lr - eta_min) * (1.0 + math.cos(math.pi * t_cur / t_i)) / 2.0


class CosineWarmRestarts:
    """Scheduler object for FlatAdamW with the reference's call pattern: ``scheduler.step(epoch + batch_idx / n_batches)``."""

    def __init__(self, optimizer: FlatAdamW, T_0: int = 5, T_mult: int = 2, eta_min: float = 1e-6):
        self.optimizer, self.T_0, self.T_mult, self.eta_min = optimizer, T_0, T_mult, eta_min
        self.base_lrs = list(optimizer.base_lrs)
        self.last_epoch = 0.0

    def step(self, epoch: Optional[float] = None) -> None:
        self.last_epoch = self.last_epoch + 1 if epoch is None else float(epoch)
        self.optimizer.set_lrs([cosine_warm_restarts(self.last_epoch, b, self.T_0, self.T_mult, self.eta_min) for b in self.base_lrs])

    def get_last_lr(self) -> List[float]:
        return self.optimizer.get_lrs()
