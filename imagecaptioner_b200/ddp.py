"""Batch-sharded data parallelism for the KD step: one process per GPU, one flat gradient all-reduce.

The reference has no distributed code; BASELINE.json's north_star adds data
parallelism over the batch with NCCL only for the gradient all-reduce.  The step shards with no
data-path collective (samples are independent through decoder, projector and every loss term), so
the only exchanges are

  * the gradient all-reduce (sum) over ONE flat fp32 buffer holding every trainable parameter's
    gradient (7.33 M floats for the default student + projector), scaled by 1/world afterwards, and
  * one int32 all-reduce of the non-PAD target count, because CrossEntropyLoss(ignore_index=0) divides
    by the GLOBAL count (reference src/distillation_utils.py:22); KL 'batchmean', the MSE terms and the
    cosine term have equal per-rank denominators, so their mean of per-rank means is already exact.

Works with any torch.distributed backend (NCCL on the B200 box, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


class FlatGradAllReducer:
    """Owns one contiguous fp32 buffer; every parameter's ``.grad`` is a view into it, so autograd
    accumulates straight into the buffer and the collective is a single call with no packing copies."""

    def __init__(self, params: Iterable[torch.nn.Parameter], group: Optional[dist.ProcessGroup] = None,
                 offsets: Optional[List[int]] = None, numel: Optional[int] = None, single_process: bool = False):
        """``offsets`` / ``numel``: optional explicit layout (element offset of every trainable parameter, total length), used
        by optim.FlatAdamW to keep the gradient buffer congruent with its parameter and moment buffers."""
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev = self.params[0].device
        self.group = group
        # single_process: no exchange even when torch.distributed is initialised (a one-GPU reference run inside a multi-rank job)
        self.world_size = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized() and not single_process) else 1
        if offsets is None:
            offsets, off = [], 0
            for p in self.params:
                offsets.append(off)
                off += p.numel()
            numel = off
        if len(offsets) != len(self.params) or numel is None or any(o + p.numel() > numel for o, p in zip(offsets, self.params)):
            raise ValueError("offsets / numel do not describe the trainable parameters")
        self.offsets = list(offsets)
        self.flat = torch.zeros(numel, dtype=torch.float32, device=dev)
        self.attach_views()

    def _slot(self, i: int) -> torch.Tensor:
        p, o = self.params[i], self.offsets[i]
        return self.flat[o:o + p.numel()].view_as(p)

    def grad_views(self):
        """parameter data_ptr -> its fp32 slot in the flat buffer (for _ops.set_grad_destinations)."""
        return {p.data_ptr(): self._slot(i) for i, p in enumerate(self.params)}

    def detach_grads(self) -> None:
        """`.grad = None` on every parameter: the next backward's gradient tensors are adopted as-is (no add kernel)."""
        for p in self.params:
            p.grad = None

    def attach_views(self) -> None:
        """Point every parameter's .grad at its slot of the flat buffer (host-side only, no kernel)."""
        base = self.flat.data_ptr()
        for i, p in enumerate(self.params):
            if p.grad is None or p.grad.data_ptr() != base + self.offsets[i] * 4 or p.grad.dtype != torch.float32:
                p.grad = self._slot(i)

    def zero_grad(self) -> None:
        """Zero the buffer in place (keeps the .grad views; do NOT call optimizer.zero_grad(set_to_none=True))."""
        self.flat.zero_()
        self.attach_views()                        # re-attach views if something replaced them

    def allreduce(self, async_op: bool = False):
        """Sum over ranks, then average.  With async_op the caller waits on the returned work and calls finish()."""
        if self.world_size == 1:
            return None
        work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group, async_op=async_op)
        if async_op:
            return work
        self.finish()
        return None

    def finish(self) -> None:
        if self.world_size > 1:
            self.flat.mul_(1.0 / self.world_size)


def attach_loss_group(loss_module, group: Optional[dist.ProcessGroup] = None) -> None:
    """Make DistillationLoss use the global non-PAD count (and the matching CE scale) under data parallelism."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        loss_module.process_group = group if group is not None else dist.group.WORLD
        loss_module.world_size = dist.get_world_size(group)
    else:
        loss_module.process_group = None
        loss_module.world_size = 1


def shard_batch(n_global: int, rank: int, world_size: int) -> slice:
    """Contiguous, even split of a global batch (the bench uses weak scaling: fixed per-rank batch)."""
    if n_global % world_size != 0:
        raise ValueError(f"global batch {n_global} is not divisible by world size {world_size}")
    per = n_global // world_size
    return slice(rank * per, (rank + 1) * per)
