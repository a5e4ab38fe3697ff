"""Data-parallel plumbing for SpotV2Net training.

Graph snapshots are independent, so ranks shard each global batch by snapshot and exchange only
gradients: every parameter's ``.grad`` is a view into ONE flat fp32 arena and a single all-reduce per
step moves it (NCCL over NVLink/NVSwitch on GPUs; gloo in the CPU tests).  The reference has no
distributed code (its only parallelism is one process per seed, 5_train_SpotV2Net.py:214-218).
"""
from __future__ import annotations

from typing import Iterable, List, Sequence

import torch
import torch.distributed as dist


class FlatGradArena:
    """One contiguous buffer holding the gradients of ``params`` in order."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev, dt = self.params[0].device, self.params[0].dtype
        sizes = [p.numel() for p in self.params]
        self.flat = torch.zeros(sum(sizes), device=dev, dtype=dt)
        self.views = [v.view_as(p) for v, p in zip(torch.split(self.flat, sizes), self.params)]
        for p, v in zip(self.params, self.views):
            p.grad = v                      # autograd accumulates in place, so the view survives backward

    def zero(self) -> None:
        self.flat.zero_()
        for p, v in zip(self.params, self.views):   # re-attach if an optimizer set grads to None
            if p.grad is None or p.grad.data_ptr() != v.data_ptr():
                p.grad = v

    def all_reduce(self, group=None, average: bool = True) -> None:
        """Sum (or average) the arena over the ranks of ``group``; no-op without a process group."""
        if not (dist.is_available() and dist.is_initialized()):
            return
        world = dist.get_world_size(group)
        if world == 1:
            return
        dist.all_reduce(self.flat, group=group)
        if average:
            self.flat.mul_(1.0 / world)

    @property
    def nbytes(self) -> int:
        return self.flat.numel() * self.flat.element_size()


def shard_snapshots(indices: Sequence[int] | torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """Rank r takes snapshots r::world of the global batch (equal shards keep mean-loss gradients exact)."""
    idx = torch.as_tensor(indices)
    if idx.numel() % world:
        raise ValueError(f"global batch of {idx.numel()} snapshots does not split evenly over {world} ranks")
    return idx[rank::world]
