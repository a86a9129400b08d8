"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink on the B200 box,
gloo in the CPU tests).  The reference has no distributed code at all; the only
collective the training path needs when every rank holds its own box replica is the gradient
all-reduce below.
"""
from __future__ import annotations

import os
from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


def init_from_env(backend: Optional[str] = None) -> (int, int, int):
    """(rank, world_size, local_rank) from the torchrun environment; initialises the process group."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, world, local


class GradientBucket:
    """One flat fp32 buffer for all parameter gradients -> a single all-reduce per step
    (1.6 M floats = 6.5 MB at L=128: latency-sized, so one bucket)."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params: List[torch.nn.Parameter] = [p for p in params]
        self.flat: Optional[torch.Tensor] = None

    def _ensure(self):
        live = [p for p in self.params if p.grad is not None]
        total = sum(p.grad.numel() for p in live)
        if self.flat is None or self.flat.numel() != total or self.flat.device != live[0].grad.device:
            self.flat = torch.empty(total, dtype=torch.float32, device=live[0].grad.device)
        return live

    def all_reduce(self, average: bool = True, group=None) -> None:
        """Sums (or averages) `.grad` over ranks in place.  Parameters whose grad is None (the dead
        edge stream in message='sender' mode) are skipped consistently on every rank."""
        if not dist.is_initialized() or dist.get_world_size(group) == 1:
            return
        live = self._ensure()
        torch._foreach_copy_(list(torch.split(self.flat, [p.grad.numel() for p in live])),
                             [p.grad.reshape(-1) for p in live])
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        if average:
            self.flat.div_(dist.get_world_size(group))
        torch._foreach_copy_([p.grad.reshape(-1) for p in live],
                             list(torch.split(self.flat, [p.grad.numel() for p in live])))


def max_over_ranks(value: float, device) -> float:
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
