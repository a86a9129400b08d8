"""Fused optimizer step for the training loop around the hot path.

`FusedAdam` does what the reference's `torch.optim.Adam(simulator.parameters(), lr, weight_decay)` +
`ExponentialLR(optimizer, gamma)` do (train.py:183-187, 263-265, 316), in ONE kernel launch per step for the whole
model: the parameters are re-pointed at views of one flat FP32 buffer (so a single pass covers all ~110 tensors),
the gradients are gathered into a flat buffer with one multi-tensor copy — the same buffer the gradient all-reduce
of `distributed.GradientBucket` uses — and `cgnn_adam_step` (csrc/optim.cu) updates parameters and both moments.
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch

from ._lib import check, lib, ptr, stream_ptr


class FusedAdam:
    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("FusedAdam: no parameters")
        dev = self.params[0].device
        if dev.type != "cuda":
            raise RuntimeError("FusedAdam runs on CUDA parameters only (no CPU path)")
        for p in self.params:
            if isinstance(p, torch.nn.parameter.UninitializedParameter):
                raise RuntimeError("FusedAdam: materialise the lazy layers (one forward) before building the optimizer")
            if p.dtype != torch.float32 or p.device != dev:
                raise TypeError("FusedAdam: parameters must be float32 on one CUDA device")
        self.lr, self.betas, self.eps, self.weight_decay = float(lr), (float(betas[0]), float(betas[1])), float(eps), float(weight_decay)
        self.step_count = 0
        self.sizes = [p.numel() for p in self.params]
        n = sum(self.sizes)
        self.flat = torch.empty(n, dtype=torch.float32, device=dev)
        # parameters become views of the flat buffer (values preserved); state_dict / load_state_dict keep working
        off = 0
        with torch.no_grad():
            for p, sz in zip(self.params, self.sizes):
                view = self.flat[off:off + sz].view_as(p)
                view.copy_(p)
                p.data = view
                off += sz
        self.grad = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)

    def zero_grad(self, set_to_none: bool = True) -> None:
        for p in self.params:
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()

    def gather_grads(self) -> torch.Tensor:
        """Copies every `.grad` into the flat gradient buffer (zeros where a parameter got no gradient: the dead edge
        stream in message='sender' mode) and returns it — the buffer to all-reduce when training on several GPUs."""
        views = list(torch.split(self.grad, self.sizes))
        live = [(v, p.grad.reshape(-1)) for v, p in zip(views, self.params) if p.grad is not None]
        if len(live) != len(self.params):
            self.grad.zero_()
        if live:
            torch._foreach_copy_([a for a, _ in live], [b for _, b in live])
        return self.grad

    def step(self, grad_scale: float = 1.0, gathered: bool = False) -> None:
        """One Adam update with the current `lr`.  `gathered=True`: the flat gradient buffer already holds the
        (all-reduced) gradients from `gather_grads()`."""
        if not gathered:
            self.gather_grads()
        self.step_count += 1
        with torch.cuda.device(self.flat.device):
            check(lib().cgnn_adam_step(ptr(self.flat), ptr(self.grad), ptr(self.exp_avg), ptr(self.exp_avg_sq), self.flat.numel(),
                                       self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay, self.step_count,
                                       float(grad_scale), stream_ptr(self.flat.device)), "cgnn_adam_step")


class ExponentialLR:
    """`torch.optim.lr_scheduler.ExponentialLR` for `FusedAdam`: lr <- lr * gamma per `step()` (train.py:186, 316)."""

    def __init__(self, optimizer: FusedAdam, gamma: float):
        self.optimizer, self.gamma = optimizer, float(gamma)
        self.base_lr = optimizer.lr
        self.last_epoch = 0

    def step(self) -> None:
        self.last_epoch += 1
        self.optimizer.lr = self.base_lr * self.gamma ** self.last_epoch

    def get_last_lr(self):
        return [self.optimizer.lr]
