"""Fused optimizer step for the training loop around the hot path.

`FusedAdam` does what the reference's `torch.optim.Adam(simulator.parameters(), lr, weight_decay)` +
`ExponentialLR(optimizer, gamma)` do (train.py:183-187, 263-265, 316), in ONE kernel launch per step for the whole
model: the parameters are re-pointed at views of one flat FP32 buffer (so a single pass covers all ~110 tensors),
the gradients are gathered into a flat buffer with one multi-tensor copy and `cgnn_adam_step_masked`
(csrc/optim.cu) updates parameters and both moments.

* Parameters whose `.grad` is None (the dead edge stream of `message="sender"`) are skipped entirely -- no weight
  decay, no moment update -- as `torch.optim.Adam` does.
* `step_overlapped` is the multi-GPU form: the flat gradient is all-reduced in buckets and the update of bucket i
  runs while bucket i+1 is still in flight (the collective is the only one of the training path).
* `state_dict()` / `load_state_dict()` use `torch.optim.Adam`'s layout (per-parameter `step`, `exp_avg`,
  `exp_avg_sq`), so either optimizer resumes from the other's checkpoint; `save_training_state` /
  `load_training_state` add what the reference's checkpoints lack for a true resume (train.py:329-351 saves the
  model weights only): optimizer, scheduler, epoch and the RNG streams `preprocess` draws its noise from.
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
        self.sizes = [p.numel() for p in self.params]
        self.steps = [0] * len(self.params)               # per parameter, like torch.optim.Adam: a step without gradient does not count
        n = sum(self.sizes)
        self.flat = torch.empty(n, dtype=torch.float32, device=dev)
        # parameters become views of the flat buffer (values preserved); state_dict / load_state_dict keep working.
        # Build the optimizer BEFORE capturing a GraphedTrainStep (a captured step notices moved parameters and recaptures).
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
        self._has = tuple(True for _ in self.params)      # which parameters had a gradient at the last gather
        self._masks = {}                                  # tuple of parameter indices -> flat 0/1 mask selecting them

    def zero_grad(self, set_to_none: bool = True) -> None:
        for p in self.params:
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()

    def gather_grads(self) -> torch.Tensor:
        """Copies every `.grad` into the flat gradient buffer and returns it -- the buffer to all-reduce when
        training on several GPUs.  Parameters without a gradient are masked out of the update."""
        views = list(torch.split(self.grad, self.sizes))
        self._has = tuple(p.grad is not None for p in self.params)
        live = [(v, p.grad.reshape(-1)) for v, p in zip(views, self.params) if p.grad is not None]
        if live:
            torch._foreach_copy_([a for a, _ in live], [b for _, b in live])
        return self.grad

    @property
    def step_count(self) -> int:
        return max(self.steps) if self.steps else 0

    def _mask(self, idxs) -> torch.Tensor:
        m = self._masks.get(idxs)
        if m is None:
            m = torch.zeros_like(self.flat)
            for i, v in enumerate(torch.split(m, self.sizes)):
                if i in set(idxs):
                    v.fill_(1.0)
            self._masks[idxs] = m
        return m

    def _advance(self):
        """Counts this step for every parameter that has a gradient; returns [(step, mask | None)] -- one launch per
        distinct step count (one, without a mask, unless some parameter has ever been skipped)."""
        for i, h in enumerate(self._has):
            if h:
                self.steps[i] += 1
        groups = {}
        for i, h in enumerate(self._has):
            if h:
                groups.setdefault(self.steps[i], []).append(i)
        if len(groups) == 1 and all(self._has):
            return [(next(iter(groups)), None)]
        return [(st, self._mask(tuple(idxs))) for st, idxs in sorted(groups.items())]

    def _launch(self, lo: int, hi: int, grad_scale: float, step: int, mask: Optional[torch.Tensor]) -> None:
        sl = slice(lo, hi)
        with torch.cuda.device(self.flat.device):
            check(lib().cgnn_adam_step_masked(ptr(self.flat[sl]), ptr(self.grad[sl]), ptr(self.exp_avg[sl]), ptr(self.exp_avg_sq[sl]),
                                              ptr(None if mask is None else mask[sl]), hi - lo, self.lr, self.betas[0], self.betas[1],
                                              self.eps, self.weight_decay, step, float(grad_scale), stream_ptr(self.flat.device)),
                  "cgnn_adam_step")

    def step(self, grad_scale: float = 1.0, gathered: bool = False) -> None:
        """One Adam update with the current `lr`.  `gathered=True`: the flat gradient buffer already holds the
        (all-reduced) gradients from `gather_grads()`."""
        if not gathered:
            self.gather_grads()
        for st, mask in self._advance():
            self._launch(0, self.flat.numel(), grad_scale, st, mask)

    def step_overlapped(self, group=None, n_buckets: int = 4, average: bool = False) -> None:
        """Multi-GPU step: gathers the gradients, all-reduces (SUM) the flat buffer in `n_buckets` slices issued back to
        back on the communication stream, and updates each slice as soon as ITS reduction has landed -- the update of
        bucket i overlaps the reduction of bucket i+1.  `average` folds 1/world into the update (replica training);
        slab training sums (every rank holds a partial gradient of one global loss)."""
        import torch.distributed as dist
        self.gather_grads()
        launches = self._advance()
        n = self.flat.numel()
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        if world == 1:
            for st, mask in launches:
                self._launch(0, n, 1.0, st, mask)
            return
        per = -(-n // max(1, n_buckets))
        per = -(-per // 4) * 4                            # slices stay 16-byte aligned
        bounds = list(range(0, n, per)) + [n]
        works = [dist.all_reduce(self.grad[a:b], op=dist.ReduceOp.SUM, group=group, async_op=True)
                 for a, b in zip(bounds[:-1], bounds[1:])]
        scale = 1.0 / world if average else 1.0
        for (a, b), w in zip(zip(bounds[:-1], bounds[1:]), works):
            w.wait()                                      # the current stream waits for this bucket only
            for st, mask in launches:
                self._launch(a, b, scale, st, mask)

    # -- checkpoint / resume (torch.optim.Adam layout) ---------------------------------------------------------------
    def state_dict(self) -> dict:
        state = {}
        for i, (m, v, p) in enumerate(zip(torch.split(self.exp_avg, self.sizes), torch.split(self.exp_avg_sq, self.sizes), self.params)):
            if self.steps[i] > 0:                        # torch.optim.Adam creates a parameter's state at its first step
                state[i] = {"step": torch.tensor(float(self.steps[i])), "exp_avg": m.view_as(p).clone(), "exp_avg_sq": v.view_as(p).clone()}
        group = {"lr": self.lr, "betas": self.betas, "eps": self.eps, "weight_decay": self.weight_decay, "amsgrad": False,
                 "params": list(range(len(self.params)))}
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd: dict) -> None:
        group = sd["param_groups"][0]
        if len(group["params"]) != len(self.params):
            raise ValueError(f"optimizer state has {len(group['params'])} parameters, the model {len(self.params)}")
        self.lr = float(group["lr"])
        self.betas = (float(group["betas"][0]), float(group["betas"][1]))
        self.eps, self.weight_decay = float(group["eps"]), float(group["weight_decay"])
        self.steps = [int(float(sd["state"][i]["step"])) if i in sd["state"] else 0 for i in range(len(self.params))]
        self.exp_avg.zero_()
        self.exp_avg_sq.zero_()
        for i, (m, v) in enumerate(zip(torch.split(self.exp_avg, self.sizes), torch.split(self.exp_avg_sq, self.sizes))):
            st = sd["state"].get(i)
            if st is not None:
                m.copy_(st["exp_avg"].reshape(-1))
                v.copy_(st["exp_avg_sq"].reshape(-1))


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

    def state_dict(self) -> dict:
        return {"gamma": self.gamma, "base_lrs": [self.base_lr], "last_epoch": self.last_epoch, "_last_lr": [self.optimizer.lr]}

    def load_state_dict(self, sd: dict) -> None:
        self.gamma, self.base_lr, self.last_epoch = float(sd["gamma"]), float(sd["base_lrs"][0]), int(sd["last_epoch"])
        self.optimizer.lr = self.base_lr * self.gamma ** self.last_epoch


def save_training_state(path: str, model: torch.nn.Module, optimizer, scheduler=None, epoch: int = 0, extra: Optional[dict] = None) -> None:
    """Everything a bit-exact resume needs.  The reference saves `simulator.state_dict()` only (train.py:329-351), so a
    restarted run loses the Adam moments, restarts the learning-rate schedule and redraws the noise."""
    state = {"model": model.state_dict(), "optimizer": optimizer.state_dict(),
             "scheduler": None if scheduler is None else scheduler.state_dict(), "epoch": int(epoch),
             "rng_cpu": torch.get_rng_state(),
             "rng_cuda": torch.cuda.get_rng_state_all() if torch.cuda.is_available() else None,
             "extra": extra or {}}
    torch.save(state, path)


def load_training_state(path: str, model: torch.nn.Module, optimizer, scheduler=None, map_location=None) -> dict:
    """Restores what `save_training_state` wrote; returns {'epoch', 'extra'}.  The model's lazy layers must be
    materialised (one forward) and the optimizer built over them first."""
    state = torch.load(path, map_location=map_location, weights_only=False)
    model.load_state_dict(state["model"])
    optimizer.load_state_dict(state["optimizer"])
    if scheduler is not None and state.get("scheduler") is not None:
        scheduler.load_state_dict(state["scheduler"])
    torch.set_rng_state(state["rng_cpu"].cpu() if torch.is_tensor(state["rng_cpu"]) else state["rng_cpu"])
    if state.get("rng_cuda") is not None and torch.cuda.is_available():
        torch.cuda.set_rng_state_all([s.cpu() for s in state["rng_cuda"]])
    return {"epoch": state["epoch"], "extra": state.get("extra", {})}
