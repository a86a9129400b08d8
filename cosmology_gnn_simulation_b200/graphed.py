"""Whole-step CUDA-graph capture of a training application (forward + loss + backward).

A training step issues ~2000 kernel launches; at ~10 us of host time each the CPU needs about two thirds of
the GPU time just to enqueue them (`tools/e2e_breakdown.py`).  `GraphedTrainStep` captures the launches once
for a given graph SIZE (N, k) into a `torch.cuda.CUDAGraph` -- the kernels, their tensor maps and every
workspace live at fixed addresses inside the capture's private memory pool -- and replays them after copying the
next sample's tensors (x, edge_attr, senders, targets) into the captured input buffers.  The neighbour list is
data, not structure: the same captured step serves every sample with the same particle count.
"""
from __future__ import annotations

from typing import Callable, Dict

import torch

from .graph import Data


class GraphedTrainStep:
    def __init__(self, model, loss_fn: Callable, warmup: int = 3):
        """`loss_fn(predictions, graph) -> {'loss': scalar tensor, ...}` (e.g. a closure over combined_loss)."""
        self.model, self.loss_fn, self.warmup = model, loss_fn, warmup
        self.key = None
        self.graph = None
        self.launches_per_step = 0
        self._ws = None               # the capture's private workspaces (kept alive with the graph)
        self._grads = None            # the gradient tensors the captured kernels write
        self._param_ptrs = None

    def capture(self, g) -> None:
        """Captures the step for graphs of g's size.  Must run before the model has taken an eager backward on the
        legacy default stream (the gradient accumulators would otherwise be tied to it and invalidate the capture)."""
        from . import _lib
        dev = g.x.device
        self.static = Data(x=g.x.clone(), edge_index=None, edge_attr=g.edge_attr.clone(), y_acc=g.y_acc.clone(),
                           y_temp_rate=g.y_temp_rate.clone())
        self.static._cgnn_senders = g._cgnn_senders.clone()
        self.static._cgnn_k = g._cgnn_k
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        # the capture records raw workspace addresses: it gets workspaces of its own, which no later eager call
        # (a bigger validation graph, a rollout) can grow, free or move
        with _lib.workspace.scope() as held:
            with torch.cuda.stream(side):                   # warm-up on a side stream (lazy inits, workspaces, autotuned sizes)
                for _ in range(self.warmup):
                    for p in self.model.parameters():
                        p.grad = None
                    self.loss_fn(self.model(self.static), self.static)["loss"].backward()
            torch.cuda.current_stream(dev).wait_stream(side)
            for p in self.model.parameters():
                p.grad = None                               # gradients are (re)created inside the capture's pool
            self.graph = torch.cuda.CUDAGraph()
            l0 = _lib.launch_count()
            with torch.cuda.graph(self.graph):
                out = self.loss_fn(self.model(self.static), self.static)
                out["loss"].backward()
                self.losses = {k_: v.detach() for k_, v in out.items()}
            self.launches_per_step = _lib.launch_count() - l0
        self._ws = held
        params = list(self.model.parameters())
        self._grads = [p.grad for p in params]              # what the replayed kernels write into
        self._param_ptrs = [p.data_ptr() for p in params]   # what they read
        self.key = (tuple(g.x.shape), tuple(g.edge_attr.shape), g._cgnn_k)

    def __call__(self, g) -> Dict[str, torch.Tensor]:
        """Runs one training application on graph `g`; returns the loss dict (static tensors, valid until the
        next call).  Parameter gradients land in `.grad` as after a plain backward."""
        key = (tuple(g.x.shape), tuple(g.edge_attr.shape), g._cgnn_k)
        params = list(self.model.parameters())
        if key != self.key or [p.data_ptr() for p in params] != self._param_ptrs:
            # another graph size, or the parameters moved (an optimizer that re-points them into a flat buffer was built
            # after the capture): the captured kernels would read stale storage
            self.capture(g)
        self.static.x.copy_(g.x)
        self.static.edge_attr.copy_(g.edge_attr)
        self.static.y_acc.copy_(g.y_acc)
        self.static.y_temp_rate.copy_(g.y_temp_rate)
        self.static._cgnn_senders.copy_(g._cgnn_senders)
        self.graph.replay()
        for p, gr in zip(params, self._grads):              # `zero_grad(set_to_none=True)` detaches them: hand them back
            p.grad = gr
        return self.losses
