"""Fused training loss (train.py:107-118,193,255-260; duplicated at validation.py:5-16,63-69).

    total = w_acc * MSE(acc, y_acc) + w_temp * MSE(temp_rate, y_temp_rate)
            + w_mom / G * sum_g || dt * sum_{i in g} acc_i ||^2          (normalised predictions)

One library call (`cgnn_loss_fwd_bwd`) produces the four scalars and the gradient seeds with
deterministic two-stage reductions; nothing is synchronised with the host.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import ops


class _CombinedLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, acc, temp, y_acc, y_temp, graph_ptr, num_graphs, dt, w_acc, w_temp, w_mom):
        want = acc.requires_grad or temp.requires_grad
        losses, d_acc, d_temp = ops.loss_fwd_bwd(acc.contiguous(), temp.contiguous(), y_acc.contiguous(),
                                                 y_temp.contiguous(), graph_ptr, num_graphs, dt, w_acc, w_temp,
                                                 w_mom, want_grads=True)
        ctx.save_for_backward(d_acc, d_temp)
        del want
        return losses

    @staticmethod
    def backward(ctx, g):
        d_acc, d_temp = ctx.saved_tensors
        scale = g[0]            # only the total (component 0) is differentiable; 1..3 are reports
        return d_acc * scale, d_temp * scale, None, None, None, None, None, None, None, None


def graph_ptr_of(batch_graph, n: int, device) -> (Optional[torch.Tensor], int):
    """int32 node offsets [G+1] of a batched graph (None for a single graph)."""
    num_graphs = int(getattr(batch_graph, "num_graphs", 1) or 1)
    if num_graphs == 1:
        return None, 1
    ptr = getattr(batch_graph, "ptr", None)
    if ptr is None:
        counts = torch.bincount(batch_graph.batch, minlength=num_graphs)
        ptr = torch.cat([counts.new_zeros(1), counts.cumsum(0)])
    return ptr.to(device=device, dtype=torch.int32).contiguous(), num_graphs


def combined_loss(predictions: Dict[str, torch.Tensor], batch_graph, dt: float, acc_loss_weight: float = 1.0,
                  temp_rate_loss_weight: float = 1.0, momentum_loss_weight: float = 0.0) -> Dict[str, torch.Tensor]:
    """Returns {'loss', 'acc_loss', 'temp_rate_loss', 'momentum_loss'} as 0-d device tensors;
    'loss' carries the autograd graph."""
    acc, temp = predictions["acceleration"], predictions["temp_rate"]
    gptr, num_graphs = graph_ptr_of(batch_graph, acc.shape[0], acc.device)
    out = _CombinedLossFn.apply(acc, temp, batch_graph.y_acc, batch_graph.y_temp_rate, gptr, num_graphs,
                                float(dt), float(acc_loss_weight), float(temp_rate_loss_weight),
                                float(momentum_loss_weight))
    det = out.detach()
    return {"loss": out[0], "acc_loss": det[1], "temp_rate_loss": det[2], "momentum_loss": det[3]}


def momentum_conservation_loss(accelerations, batch_graph, dt, momentum_weight):
    """Signature of train.py:107 for callers that want only this term (autograd through torch ops)."""
    num_graphs = int(getattr(batch_graph, "num_graphs", 1) or 1)
    dv = accelerations * dt
    if num_graphs == 1:
        tot = dv.sum(dim=0, keepdim=True)
    else:
        tot = torch.zeros((num_graphs, dv.shape[1]), dtype=dv.dtype, device=dv.device).index_add_(0, batch_graph.batch, dv)
    return momentum_weight * (tot ** 2).sum() / num_graphs
