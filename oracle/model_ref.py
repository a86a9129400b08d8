"""Oracle: encode-process-decode Interaction Network, restated as pure functions.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Works in fp32 or fp64 on
CPU tensors and is differentiable through torch autograd, so it is the truth
for outputs *and* gradients.

Follows the reference line by line, but over a flat parameter dict keyed by the
reference's own `state_dict` names, e.g.
`processor.3.edge_model.0.2.weight`:

* MLP            graph_network.py:15-32   Linear(+ReLU) x n_hidden, then Linear
* MLP + LN       graph_network.py:133-135 Sequential(mlp, LayerNorm(latent))
* encoder        graph_network.py:52-64
* one MP step    graph_network.py:83-101
* residuals      graph_network.py:177-183 (both updates read the OLD latents)
* decoders       graph_network.py:151-152,158-164 (no LayerNorm)
* loss           train.py:107-118,193,255-260

`message` selects what PyG's `propagate` sums at the receivers
(graph_network.py:92):
  "sender" - reference-actual: PyG 2.6.1's default `message(x_j)` returns the
             sender node latent and the `edge_attr=` kwarg is ignored
             (un-vendored dependency; semantics restated, parity unpinned).
  "edge"   - the intended Interaction Network: the updated edge latent
             (LayerNorm output, before the residual) is the message.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch

LN_EPS = 1e-5  # nn.LayerNorm default, graph_network.py:135


# ----------------------------------------------------------------------------
# parameters
# ----------------------------------------------------------------------------
def mlp_names(prefix: str, n_hidden: int):
    """Linear layers of `build_mlp` sit at Sequential indices 0,2,4,... (ReLU between)."""
    return [f"{prefix}.{2 * i}" for i in range(n_hidden + 1)]


def param_shapes(latent: int, hidden: int, n_hidden: int, n_steps: int, out_size: int,
                 node_in: int, edge_in: int) -> Dict[str, tuple]:
    """All tensors of the reference `state_dict`, in registration order
    (graph_network.py:137-152: encoder node, encoder edge, then per step node_model,
    edge_model (graph_network.py:80-81), then the two decoders)."""
    shapes: Dict[str, tuple] = {}

    def add_mlp(prefix, fan_in, fan_out):
        widths = [fan_in] + [hidden] * n_hidden + [fan_out]
        for name, (i, o) in zip(mlp_names(prefix, n_hidden), zip(widths[:-1], widths[1:])):
            shapes[name + ".weight"] = (o, i)
            shapes[name + ".bias"] = (o,)

    def add_mlp_ln(prefix, fan_in):
        add_mlp(prefix + ".0", fan_in, latent)
        shapes[prefix + ".1.weight"] = (latent,)
        shapes[prefix + ".1.bias"] = (latent,)

    add_mlp_ln("encoder.node_model", node_in)
    add_mlp_ln("encoder.edge_model", edge_in)
    for t in range(n_steps):
        add_mlp_ln(f"processor.{t}.node_model", 2 * latent)
        add_mlp_ln(f"processor.{t}.edge_model", 3 * latent)
    add_mlp("decoder_acc", latent, out_size)
    add_mlp("decoder_temp_rate", latent, 1)
    return shapes


def init_params(latent: int, hidden: int, n_hidden: int, n_steps: int, out_size: int,
                node_in: int = 17, edge_in: int = 4, seed: int = 0,
                dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """nn.Linear default init (U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for weight and bias),
    LayerNorm gamma=1 beta=0 - the distribution the reference modules draw from."""
    g = torch.Generator().manual_seed(seed)
    params = {}
    for name, shape in param_shapes(latent, hidden, n_hidden, n_steps, out_size, node_in, edge_in).items():
        parts = name.split(".")
        is_ln = parts[-2] == "1" and parts[-3] in ("node_model", "edge_model")
        if is_ln:
            params[name] = torch.ones(shape, dtype=dtype) if name.endswith("weight") else torch.zeros(shape, dtype=dtype)
        else:
            if name.endswith("weight"):
                fan_in = shape[1]
                last_fan_in = fan_in
            else:
                fan_in = last_fan_in
            bound = 1.0 / math.sqrt(fan_in)
            params[name] = ((torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * bound).to(dtype)
    return params


# ----------------------------------------------------------------------------
# forward
# ----------------------------------------------------------------------------
def mlp(params, prefix: str, z: torch.Tensor, n_hidden: int) -> torch.Tensor:
    names = mlp_names(prefix, n_hidden)
    for i, name in enumerate(names):
        z = z @ params[name + ".weight"].T + params[name + ".bias"]
        if i < n_hidden:
            z = torch.relu(z)
    return z


def layer_norm(z, gamma, beta):
    mu = z.mean(dim=-1, keepdim=True)
    var = ((z - mu) ** 2).mean(dim=-1, keepdim=True)  # biased, like nn.LayerNorm
    return (z - mu) / torch.sqrt(var + LN_EPS) * gamma + beta


def mlp_ln(params, prefix: str, z, n_hidden: int):
    y = mlp(params, prefix + ".0", z, n_hidden)
    return layer_norm(y, params[prefix + ".1.weight"], params[prefix + ".1.bias"])


def forward(params: Dict[str, torch.Tensor], x: torch.Tensor, edge_index: torch.Tensor,
            edge_attr: torch.Tensor, n_hidden: int, n_steps: int, message: str = "sender",
            return_latents: bool = False, checkpoint_steps: bool = False):
    """`EncodeProcessDecode.forward` (graph_network.py:154-164).  `checkpoint_steps` re-runs every processor
    step in backward instead of keeping its activations (same arithmetic, same gradients): what lets the float64
    oracle of a 32 768-particle box fit the host memory."""
    assert message in ("sender", "edge")
    src, dst = edge_index[0], edge_index[1]
    h = mlp_ln(params, "encoder.node_model", x, n_hidden)            # :54
    e = mlp_ln(params, "encoder.edge_model", edge_attr, n_hidden)    # :57

    def step(t, h, e):
        edge_in = torch.cat([h[src], h[dst], e], dim=-1)             # :89  sender, receiver, edge
        u_e = mlp_ln(params, f"processor.{t}.edge_model", edge_in, n_hidden)  # :90
        msg = h[src] if message == "sender" else u_e                 # :92 (see module docstring)
        agg = torch.zeros_like(h).index_add_(0, dst, msg)
        node_in = torch.cat([h, agg], dim=-1)                        # :94
        u_n = mlp_ln(params, f"processor.{t}.node_model", node_in, n_hidden)  # :96
        return h + u_n, e + u_e                                      # :181, :182

    for t in range(n_steps):
        if checkpoint_steps:
            from torch.utils.checkpoint import checkpoint
            h, e = checkpoint(step, t, h, e, use_reentrant=False)
        else:
            h, e = step(t, h, e)
    acc = mlp(params, "decoder_acc", h, n_hidden)                    # :158
    temp_rate = mlp(params, "decoder_temp_rate", h, n_hidden)        # :159
    out = {"acceleration": acc, "temp_rate": temp_rate}
    if return_latents:
        out["h"] = h
        out["e"] = e
    return out


# ----------------------------------------------------------------------------
# loss
# ----------------------------------------------------------------------------
def loss(acc_pred, temp_pred, y_acc, y_temp, dt: float, batch: Optional[torch.Tensor] = None,
         num_graphs: int = 1, w_acc: float = 1.0, w_temp: float = 1.0, w_mom: float = 0.0):
    """train.py:255-260 with `momentum_conservation_loss` (train.py:107-118).

    MSELoss() is the mean over ALL elements; the momentum term uses the *normalised*
    predicted accelerations times dt, summed per graph, squared norm, mean over graphs."""
    acc_loss = ((acc_pred - y_acc) ** 2).mean()
    temp_loss = ((temp_pred - y_temp) ** 2).mean()
    if batch is None:
        batch = torch.zeros(acc_pred.shape[0], dtype=torch.long)
    dv = acc_pred * dt
    tot = torch.zeros(num_graphs, acc_pred.shape[1], dtype=acc_pred.dtype).index_add_(0, batch, dv)
    mom = w_mom * (tot ** 2).sum() / num_graphs
    total = w_acc * acc_loss + w_temp * temp_loss + mom
    return {"loss": total, "acc_loss": acc_loss, "temp_rate_loss": temp_loss, "momentum_loss": mom}
