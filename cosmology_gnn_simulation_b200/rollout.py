"""Device-resident autoregressive rollout (render_rollout.py:26-90).

Same arithmetic as the reference's `rollout`: build the graph from the last `window_size` frames (k-NN rebuilt
every step), predict, un-normalise, semi-implicit Euler (v' = v + a dt, x' = (x + v' dt) mod box, u' = u + du dt;
the velocity is the raw frame difference, as in the reference).  What changes is where things live: the
trajectory is one preallocated device tensor written in place (the reference re-`torch.cat`s a growing CPU
tensor every step and moves the graph to the device and the predictions back), the graph is built by the
GPU kernels, and nothing is synchronised with the host inside the loop.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from .data_utils import preprocess


@torch.no_grad()
def rollout(model, data: Dict[str, torch.Tensor], metadata: dict, noise_std: float, dt: float, box_size: float,
            window_size: int = 6, *, num_neighbors: int = 16, n_steps: Optional[int] = None,
            device=None) -> Dict[str, torch.Tensor]:
    """`data["Coordinates"]` [T,N,3] and `data["InternalEnergy"]` [T,N] or [T,N,1] give the first `window_size`
    frames (and T, the total length, unless `n_steps` says how many frames to predict).  Returns the full
    trajectories as device tensors, [T,N,3] and [T,N,1].  `num_neighbors` defaults to the reference's
    hard-coded 16 (render_rollout.py:49)."""
    dev = torch.device(device) if device is not None else next(model.parameters()).device
    model.eval()
    coords = data["Coordinates"]
    energy = data["InternalEnergy"]
    if energy.dim() == 2:
        energy = energy.unsqueeze(-1)
    w = int(window_size)
    total = coords.shape[0] if n_steps is None else w + int(n_steps)
    n = coords.shape[1]
    pos_traj = torch.empty((total, n, 3), dtype=torch.float32, device=dev)
    temp_traj = torch.empty((total, n, 1), dtype=torch.float32, device=dev)
    pos_traj[:w] = coords[:w].to(dev, dtype=torch.float32)
    temp_traj[:w] = energy[:w].to(dev, dtype=torch.float32)

    def md(key):
        return torch.tensor(metadata[key], dtype=torch.float32, device=dev)

    acc_std, acc_mean = md("acc_std"), md("acc_mean")
    rate_std, rate_mean = md("temp_rate_std"), md("temp_rate_mean")
    for t in range(w, total):
        graph = preprocess(pos_traj[t - w:t], temp_traj[t - w:t], metadata, noise_std=0.0,
                           num_neighbors=num_neighbors, box_size=box_size, dt=dt, device=dev)
        pred = model(graph)
        acc = pred["acceleration"] * acc_std + acc_mean
        rate = pred["temp_rate"] * rate_std + rate_mean
        recent = pos_traj[t - 1]
        velocity = (recent - pos_traj[t - 2]) / dt + acc * dt                   # render_rollout.py:72-76
        pos_traj[t] = torch.remainder(recent + velocity * dt, box_size)         # :77-80
        temp_traj[t] = temp_traj[t - 1] + rate * dt                             # :81
    return {"Coordinates": pos_traj, "InternalEnergy": temp_traj}


@torch.no_grad()
def rollout_slab(model, data: Dict[str, torch.Tensor], metadata: dict, noise_std: float, dt: float, box_size: float,
                 window_size: int = 6, *, num_neighbors: int = 16, n_steps: Optional[int] = None, rank: int = 0, world: int = 1,
                 group=None, device=None) -> Dict[str, torch.Tensor]:
    """`rollout` for one box slab-sharded over `world` ranks (BASELINE config 4; render_rollout.py:39-85 per rank).

    Every step the box is cut anew into equal-count x-slabs of the CURRENT positions (`preprocess_slab`), so a particle
    that has crossed a slab face simply belongs to its new owner in the next step: migration is the re-partition.  A rank
    predicts and integrates its own particles; one all-gather per step (positions + energy of the new frame, 16 bytes per
    particle) gives every rank the complete frame the next partition and k-NN need.  Every rank returns the full
    trajectories, identical on all ranks; with world == 1 the result equals `rollout` bit for bit."""
    import torch.distributed as dist
    from .data_utils import preprocess_slab
    from .slab import slab_bounds
    dev = torch.device(device) if device is not None else next(model.parameters()).device
    model.eval()
    coords = data["Coordinates"]
    energy = data["InternalEnergy"]
    if energy.dim() == 2:
        energy = energy.unsqueeze(-1)
    w = int(window_size)
    total = coords.shape[0] if n_steps is None else w + int(n_steps)
    n = coords.shape[1]
    pos_traj = torch.empty((total, n, 3), dtype=torch.float32, device=dev)
    temp_traj = torch.empty((total, n, 1), dtype=torch.float32, device=dev)
    pos_traj[:w] = coords[:w].to(dev, dtype=torch.float32)
    temp_traj[:w] = energy[:w].to(dev, dtype=torch.float32)

    def md(key):
        return torch.tensor(metadata[key], dtype=torch.float32, device=dev)

    acc_std, acc_mean = md("acc_std"), md("acc_mean")
    rate_std, rate_mean = md("temp_rate_std"), md("temp_rate_mean")
    bounds = slab_bounds(n, world)
    cap = max(bounds[p + 1] - bounds[p] for p in range(world))          # slabs differ by at most one particle: pad to the largest
    for t in range(w, total):
        graph = preprocess_slab(pos_traj[t - w:t], temp_traj[t - w:t], metadata, noise_std=0.0, num_neighbors=num_neighbors,
                                box_size=box_size, dt=dt, rank=rank, world=world, group=group, device=dev)
        pred = model(graph)
        lo, hi = graph.own_range
        own = graph.order[lo:hi]                                          # input-order ids of the particles this rank owns now
        acc = pred["acceleration"] * acc_std + acc_mean
        rate = pred["temp_rate"] * rate_std + rate_mean
        recent = pos_traj[t - 1][own]
        velocity = (recent - pos_traj[t - 2][own]) / dt + acc * dt           # render_rollout.py:72-76
        new = torch.zeros((cap, 4), dtype=torch.float32, device=dev)
        new[:hi - lo, :3] = torch.remainder(recent + velocity * dt, box_size)   # :77-80
        new[:hi - lo, 3:] = temp_traj[t - 1][own] + rate * dt                   # :81
        if world > 1:
            parts = torch.empty((world, cap, 4), dtype=torch.float32, device=dev)
            dist.all_gather_into_tensor(parts, new, group=group)
            for p in range(world):                                        # every rank knows every slab's ids: the partition is global
                ids = graph.order[bounds[p]:bounds[p + 1]]
                pos_traj[t][ids] = parts[p, :ids.numel(), :3]
                temp_traj[t][ids] = parts[p, :ids.numel(), 3:]
        else:
            pos_traj[t][own] = new[:, :3]
            temp_traj[t][own] = new[:, 3:]
    return {"Coordinates": pos_traj, "InternalEnergy": temp_traj}
