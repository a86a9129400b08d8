"""Oracle: `data_utils.preprocess` restated step by step.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Uses torch CPU float32 ops so that the
feature/target arithmetic is bit-compatible with the reference, and the numpy k-NN oracle for
the graph.  Returns a plain dict (no PyG).
"""
from __future__ import annotations

import numpy as np
import torch

from . import knn_ref


def _random_walk_noise(rate_seq: torch.Tensor, scale, dt: float) -> torch.Tensor:
    """Shared body of generate_position_noise / generate_temperature_noise
    (data_utils.py:36-54, 57-70): randn (drawn even when scale == 0), two cumsums over
    time, times dt, with a zero frame prepended."""
    steps = rate_seq.size(1)
    walk = torch.randn_like(rate_seq, dtype=torch.float32) * (scale / (steps ** 0.5))
    walk = walk.cumsum(dim=1)
    noise = walk.cumsum(dim=1) * dt
    return torch.cat((torch.zeros_like(noise, dtype=torch.float32)[:, 0:1], noise), dim=1)


def _min_image_(d: torch.Tensor, box: float) -> torch.Tensor:
    """In-place wrap with STRICT inequalities (data_utils.py:41-42, 104-105, 186-187)."""
    d[d < -1 * box / 2] += box
    d[d > box / 2] -= box
    return d


def preprocess(position_seq, temperature_seq, metadata, target_position=None,
               target_temperature=None, noise_std=0.0, num_neighbors=16, dt=None, box_size=None,
               knn="brute"):
    dt = float(dt)
    box = float(box_size)
    pos = position_seq.float().permute(1, 0, 2)                       # [N,W,3]   :86
    temp = temperature_seq.float()
    if temp.shape[0] == pos.shape[1] and temp.shape[1] == pos.shape[0]:
        temp = temp.permute(1, 0, 2)                                  # [N,W,1]   :87-88

    # :91-92 position noise (RNG draw #1) and wrap
    vel0 = _min_image_(pos[:, 1:] - pos[:, :-1], box) / dt
    pos_noise = _random_walk_noise(vel0, noise_std, dt)
    pos = torch.remainder(pos + pos_noise, box)

    # :95-97 temperature noise (RNG draw #2)
    trs = torch.tensor(metadata["temp_rate_std"], dtype=torch.float32)
    temp_noise = _random_walk_noise((temp[:, 1:] - temp[:, :-1]) / dt, noise_std * trs, dt)
    temp = temp + temp_noise

    recent = pos[:, -1]                                               # :100
    vel = _min_image_(pos[:, 1:] - pos[:, :-1], box) / dt             # :102-107
    recent_temp = temp[:, -1]                                         # :110

    f32 = lambda key: torch.tensor(metadata[key], dtype=torch.float32)
    nv = (vel - f32("vel_mean")) / f32("vel_std")                     # :127-129
    nt = (temp - f32("temp_mean")) / f32("temp_std")                  # :132-134
    x = torch.cat((nv.reshape(nv.size(0), -1), nt.reshape(nt.size(0), -1)), dim=-1)   # :138-145

    # :148-152 graph
    n = recent.shape[0]
    pos_np = recent.numpy()
    if knn == "brute":
        ext_idx = knn_ref.knn_brute(pos_np, box, num_neighbors)
    else:
        ext_idx = knn_ref.knn_kdtree(pos_np, box, num_neighbors)
    edge_index = torch.from_numpy(knn_ref.edge_index_from_ext(ext_idx, n))
    snd, rcv = edge_index[0], edge_index[1]

    # :162-164 edge features: RAW difference of wrapped positions (not minimum image)
    disp = recent[snd] - recent[rcv]
    edge_attr = torch.cat((disp, torch.norm(disp, dim=-1, keepdim=True)), dim=-1)

    y_acc = None
    if target_position is not None:                                   # :170-197
        tp = target_position.float()
        if tp.dim() == 3:
            tp = tp.permute(1, 0, 2).squeeze(1)
        elif tp.dim() == 2 and tp.shape[0] != n:
            tp = tp.reshape(-1, 3)
        tp = tp + pos_noise[:, -1]
        nxt = _min_image_(tp - recent, box) / dt
        y_acc = ((nxt - vel[:, -1]) / dt - f32("acc_mean")) / f32("acc_std")

    y_temp = None
    if target_temperature is not None:                                # :113-124, 199-214
        tt = target_temperature.float()
        if tt.dim() == 3:
            tt = tt.permute(1, 0, 2).squeeze(1)
        elif tt.dim() == 2 and tt.shape[1] != 1:
            tt = tt.reshape(-1, 1)
        if tt.shape != recent_temp.shape and tt.numel() == recent_temp.numel():
            tt = tt.reshape(recent_temp.shape)
        tt = tt + temp_noise[:, -1]
        y_temp = ((tt - recent_temp) / dt - f32("temp_rate_mean")) / f32("temp_rate_std")

    return {
        "x": x.float(), "edge_index": edge_index, "edge_attr": edge_attr,
        "y_acc": None if y_acc is None else y_acc.float(),
        "y_temp_rate": None if y_temp is None else y_temp.float(),
        "pos": recent, "dt": torch.tensor([dt], dtype=torch.float32),
        "box_size": torch.tensor([box], dtype=torch.float32),
        "ext_idx": torch.from_numpy(ext_idx),
    }
