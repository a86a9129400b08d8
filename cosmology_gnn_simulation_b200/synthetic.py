"""Synthetic particle boxes (the reference ships no data).

`make_box` returns what one `SequenceDataset` sample would hold (dataloader.py:99-131):
    Coordinates     [W+1, N, 3]   (W input frames + 1 target frame)
    InternalEnergy  [W+1, N, 1]
and a metadata dict with exactly the keys `generate_metadata.py:32-43` writes.
"""
from __future__ import annotations

import numpy as np
import torch


def positions(n: int, kind: str = "uniform", box: float = 1.0, seed: int = 0) -> np.ndarray:
    rng = np.random.default_rng(seed)
    if kind == "uniform":
        p = rng.random((n, 3), dtype=np.float32) * np.float32(box)
    elif kind == "clustered":
        # mixture of ~N/4096 isotropic Gaussians (sigma log-uniform in [0.002, 0.02] box units)
        # plus 20 % uniform background, wrapped into the box
        n_bg = n // 5
        n_cl = n - n_bg
        n_blobs = max(1, n // 4096)
        centres = rng.random((n_blobs, 3)) * box
        sigma = np.exp(rng.uniform(np.log(0.002), np.log(0.02), n_blobs)) * box
        which = rng.integers(0, n_blobs, n_cl)
        blob = centres[which] + rng.standard_normal((n_cl, 3)) * sigma[which, None]
        bg = rng.random((n_bg, 3)) * box
        p = np.concatenate([blob, bg], axis=0)
        p = np.mod(p, box).astype(np.float32)
        rng.shuffle(p, axis=0)
    elif kind == "lattice":
        m = int(round(n ** (1.0 / 3.0)))
        assert m ** 3 == n, "lattice needs a cubic particle count"
        g = (np.arange(m, dtype=np.float32) + 0.5) * np.float32(box / m)
        p = np.stack(np.meshgrid(g, g, g, indexing="ij"), axis=-1).reshape(-1, 3).astype(np.float32)
    else:
        raise ValueError(f"unknown box kind {kind!r}")
    p = np.where(p >= box, np.float32(0.0), p).astype(np.float32)
    return np.ascontiguousarray(p)


def make_box(n: int, kind: str = "uniform", window: int = 5, box: float = 1.0, dt: float = 0.01,
             seed: int = 0):
    """W+1 frames of positions advected by constant random velocities, log-normal internal energy."""
    rng = np.random.default_rng(seed + 1000)
    p0 = positions(n, kind, box, seed)
    v = (rng.standard_normal((n, 3)) * 0.1).astype(np.float32)
    frames = [p0]
    for _ in range(window):
        frames.append(np.mod(frames[-1] + v * np.float32(dt), np.float32(box)).astype(np.float32))
    coords = np.stack(frames, axis=0)
    # a little per-frame drift so that temperature rates are not identically zero
    u0 = np.exp(rng.standard_normal((n, 1)) * 0.5).astype(np.float32)
    du = (rng.standard_normal((window + 1, n, 1)) * 0.01).astype(np.float32).cumsum(axis=0)
    energy = (u0[None] * (1.0 + du)).astype(np.float32)
    # an acceleration field only for the metadata statistics
    acc = (rng.standard_normal((n, 3)) * 1.0).astype(np.float32)
    temp_rate = (energy[1:] - energy[:-1]) / dt
    metadata = {
        "temp_mean": np.mean(energy, axis=(0, 1)).tolist(),
        "temp_std": np.std(energy, axis=(0, 1)).tolist(),
        "temp_rate_mean": np.mean(temp_rate, axis=(0, 1)).tolist(),
        "temp_rate_std": np.std(temp_rate, axis=(0, 1)).tolist(),
        "vel_mean": float(np.mean(np.mean(v, axis=0))),
        "vel_std": float(np.mean(np.std(v, axis=0))),
        "acc_mean": float(np.mean(np.mean(acc, axis=0))),
        "acc_std": float(np.mean(np.std(acc, axis=0))),
        "box_size": float(box),
        "dt": float(dt),
    }
    return {
        "Coordinates": torch.from_numpy(coords),
        "InternalEnergy": torch.from_numpy(energy),
        "metadata": metadata,
    }
