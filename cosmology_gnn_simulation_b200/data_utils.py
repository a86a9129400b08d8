"""Drop-in for the reference's `data_utils` module: same function names, arguments and outputs
(data_utils.py:9-33, 36-70, 72-228), with the graph construction moved to the GPU.

What changed underneath: the 27N ghost extension + `torch_cluster.knn` KD-tree (data_utils.py:148-149,
CPU, single thread) is replaced by the sm_100a cell-list kernel (`csrc/knn.cu`), and the index
remap / edge-feature gathers (data_utils.py:150-164) by one kernel (`csrc/graph.cu`).  The
feature/target arithmetic is kept as torch ops in the reference's order so it stays bit-compatible.

The returned graph lives on the CUDA device (a later `.to(device)` is a no-op).  There is no CPU
fallback: without a CUDA device `preprocess` raises.
"""
from __future__ import annotations

import torch

from . import ops
from .graph import Data

__all__ = ["extend_positions_torch", "generate_position_noise", "generate_temperature_noise", "preprocess", "preprocess_slab"]


def extend_positions_torch(positions, box_size):
    """27 periodic ghost copies [27N,3] and their mapping [27N] (data_utils.py:9-33).
    Kept for API compatibility; `preprocess` itself never materialises the ghosts."""
    n, d = positions.size()
    if isinstance(box_size, list):
        box_size = float(box_size[0])
    axis = torch.tensor([-box_size, 0, box_size], device=positions.device, dtype=torch.float32)
    shifts = torch.cartesian_prod(*([axis] * d))                      # x slowest, like the reference
    extended = (positions.unsqueeze(0) + shifts.unsqueeze(1)).reshape(-1, d)
    mapping = torch.arange(n, device=positions.device).repeat(shifts.shape[0])
    return extended, mapping


def _wrap_displacement_(disp, box_size):
    # the reference's two masked in-place updates (data_utils.py:44-45, 104-105), as selects: same values, but no
    # boolean-mask indexing (which synchronises with the host to size its result)
    disp = torch.where(disp < -1 * box_size / 2, disp + box_size, disp)
    disp = torch.where(disp > box_size / 2, disp - box_size, disp)
    return disp


def _double_cumsum_noise(rate_seq, scale, dt):
    steps = rate_seq.size(1)
    noise = torch.randn_like(rate_seq, dtype=torch.float32) * (scale / (steps ** 0.5))
    noise = noise.cumsum(dim=1).cumsum(dim=1) * dt
    return torch.cat((torch.zeros_like(noise, dtype=torch.float32)[:, 0:1], noise), dim=1)


def generate_position_noise(position_seq, noise_std, box_size, dt):
    """Random-walk position noise [N,W,3], zero at the first frame (data_utils.py:36-54).
    Draws from the torch RNG even when noise_std == 0, like the reference."""
    position_seq = position_seq.float()
    velocity = _wrap_displacement_(position_seq[:, 1:] - position_seq[:, :-1], box_size) / dt
    return _double_cumsum_noise(velocity, noise_std, dt)


def generate_temperature_noise(temperature_seq, noise_std, temp_rate_std, dt):
    """Random-walk temperature noise scaled by temp_rate_std (data_utils.py:57-70)."""
    temperature_seq = temperature_seq.float()
    rate = (temperature_seq[:, 1:] - temperature_seq[:, :-1]) / dt
    return _double_cumsum_noise(rate, noise_std * temp_rate_std, dt)


def _cuda_device(device):
    if device is not None:
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("cgnn preprocess builds the graph on a CUDA device; no CPU path exists")
        return device
    if not torch.cuda.is_available():
        raise RuntimeError("cgnn preprocess needs a CUDA device (the k-NN graph is built by libcgnn.so); "
                           "no CPU fallback exists")
    return torch.device("cuda", torch.cuda.current_device())


def _features_and_targets(position_seq, temperature_seq, metadata, target_position, target_temperature, noise_std, dt,
                          box_size):
    """Node features and normalised targets of one sample, in the reference's order of operations
    (data_utils.py:86-145,166-214).  Returns (recent_pos [N,3], x [N,F], y_acc | None, y_temp | None)."""
    def md(key):
        return torch.tensor(metadata[key], dtype=torch.float32, device=position_seq.device)

    pos = position_seq.float().permute(1, 0, 2)                                  # [N,W,3]
    temp = temperature_seq.float()
    if temp.shape[0] == pos.shape[1] and temp.shape[1] == pos.shape[0]:
        temp = temp.permute(1, 0, 2)                                             # [N,W,1]

    pos_noise = generate_position_noise(pos, noise_std, box_size, dt)
    pos = torch.remainder(pos + pos_noise, box_size)
    temp_noise = generate_temperature_noise(temp, noise_std, md("temp_rate_std"), dt)
    temp = temp + temp_noise

    recent_pos = pos[:, -1]
    velocity = _wrap_displacement_(pos[:, 1:] - pos[:, :-1], box_size) / dt
    recent_temp = temp[:, -1]
    n = recent_pos.shape[0]

    vel_feat = ((velocity - md("vel_mean")) / md("vel_std")).reshape(n, -1)
    temp_feat = ((temp - md("temp_mean")) / md("temp_std")).reshape(n, -1)
    x = torch.cat((vel_feat, temp_feat), dim=-1).float()

    y_acc = None
    if target_position is not None:
        tp = target_position.float()
        if tp.dim() == 3:
            tp = tp.permute(1, 0, 2).squeeze(1)
        elif tp.dim() == 2 and tp.shape[0] != n:
            tp = tp.reshape(-1, 3)
        tp += pos_noise[:, -1]                       # in place, like the reference (data_utils.py:182)
        next_vel = _wrap_displacement_(tp - recent_pos, box_size) / dt
        y_acc = (((next_vel - velocity[:, -1]) / dt - md("acc_mean")) / md("acc_std")).float()

    y_temp = None
    if target_temperature is not None:
        tt = target_temperature.float()
        if tt.dim() == 3:
            tt = tt.permute(1, 0, 2).squeeze(1)
        elif tt.dim() == 2 and tt.shape[1] != 1:
            tt = tt.reshape(-1, 1)
        if tt.shape != recent_temp.shape and tt.numel() == recent_temp.numel():
            tt = tt.reshape(recent_temp.shape)
        tt += temp_noise[:, -1]                      # in place (data_utils.py:206)
        y_temp = (((tt - recent_temp) / dt - md("temp_rate_mean")) / md("temp_rate_std")).float()
    return recent_pos, x, y_acc, y_temp


def preprocess(position_seq, temperature_seq, metadata, target_position=None, target_temperature=None,
               noise_std=0.0, num_neighbors=16, dt=None, box_size=None, *, edge_disp="raw", device=None):
    """Builds the graph for one sample (data_utils.py:72-228).

    position_seq [W,N,3], temperature_seq [W,N,1]; optional targets [1,N,3] / [1,N,1].
    Returns a `Data` with x, edge_index, edge_attr, y_acc, y_temp_rate, pos, dt, box_size on the GPU.
    `edge_disp="raw"` reproduces the reference's edge displacement (difference of wrapped positions,
    data_utils.py:162); `"min_image"` uses the periodic image the neighbour was found at.
    """
    dt = float(dt)
    box_size = float(box_size)
    dev = _cuda_device(device if device is not None else (position_seq.device if position_seq.is_cuda else None))
    recent_pos, x, y_acc, y_temp = _features_and_targets(position_seq, temperature_seq, metadata, target_position,
                                                         target_temperature, noise_std, dt, box_size)
    n = recent_pos.shape[0]

    # ---- graph: cell-list k-NN + edge features on the GPU --------------------------------------
    pos_dev = recent_pos.contiguous().to(dev, non_blocking=True)
    if n * 27 < num_neighbors:
        raise ValueError(f"num_neighbors={num_neighbors} exceeds the 27*N={27 * n} periodic candidates")
    nbr_ext = ops.knn_periodic(pos_dev, box_size, int(num_neighbors))
    senders, edge_index, edge_attr = ops.edge_features(pos_dev, nbr_ext, box_size, disp=edge_disp)
    assert edge_index.shape[1] == n * num_neighbors

    def mv(t):
        return None if t is None else t.contiguous().to(dev, non_blocking=True)

    graph = Data(
        x=mv(x), edge_index=edge_index, edge_attr=edge_attr, y_acc=mv(y_acc), y_temp_rate=mv(y_temp),
        pos=pos_dev, dt=torch.tensor([dt], dtype=torch.float32, device=dev),
        box_size=torch.tensor([box_size], dtype=torch.float32, device=dev),
    )
    graph._cgnn_senders = senders            # int32 ELL neighbour table (edge e = receiver*k + rank)
    graph._cgnn_k = int(num_neighbors)
    graph._cgnn_nbr_ext = nbr_ext
    return graph


def preprocess_slab(position_seq, temperature_seq, metadata, target_position=None, target_temperature=None,
                    noise_std=0.0, num_neighbors=16, dt=None, box_size=None, *, rank=0, world=1, group=None,
                    edge_disp="raw", device=None):
    """`preprocess` for one rank of a slab-sharded box (slab.py).  Every rank passes the SAME full sample;
    it gets back the graph of its own x-slab: `x`, targets and `pos` for the owned particles, the k in-edges
    of every owned particle (`edge_attr`, local int32 senders) and `halo` (the exchange plan).
    `graph.order` [N] is the x-sort permutation (global id -> index in the input arrays) and
    `graph.own_range` the owned global ids, so results can be scattered back to the input order."""
    from . import slab as _slab
    dt = float(dt)
    box_size = float(box_size)
    dev = _cuda_device(device if device is not None else (position_seq.device if position_seq.is_cuda else None))
    recent_pos, x, y_acc, y_temp = _features_and_targets(position_seq, temperature_seq, metadata, target_position,
                                                         target_temperature, noise_std, dt, box_size)
    n = recent_pos.shape[0]
    if n * 27 < num_neighbors:
        raise ValueError(f"num_neighbors={num_neighbors} exceeds the 27*N={27 * n} periodic candidates")
    pos_dev = recent_pos.contiguous().to(dev, non_blocking=True)
    order = torch.sort(pos_dev[:, 0], stable=True)[1]                 # identical on every rank (same data)
    pos_sorted = pos_dev[order].contiguous()
    bounds = _slab.slab_bounds(n, world)
    lo, hi = bounds[rank], bounds[rank + 1]
    nbr_ext = ops.knn_periodic(pos_sorted, box_size, int(num_neighbors), query_range=(lo, hi - lo))
    senders_g, _, edge_attr = ops.edge_features(pos_sorted, nbr_ext, box_size, disp=edge_disp, want_edge_index=False, q0=lo)
    plan, senders_local, halo_gid = _slab.plan_from_global_senders(senders_g, bounds, rank, world, group)
    own = order[lo:hi]

    def pick(t):
        return None if t is None else t.to(dev, non_blocking=True)[own].contiguous()

    graph = Data(
        x=pick(x), edge_index=None, edge_attr=edge_attr, y_acc=pick(y_acc), y_temp_rate=pick(y_temp),
        pos=pos_sorted[lo:hi].contiguous(), dt=torch.tensor([dt], dtype=torch.float32, device=dev),
        box_size=torch.tensor([box_size], dtype=torch.float32, device=dev),
    )
    graph._cgnn_senders = senders_local.contiguous()
    graph._cgnn_k = int(num_neighbors)
    graph.halo = plan
    graph.order = order
    graph.own_range = (lo, hi)
    graph.halo_gid = halo_gid
    return graph
