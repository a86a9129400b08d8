"""Drop-in for the reference's `data_utils` module: same function names, arguments and outputs
(data_utils.py:9-33, 36-70, 72-228), with the graph construction moved to the GPU.

What changed underneath: the 27N ghost extension + `torch_cluster.knn` KD-tree (data_utils.py:148-149,
CPU, single thread) is replaced by the sm_100a cell-list kernel (`csrc/knn.cu`), and the index
remap / edge-feature gathers (data_utils.py:150-164) by one kernel (`csrc/graph.cu`).  The
feature/target arithmetic (wrap, minimum-image velocities, normalisation, flattening, targets) is one kernel
(`csrc/features.cu`) that performs the reference's float32 operations one by one, bit-identically.

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


def _scalar(metadata, key) -> float:
    """A metadata statistic as the float32 the reference's `torch.tensor(metadata[key], dtype=float32)` holds
    (scalars or 1-element lists, generate_metadata.py:32-43)."""
    v = metadata[key]
    while isinstance(v, (list, tuple)):
        if len(v) != 1:
            raise ValueError(f"metadata[{key!r}] must be a scalar or a 1-element list")
        v = v[0]
    return float(torch.tensor(float(v), dtype=torch.float32))


def _noise(position_seq_nw, temperature_seq_nw, metadata, noise_std, dt, box_size):
    """The two random-walk noises (data_utils.py:36-70), drawn from the generator of the INPUT's device in the
    reference's order.  With noise_std == 0 the noise is identically zero but the reference still draws: the
    generator is advanced by the same two draws and no noise array is produced (None, None)."""
    if noise_std == 0:
        # the same `randn_like` calls on tensors of the same shape AND memory layout as the reference's velocity / rate
        # sequences (a generator fills a permuted tensor differently from a contiguous one)
        torch.randn_like(position_seq_nw[:, 1:] - position_seq_nw[:, :-1], dtype=torch.float32)
        torch.randn_like(temperature_seq_nw[:, 1:] - temperature_seq_nw[:, :-1], dtype=torch.float32)
        return None, None
    rate_std = torch.tensor(metadata["temp_rate_std"], dtype=torch.float32, device=position_seq_nw.device)
    pos_noise = generate_position_noise(position_seq_nw, noise_std, box_size, dt)
    temp_noise = generate_temperature_noise(temperature_seq_nw, noise_std, rate_std, dt)
    return pos_noise, temp_noise


def _features_and_targets(position_seq, temperature_seq, metadata, target_position, target_temperature, noise_std, dt,
                          box_size, dev):
    """Node features and normalised targets of one sample (data_utils.py:86-145,166-214) by ONE kernel launch
    (`cgnn_preprocess_features`, csrc/features.cu): bit-identical to the reference's CPU arithmetic.
    Returns device tensors (recent_pos [N,3], x [N,F], y_acc [N,3] | None, y_temp [N,1] | None)."""
    from ctypes import c_float
    from ._lib import check, lib, ptr, stream_ptr
    pos = position_seq.float()                                                   # [W,N,3], time-major
    w, n = pos.shape[0], pos.shape[1]
    temp = temperature_seq.float()
    if temp.shape[0] == w and temp.shape[1] == n:                                # the reference permutes exactly in this case (:87-88)
        temp_wn = temp.reshape(w, n)
    else:                                                                        # already [N,W,1]
        temp_wn = temp.reshape(n, w).t()
    pos_noise, temp_noise = _noise(pos.permute(1, 0, 2), temp_wn.t().unsqueeze(-1), metadata, float(noise_std), dt, box_size)

    def to_dev(t):
        return None if t is None else t.to(dev, non_blocking=True).contiguous()

    tp = tt = None
    if target_position is not None:
        tp = target_position.float()
        if tp.dim() == 3:
            tp = tp.permute(1, 0, 2).squeeze(1)
        elif tp.dim() == 2 and tp.shape[0] != n:
            tp = tp.reshape(-1, 3)
    if target_temperature is not None:
        tt = target_temperature.float()
        if tt.dim() == 3:
            tt = tt.permute(1, 0, 2).squeeze(1)
        tt = tt.reshape(n, 1) if tt.numel() == n else tt
    pos_d, temp_d, pn_d, tn_d, tp_d, tt_d = (to_dev(t) for t in (pos, temp_wn, pos_noise, temp_noise, tp, tt))
    f = 3 * (w - 1) + w
    recent_pos = torch.empty((n, 3), dtype=torch.float32, device=dev)
    x = torch.empty((n, f), dtype=torch.float32, device=dev)
    y_acc = torch.empty((n, 3), dtype=torch.float32, device=dev) if tp is not None else None
    y_temp = torch.empty((n, 1), dtype=torch.float32, device=dev) if tt is not None else None
    stats = (c_float * 8)(*[_scalar(metadata, k) for k in ("vel_mean", "vel_std", "temp_mean", "temp_std", "acc_mean", "acc_std",
                                                           "temp_rate_mean", "temp_rate_std")])
    with torch.cuda.device(dev):
        check(lib().cgnn_preprocess_features(ptr(pos_d), ptr(temp_d), ptr(pn_d), ptr(tn_d), ptr(tp_d), ptr(tt_d), n, w, float(box_size),
                                             float(dt), stats, ptr(recent_pos), ptr(x), ptr(y_acc), ptr(y_temp), stream_ptr(dev)),
              "cgnn_preprocess_features")
    if pos_noise is not None:
        # the reference adds the last noise frame to the caller's target tensors in place (data_utils.py:182,206);
        # done after the launch, so a target that already lives on the device is read un-noised by the kernel
        if tp is not None:
            tp += pos_noise[:, -1].to(tp.device)
        if tt is not None:
            tt += temp_noise[:, -1].to(tt.device)
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
                                                         target_temperature, noise_std, dt, box_size, dev)
    n = recent_pos.shape[0]

    # ---- graph: cell-list k-NN + edge features on the GPU --------------------------------------
    pos_dev = recent_pos
    if n * 27 < num_neighbors:
        raise ValueError(f"num_neighbors={num_neighbors} exceeds the 27*N={27 * n} periodic candidates")
    nbr_ext = ops.knn_periodic(pos_dev, box_size, int(num_neighbors))
    senders, edge_index, edge_attr = ops.edge_features(pos_dev, nbr_ext, box_size, disp=edge_disp)
    assert edge_index.shape[1] == n * num_neighbors

    graph = Data(
        x=x, edge_index=edge_index, edge_attr=edge_attr, y_acc=y_acc, y_temp_rate=y_temp,
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
    if world > 1 and float(noise_std) == 0.0 and not position_seq.is_cuda and position_seq.dim() == 3:
        # host inputs: every rank copies only its share of the particles over PCIe, the shares travel between the GPUs over
        # NVLink (slab.sharded_to_device) -- host-to-device bytes per rank do not grow with the number of ranks.  (With noise
        # the reference draws it on the INPUT's device, data_utils.py:36-70: host inputs then stay on the host until after.)
        w_, n_ = position_seq.shape[0], position_seq.shape[1]

        def up(t, pdim):
            return _slab.sharded_to_device(t.float(), pdim, rank, world, dev, group)

        position_seq = up(position_seq, 1)
        if not temperature_seq.is_cuda:
            temperature_seq = up(temperature_seq, 1 if (temperature_seq.shape[0] == w_ and temperature_seq.shape[1] == n_) else 0)
        if target_position is not None and not target_position.is_cuda:
            target_position = up(target_position, 1 if target_position.dim() == 3 else 0)
        if target_temperature is not None and not target_temperature.is_cuda:
            target_temperature = up(target_temperature, 1 if target_temperature.dim() == 3 else 0)
    recent_pos, x, y_acc, y_temp = _features_and_targets(position_seq, temperature_seq, metadata, target_position,
                                                         target_temperature, noise_std, dt, box_size, dev)
    n = recent_pos.shape[0]
    if n * 27 < num_neighbors:
        raise ValueError(f"num_neighbors={num_neighbors} exceeds the 27*N={27 * n} periodic candidates")
    pos_dev = recent_pos
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
