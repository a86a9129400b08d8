"""Graph containers with the slice of the PyG `Data` / `Batch` API the reference's callers use
(data_utils.py:218-227, train.py:247, validation.py:56, train.py:111-112).

They exist because torch_geometric is an optional dependency here; the model accepts any object
exposing `.x`, `.edge_index`, `.edge_attr` (a real PyG `Data` works too).  Private `_cgnn_*`
attributes carry the int32 ELL neighbour table through `.to()` and batching so the model does
not have to re-derive it from `edge_index`.
"""
from __future__ import annotations

from typing import Iterable, List

import torch


class Data:
    def __init__(self, **kwargs):
        for key, value in kwargs.items():
            setattr(self, key, value)

    # -- PyG-like helpers ---------------------------------------------------------------------
    def keys(self) -> List[str]:
        return [k for k, v in self.__dict__.items() if v is not None and not k.startswith("__")]

    def __contains__(self, key):
        return getattr(self, key, None) is not None

    @property
    def num_nodes(self):
        x = getattr(self, "x", None)
        if x is not None:
            return x.shape[0]
        pos = getattr(self, "pos", None)
        return None if pos is None else pos.shape[0]

    @property
    def num_edges(self):
        ei = getattr(self, "edge_index", None)
        return 0 if ei is None else ei.shape[1]

    def to(self, device, non_blocking: bool = False):
        out = self.__class__.__new__(self.__class__)
        for key, value in self.__dict__.items():
            if torch.is_tensor(value):
                value = value.to(device, non_blocking=non_blocking)
            out.__dict__[key] = value
        return out

    def cpu(self):
        return self.to("cpu")

    def cuda(self, device=None):
        return self.to("cuda" if device is None else device)

    def clone(self):
        out = self.__class__.__new__(self.__class__)
        for key, value in self.__dict__.items():
            out.__dict__[key] = value.clone() if torch.is_tensor(value) else value
        return out

    def __repr__(self):
        parts = []
        for key, value in self.__dict__.items():
            if key.startswith("_"):
                continue
            parts.append(f"{key}={list(value.shape)}" if torch.is_tensor(value) else f"{key}={value!r}")
        return f"{self.__class__.__name__}({', '.join(parts)})"


class Batch(Data):
    """Concatenation of graphs with node-offset edge indices, `batch` vector, `ptr`, `num_graphs`."""

    @classmethod
    def from_data_list(cls, data_list: Iterable[Data]) -> "Batch":
        data_list = list(data_list)
        if not data_list:
            raise ValueError("empty data list")
        out = cls()
        offsets = [0]
        for d in data_list:
            offsets.append(offsets[-1] + d.num_nodes)
        keys = [k for k in data_list[0].__dict__.keys()]
        for key in keys:
            vals = [getattr(d, key, None) for d in data_list]
            if key == "_cgnn_nbr_ext":          # per-graph ghost indices: meaningless after batching
                continue
            if key == "_cgnn_k":
                setattr(out, key, vals[0] if len(set(vals)) == 1 else None)
                continue
            if any(v is None for v in vals):
                setattr(out, key, None)
                continue
            if not torch.is_tensor(vals[0]):
                setattr(out, key, vals)
                continue
            if key == "edge_index":
                setattr(out, key, torch.cat([v + off for v, off in zip(vals, offsets)], dim=1))
            elif key == "_cgnn_senders":
                setattr(out, key, torch.cat([v + off for v, off in zip(vals, offsets)], dim=0))
            else:
                setattr(out, key, torch.cat(vals, dim=0))
        if getattr(out, "_cgnn_k", None) is None:
            out._cgnn_senders = None      # mixed in-degree: the model re-derives (and rejects) it
        dev = data_list[0].x.device if getattr(data_list[0], "x", None) is not None else "cpu"
        out.batch = torch.cat([torch.full((d.num_nodes,), i, dtype=torch.long, device=dev)
                               for i, d in enumerate(data_list)])
        out.ptr = torch.tensor(offsets, dtype=torch.long, device=dev)
        out.num_graphs = len(data_list)
        return out
