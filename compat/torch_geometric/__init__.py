"""Stand-in for the slice of `torch_geometric` the reference's callers touch (train.py:5-6,150-162,247;
validation.py:2,56): `torch_geometric.data.{Data,Batch}` and `torch_geometric.loader.DataLoader`.
Used only when PyTorch-Geometric itself is not installed (it is not in this image); a real install
found elsewhere on sys.path takes this module's place.  Nothing here is on the hot path: `Data` / `Batch` are the
containers of cosmology_gnn_simulation_b200/graph.py, `DataLoader` is torch's with PyG's collate rule."""
import os
import sys

from _shim import prefer_real  # noqa: E402

if not prefer_real(__name__):
    from . import data, loader  # noqa: F401
    __version__ = "0.0+cgnn-shim"
